"""Patch::CalculateTexCoords (Structure/Patch.cpp:40-108): oracle known answers on the CPU, CUDA
kernel against the oracle (bit-exact floats) on the GPU."""
import numpy as np
import pytest

from oracle import patch_texcoords as oracle_patch
from texturefusion_b200 import synth


def keyframe():
    cam = synth.Camera()
    return cam, synth.make_sequence(1, cam=cam, total=300, keyframe_every=1, start=70).frames[0]


def pseudo_meshes(cam, kf, n_patches=200, seed=3):
    """Vertices back-projected from the key-frame's own depth (so that most of them map correctly),
    plus patches that leave the image, one with wrong colours, one empty."""
    rng = np.random.RandomState(seed)
    R, t = kf.pose[:3, :3].astype(np.float64), kf.pose[:3, 3].astype(np.float64)
    offs, verts, cols = [0], [], []
    H, W = kf.depth.shape
    for p in range(n_patches):
        n = 0 if p == 7 else int(rng.randint(3, 120))
        cu, cv = rng.uniform(20, W - 20), rng.uniform(20, H - 20)
        spread = 60.0 if p % 17 == 0 else 12.0  # some patches straddle the image border
        u = cu + rng.uniform(-spread, spread, n) + (W * 0.6 if p % 23 == 5 else 0)
        v = cv + rng.uniform(-spread, spread, n)
        ui, vi = np.clip(np.round(u).astype(int), 0, W - 1), np.clip(np.round(v).astype(int), 0, H - 1)
        d = kf.depth[vi, ui].astype(np.float64)
        d = np.where(d > 0, d, 1.5) + (rng.uniform(-1.5, 1.5, n) if p % 29 == 3 else rng.normal(0, 0.002, n))
        pc = np.stack([(u - cam.cx) / cam.fx * d, (v - cam.cy) / cam.fy * d, d], 1)
        verts.append((pc @ R.T + t).astype(np.float32))
        c = kf.rgb[vi, ui].astype(np.float32) / 255.0 + rng.normal(0, 0.02, (n, 3)).astype(np.float32)
        if p % 31 == 11:
            c = 1.0 - c  # wrong colours -> wrong_mapping
        cols.append(c.astype(np.float32))
        offs.append(offs[-1] + n)
    return np.array(offs, np.int64), np.concatenate(verts), np.concatenate(cols)


def world_to_camera(pose):
    return np.linalg.inv(pose.astype(np.float64)).astype(np.float32)  # SE3d inverse, then cast<float>


def test_oracle_known_answers():
    cam, kf = keyframe()
    T = world_to_camera(kf.pose)
    # one vertex exactly on the optical axis, 2 m away: pixel = (int(cx) + 0.5, int(cy) + 0.5)
    p = (kf.pose[:3, :3].astype(np.float64) @ np.array([0.0, 0.0, 2.0]) + kf.pose[:3, 3]).astype(np.float32)
    tc, col, res = oracle_patch(kf.rgb, kf.depth, T, cam, [0, 1], p[None], np.zeros((1, 3), np.float32))
    x, y, w, h, wrong, flag = res[0]
    assert flag == 0 and (w, h) == (5, 5)            # Box(min-2, min-2, 0+5, 0+5)
    assert abs(tc[0, 0] + x - 319.5) < 1e-3 and abs(tc[0, 1] + y - 239.5) < 1e-3   # truncated cx=319, +0.5
    # texcolor = bilinear sample / 255 with the c2-for-c4 slip; at fraction (.5,.5): (c1 + 2 c2 + c3)/4
    c1, c2, c3 = kf.rgb[239, 319].astype(np.float32), kf.rgb[239, 320].astype(np.float32), kf.rgb[240, 319].astype(np.float32)
    assert np.allclose(col[0] * 255.0, (c1 + 2 * c2 + c3) / 4, atol=0.51)
    # a vertex behind the image border is clamped and flagged
    q = (kf.pose[:3, :3].astype(np.float64) @ np.array([5.0, 0.0, 2.0]) + kf.pose[:3, 3]).astype(np.float32)
    tc, col, res = oracle_patch(kf.rgb, kf.depth, T, cam, [0, 1], q[None], np.zeros((1, 3), np.float32))
    assert res[0][5] == -1 and tc[0, 0] + res[0][0] == cam.width
    # an empty patch: no box, not wrong
    tc, col, res = oracle_patch(kf.rgb, kf.depth, T, cam, [0, 0], np.zeros((0, 3), np.float32), np.zeros((0, 3), np.float32))
    assert res[0].tolist() == [0, 0, 0, 0, 0, 0]


def test_oracle_wrong_mapping_vote():
    cam, kf = keyframe()
    offs, v, c = pseudo_meshes(cam, kf)
    tc, col, res = oracle_patch(kf.rgb, kf.depth, world_to_camera(kf.pose), cam, offs, v, c)
    wrong = res[:, 4]
    assert wrong[11] == 1 and wrong[42] == 1          # inverted colours (p % 31 == 11)
    assert wrong[3] == 1                               # depth off by up to 1.5 m (p % 29 == 3)
    assert wrong[[0, 1, 2, 4, 6]].sum() == 0
    assert (res[:, 5] == -1).any() and (res[:, 5] == 0).any()
    assert res[7].tolist() == [0, 0, 0, 0, 0, 0]


@pytest.mark.gpu
def test_gpu_patch_texcoords_bit_exact():
    from texturefusion_b200 import capi
    cam, kf = keyframe()
    offs, v, c = pseudo_meshes(cam, kf, n_patches=600, seed=9)
    T = world_to_camera(kf.pose)
    g = capi.Map(0.005)
    g.upload_frame(kf.index, kf.depth, None, kf.quality)
    g.upload_keyframe_rgb(kf.index, kf.rgb, kf.color_valid)
    tc, col, res = g.patch_texcoords(kf.index, T, cam, offs, v, c)
    otc, ocol, ores = oracle_patch(kf.rgb, kf.depth, T, cam, offs, v, c)
    assert np.array_equal(res, ores)
    assert np.array_equal(tc.view(np.uint32), otc.view(np.uint32))
    assert np.array_equal(col.view(np.uint32), ocol.view(np.uint32))
    # the boxes feed straight into the atlas update
    patches = [(g.atlas_alloc_slot((i, 0, 9)), kf.index, *r[:4]) for i, r in enumerate(res) if r[2] > 0 and r[3] > 0]
    g.atlas_update(patches)
    with pytest.raises(capi.TexFusionError):
        g.patch_texcoords(12345, T, cam, offs, v, c)


def test_restatement_equals_reference_patch_sources():
    """Structure/Patch.cpp's own CalculateTexCoords / bilinear / bilinear_depth (oracle/_ref/libtexfusion_ref_patch.so)
    against the restatement the GPU kernel is checked with: texcoords, texcolours, boxes, votes, flags, bit for bit."""
    from oracle import have_patch_ref
    if not have_patch_ref():
        pytest.skip("oracle/_ref/libtexfusion_ref_patch.so not built")
    cam, kf = keyframe()
    T = world_to_camera(kf.pose)
    for seed in (3, 4):
        off, v, c = pseudo_meshes(cam, kf, n_patches=300, seed=seed)
        a = oracle_patch(kf.rgb, kf.depth, T, cam, off, v, c, impl="port")
        b = oracle_patch(kf.rgb, kf.depth, T, cam, off, v, c, impl="ref")
        assert np.array_equal(a[2], b[2])
        assert np.array_equal(a[0].view(np.uint32), b[0].view(np.uint32))
        assert np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32))
        assert (a[2][:, 4] == 1).sum() > 0 and (a[2][:, 5] == -1).sum() > 0  # wrong_mapping and off-image cases occur
