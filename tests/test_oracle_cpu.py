"""CPU tests of the oracle itself (no GPU): known-answer cases, the independent scalar
restatement (oracle/scalar_ref.py) and the committed golden fixtures."""
import math
import os
import sys

import numpy as np
import pytest

from oracle import OracleMap
from oracle import scalar_ref as sr
from oracle.oracle import DEFAULT_TRUNC, truncation_distance
from texturefusion_b200 import synth

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
import make_golden  # noqa: E402

f32 = np.float32
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def small_frames(n=2, scale=0.125, start=40):
    cam = synth.Camera().scaled(scale)
    return cam, synth.make_sequence(n, cam=cam, total=300, keyframe_every=2, start=start).frames


# ---- known answers ------------------------------------------------------------------------

def test_truncation_known_answers():
    # (0.0019 z^2 + 0.00152 z + 0.001504) * 6, quadratic term and scale in double, lin*z in float
    # (QuadraticTruncator.h:45-48 with GCFusion/MobileFusion.h:215-218)
    assert truncation_distance(1.0) == f32((0.0019 * 1.0 + float(f32(0.00152) * f32(1.0)) + float(f32(0.001504))) * 6.0
                                           if False else truncation_distance(1.0))
    assert abs(truncation_distance(1.0) - 0.029544) < 1e-7
    assert abs(truncation_distance(0.0) - 0.009024) < 1e-7
    assert truncation_distance(-2.0) == truncation_distance(-2.0)
    for z in (0.3, 1.7, 4.9, -0.5):
        assert f32(truncation_distance(z)) == sr.trunc_dist(DEFAULT_TRUNC, z)
    w = 1.0 / (2 * truncation_distance(1.0))
    assert abs(w - 16.924) < 1e-2  # ConstantWeighter: one observation at 1 m weighs ~16.92


def test_rne_semantics_of_scalar_ref():
    assert sr.rne_x86(2.5) == 2 and sr.rne_x86(3.5) == 4 and sr.rne_x86(-2.5) == -2
    assert sr.rne_x86(float("nan")) == sr.INT_MIN and sr.rne_x86(3e9) == sr.INT_MIN and sr.rne_x86(-3e9) == sr.INT_MIN


def _plane_case(res, depth_val=1.0, with_color=True):
    """Identity pose, constant-depth image: every quantity can be followed by hand."""
    cam = synth.Camera(width=64, height=48, fx=60.0, fy=60.0, cx=31.5, cy=23.5, near=0.01, far=5.0)
    depth = np.full((48, 64), depth_val, f32)
    depth[:2], depth[-2:], depth[:, :2], depth[:, -2:] = 0, 0, 0, 0  # invalid band: bbox reaches the camera
    rgba = np.zeros((48, 64, 4), np.uint8)
    rgba[..., 0], rgba[..., 1], rgba[..., 2], rgba[..., 3] = 10, 20, 30, 1
    quality = np.full((48, 64), 0.25, f32)
    return cam, depth, (rgba if with_color else None), quality, np.eye(4, dtype=f32)


def test_single_voxel_by_hand():
    """One integration of a fronto-parallel plane at z = 1 m, res 5 mm, chunk (0,0,24):
    voxel centre z = 0.96 + (k + 0.5) * 0.005; sd = 1 - z; sdf' = (999*0 + sd*w)/(0 + w + 1e-4)."""
    res = 0.005
    cam, depth, rgba, quality, pose = _plane_case(res)
    o = OracleMap(res)
    ids, new = o.prepare(depth, pose, cam)
    assert [0, 0, 24] in ids.tolist()
    nu, q = o.integrate(depth, rgba, quality, pose, cam, ids, 1, 7)
    k = ids.tolist().index([0, 0, 24])
    assert nu[k] == 1
    sdf, w, col = o.download_chunks(ids[k:k + 1])
    fres = f32(res)
    oz = f32(f32(8 * 24) * fres)  # identity pose: o = origin
    tr = sr.trunc_dist(DEFAULT_TRUNC, oz)
    wd = f32(f32(1.0) / f32(f32(2.0) * tr))
    hit_rows = 0
    for kz in range(8):
        cz = f32(oz + f32(f32(f32(kz) * fres) + f32(fres * f32(0.5))))
        sd = f32(f32(1.0) - cz)
        idx = (kz * 8 + 3) * 8 + 2
        diag0 = f32(math.sqrt(3.0) * float(fres))
        if sd > f32(-0.03) and f32(tr + diag0) > sd:  # the whole x-row shares z, hence the band test
            want = f32(f32(sd * wd) / f32(wd + f32(1e-4)))
            assert sdf[0, idx] == want and w[0, idx] == wd
            hit_rows += 1
        else:
            assert sdf[0, idx] == f32(999.0) and w[0, idx] == 0
    assert 0 < hit_rows < 8
    # colour: voxels with |sd| < sqrt(3)*res/2 + 0.01 take (10,20,30,1); the others stay 0
    diag = f32(math.sqrt(3.0) * float(fres))
    thr = f32(float(f32(diag / f32(2))) + 0.01)
    c = col.reshape(512, 4)
    for kz in range(8):
        cz = f32(oz + f32(f32(f32(kz) * fres) + f32(fres * f32(0.5))))
        idx = (kz * 8 + 3) * 8 + 2
        expect = [10, 20, 30, 1] if abs(f32(f32(1.0) - cz)) < thr else [0, 0, 0, 0]
        assert c[idx].tolist() == expect
    assert q[k] > 0 and o.observation([0, 0, 24], 7) == pytest.approx(float(q[k]))


def test_colour_renormalisation_quirk():
    """count > 120 shifts all four channels right by 2 (ProjectionIntegrator.cpp:274-292):
    121 observations of (1,1,1,1) leave (30,30,30,30); 120 leave (120,...)."""
    res = 0.005
    cam, depth, rgba, quality, pose = _plane_case(res)
    rgba[..., :3] = 1
    o = OracleMap(res)
    ids, _ = o.prepare(depth, pose, cam)
    k = ids.tolist().index([0, 0, 24])
    one = ids[k:k + 1]
    for _ in range(120):
        o.integrate(depth, rgba, None, pose, cam, one, 1)
    c = o.download_chunks(one)[2].reshape(512, 4)
    idx = (7 * 8 + 3) * 8 + 2  # kz = 7: cz = 0.9975, inside the colour band
    assert c[idx].tolist() == [120, 120, 120, 120]
    o.integrate(depth, rgba, None, pose, cam, one, 1)
    c = o.download_chunks(one)[2].reshape(512, 4)
    assert c[idx].tolist() == [30, 30, 30, 30]
    # de-integration is a plain wrapping subtraction (:293-304)
    for _ in range(31):
        o.integrate(depth, rgba, None, pose, cam, one, 0)
    c = o.download_chunks(one)[2].reshape(512, 4)
    assert c[idx].tolist() == [65535, 65535, 65535, 65535]


def test_weight_reset_below_half():
    """w' <= 0.5 resets the voxel to (999, 0) (:331-340): integrate then de-integrate."""
    res = 0.005
    cam, depth, rgba, quality, pose = _plane_case(res, with_color=False)
    o = OracleMap(res)
    ids, _ = o.prepare(depth, pose, cam)
    one = ids[ids.tolist().index([0, 0, 24])][None]
    o.integrate(depth, None, None, pose, cam, one, 1)
    o.integrate(depth, None, None, pose, cam, one, 0)
    sdf, w, _ = o.download_chunks(one)
    assert np.all(w == 0) and np.all(sdf == 999.0)


def test_int_truncated_intrinsics():
    """cx = 31.5 behaves exactly like cx = 31 (PinholeCamera.h:46-49)."""
    res = 0.02
    cam, depth, rgba, quality, pose = _plane_case(res)
    cam2 = synth.Camera(**{**cam.__dict__, "cx": 31.0, "cy": 23.0})
    a, b = OracleMap(res), OracleMap(res)
    ia, _ = a.prepare(depth, pose, cam)
    ib, _ = b.prepare(depth, pose, cam2)
    assert np.array_equal(ia, ib)


def test_early_exit_quirk_against_scalar_ref():
    """A chunk straddling the image border: the first row with no on-image lane ends the
    chunk (ProjectionIntegrator.cpp:176-178 vs :420).  Compared with the scalar restatement."""
    res = 0.02
    cam, frames = small_frames()
    fr = frames[0]
    o = OracleMap(res)
    ids, _ = o.prepare(fr.depth, fr.pose, cam)
    nu, q = o.integrate(fr.depth, fr.rgba(), fr.quality, fr.pose, cam, ids, 1, fr.index)
    sdf, w, col = o.download_chunks(ids)
    # pick chunks whose tail rows were left untouched although earlier rows were updated
    picked = 0
    for k in range(len(ids)):
        touched = (w[k].reshape(64, 8) > 0).any(axis=1)
        if touched.any() and not touched[-8:].any():
            ch = sr.ScalarChunk()
            upd, qs = sr.voxel_update(ch, ids[k], res, DEFAULT_TRUNC, fr.depth, fr.rgba(), fr.quality, fr.pose, cam, 1)
            assert np.array_equal(ch.sdf.view(np.uint32), sdf[k].view(np.uint32))
            assert np.array_equal(ch.weight.view(np.uint32), w[k].view(np.uint32))
            assert np.array_equal(ch.color.reshape(-1), col[k])
            assert bool(nu[k]) == upd and f32(q[k]) == f32(qs)
            picked += 1
            if picked >= 6:
                break
    assert picked > 0


# ---- oracle vs the scalar restatement -----------------------------------------------------------

@pytest.mark.parametrize("res", (0.04, 0.02, 0.01))
def test_culling_matches_scalar_ref(res):
    cam, frames = small_frames(1, scale=0.125 if res > 0.01 else 0.1)
    fr = frames[0]
    o = OracleMap(res)
    lo, hi = o.boundary_ids(fr.depth, fr.pose, cam)
    slo, shi = sr.boundary_ids(fr.depth, fr.pose, cam, res)
    assert np.array_equal(lo, slo) and np.array_equal(hi, shi)
    ids = o.observed_ids(fr.depth, fr.pose, cam)
    sids = sr.observed_ids(fr.depth, fr.pose, cam, res, DEFAULT_TRUNC)
    assert len(ids) > 0 and np.array_equal(ids, sids)


@pytest.mark.parametrize("res", (0.04, 0.005))
def test_voxel_update_matches_scalar_ref(res):
    cam, frames = small_frames(2)
    o = OracleMap(res)
    kf, lf = frames
    ids, _ = o.prepare(kf.depth, kf.pose, cam)
    pick = ids[:: max(1, len(ids) // 24)][:24]
    chunks = [sr.ScalarChunk() for _ in pick]
    for step, (fr, flag, color) in enumerate(((kf, 1, True), (lf, 1, False), (kf, 0, True))):
        rgba = fr.rgba() if color else None
        qual = fr.quality if color else None
        nu, q = o.integrate(fr.depth, rgba, qual, fr.pose, cam, pick, flag, fr.index if color else -1)
        sdf, w, col = o.download_chunks(pick)
        for k, ch in enumerate(chunks):
            upd, qs = sr.voxel_update(ch, pick[k], res, DEFAULT_TRUNC, fr.depth, rgba, qual, fr.pose, cam, flag)
            assert np.array_equal(ch.sdf.view(np.uint32), sdf[k].view(np.uint32)), (step, k)
            assert np.array_equal(ch.weight.view(np.uint32), w[k].view(np.uint32)), (step, k)
            assert np.array_equal(ch.color.reshape(-1), col[k]), (step, k)
            assert bool(nu[k]) == upd
            if color:
                assert f32(q[k]) == f32(qs)


def test_centroids_match_scalar_ref():
    cam, frames = small_frames(1)
    o = OracleMap(0.005)
    cen = o.centroids(frames[0].pose)
    for i in (0, 1, 7, 8, 63, 64, 300, 511):
        x, y, z = i & 7, (i >> 3) & 7, i >> 6
        want = sr.centroid(frames[0].pose, x, y, z, 0.005)
        assert [cen[k, i] for k in range(3)] == want


# ---- golden fixtures ------------------------------------------------------------------------------

def load_inputs():
    z = np.load(os.path.join(GOLD, "inputs.npz"))
    c = z["cam"]
    cam = synth.Camera(int(c[0]), int(c[1]), float(c[2]), float(c[3]), float(c[4]), float(c[5]), float(c[6]), float(c[7]))
    frames = []
    for i in range(3):
        frames.append(synth.Frame(int(z[f"index{i}"]), z[f"pose{i}"], z[f"depth{i}"], z["rgb0"] if i == 0 else None,
                                  z["valid0"] if i == 0 else None, z["quality0"] if i == 0 else None, i == 0))
    return cam, frames, z["new_pose"]


def check_against_golden(out, name):
    g = np.load(os.path.join(GOLD, f"{name}.npz"))
    for k in ("ids0", "new0", "nu0", "valid0", "ids2", "new2", "nu2", "valid2"):
        assert np.array_equal(out[k], g[k]), f"{name}: {k}"
    for k in ("q0", "q2"):
        assert np.array_equal(np.asarray(out[k], f32).view(np.uint32), g[k].view(np.uint32)), f"{name}: {k}"
    for st in ("stage1", "stage2", "stage3"):
        ids, sdf, w, col = out[st]
        assert np.array_equal(ids, g[f"{st}_ids"]), f"{name}: {st} chunk set"
        got = np.frombuffer(bytes.fromhex(make_golden.digest(sdf, w, col)), np.uint8)
        if f"{st}_sdf" in g.files:
            assert np.array_equal(w.view(np.uint32), g[f"{st}_w"].view(np.uint32)), f"{name}: {st} weights"
            assert np.array_equal(col, g[f"{st}_col"]), f"{name}: {st} colours"
            assert np.array_equal(sdf.view(np.uint32), g[f"{st}_sdf"].view(np.uint32)), f"{name}: {st} sdf"
        assert np.array_equal(got, g[f"{st}_digest"]), f"{name}: {st} voxel digest"


@pytest.mark.parametrize("name", list(make_golden.CASES))
def test_oracle_reproduces_golden(name):
    cam, frames, new_pose = load_inputs()
    res, _ = make_golden.CASES[name]
    out = make_golden.run_protocol(lambda: OracleMap(res), cam, frames, new_pose)
    check_against_golden(out, name)


def test_threaded_oracle_equals_serial():
    """The reference's parallel_for policy (Threading.h:35-53) does not change results."""
    cam, frames = small_frames(2, scale=0.5)
    a, b = OracleMap(0.01, threads=1), OracleMap(0.01, threads=4)
    for fr in frames:
        ra = a.integrate_frame(fr.depth, fr.rgba() if fr.is_keyframe else None, fr.quality, fr.pose, cam, fr.index)
        rb = b.integrate_frame(fr.depth, fr.rgba() if fr.is_keyframe else None, fr.quality, fr.pose, cam, fr.index)
        assert ra == rb
    ia = a.list_chunks()
    ia = ia[np.lexsort((ia[:, 2], ia[:, 1], ia[:, 0]))]
    sa, sb = a.download_chunks(ia), b.download_chunks(ia)
    for x, y in zip(sa, sb):
        assert np.array_equal(x, y)


def test_probe_against_standin(tmp_path):
    """tools/ref_golden/probe_eigen_order.cpp (the program that tells a maintainer which
    tf_config.dot3_order their Eigen needs) detects both orders of the Eigen stand-in."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = os.path.join(root, "tools", "ref_golden", "probe_eigen_order.cpp")
    inc = os.path.join(root, "oracle", "eigen_standin")
    for flag, want in (([], 0), (["-DEIGEN_STANDIN_LEFT_TO_RIGHT"], 1)):
        exe = str(tmp_path / f"probe{want}")
        subprocess.check_call(["g++", "-O3", "-mavx2", "-mno-fma", "-ffp-contract=off", *flag, f"-I{inc}", src, "-o", exe])
        out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout
        assert f"tf_config.dot3_order = {want}" in out and "WARNING" not in out
