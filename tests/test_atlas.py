"""Texture atlas (Structure/Atlas.cpp:32-91): slot placement, copy and resize.
CPU part: the oracle against hand-computed placement and against cv2.resize (the OpenCV code the
reference calls is not in its tree; +-1 LSB is the north_star tolerance).  GPU part: the CUDA
atlas against the oracle (bit-exact) and cv2 (+-1 LSB)."""
import numpy as np
import pytest

from oracle import OracleMap
from texturefusion_b200 import synth

ATLAS = 96 * 72 * 2


def keyframe(scale=1.0):
    cam = synth.Camera() if scale == 1.0 else synth.Camera().scaled(scale)
    return cam, synth.make_sequence(1, cam=cam, total=300, keyframe_every=1, start=50).frames[0]


def boxes_for(rng, w, h, pw, ph, n):
    """Crops that fit the slot, crops that need shrinking in x, in y, in both, and edge cases."""
    out = [(0, 0, pw, ph), (0, 0, 1, 1), (w - 2 * pw - 1, h - 2 * ph - 1, 2 * pw + 1, 2 * ph + 1), (5, 7, pw + 1, ph - 1),
           (9, 3, pw - 1, ph + 1), (0, 0, w - 1, h - 1)]
    while len(out) < n:
        bw, bh = int(rng.randint(1, 4 * pw)), int(rng.randint(1, 4 * ph))
        bw, bh = min(bw, w - 1), min(bh, h - 1)
        out.append((int(rng.randint(0, w - bw)), int(rng.randint(0, h - bh)), bw, bh))
    return out


@pytest.mark.parametrize("res,pw,ph", [(0.005, 24, 18), (0.01, 48, 36), (0.02, 96, 72), (0.04, 192, 144)])
def test_patch_size_and_slot_placement(res, pw, ph):
    """PATCH_WIDTH = floor(4800 res), PATCH_HEIGHT = floor(3600 res) (Atlas.h:62-65); AddPatch walks
    the atlas row by row and wraps when x + PATCH_WIDTH >= 13824 (Atlas.cpp:43-64)."""
    o = OracleMap(res)
    assert o.atlas_patch_size() == (pw, ph)
    per_row = 0
    x = 0
    while True:  # replay the reference's advance rule
        per_row += 1
        if x + pw >= ATLAS:
            break
        x += pw
    locs = [o.atlas_alloc_slot((i, 0, 0)) for i in range(per_row + 2)]
    assert locs[0] == 0 and locs[1] == pw
    assert locs[per_row] == ph * ATLAS and locs[per_row + 1] == ph * ATLAS + pw
    assert o.atlas_alloc_slot((1, 0, 0)) == locs[1]  # an existing patch keeps its slot


def test_oracle_resize_matches_cv2():
    import cv2
    cam, kf = keyframe()
    res = 0.005
    o = OracleMap(res)
    pw, ph = o.atlas_patch_size()
    rng = np.random.RandomState(1)
    worst = 0
    for i, (x, y, w, h) in enumerate(boxes_for(rng, cam.width, cam.height, pw, ph, 40)):
        loc = o.atlas_alloc_slot((i, 1, 2))
        o.atlas_update(loc, kf.rgb, (x, y, w, h))
        oy, ox = loc // ATLAS, loc % ATLAS
        rows = o.atlas_download(oy * ATLAS, (oy + ph) * ATLAS).reshape(ph, ATLAS, 3)
        crop = kf.rgb[y:y + h, x:x + w]
        if w > pw or h > ph:
            want = cv2.resize(crop, (pw, ph), interpolation=cv2.INTER_LINEAR)
            got = rows[:, ox:ox + pw]
        else:
            want, got = crop, rows[:h, ox:ox + w]
        diff = np.abs(got.astype(int) - want.astype(int)).max()
        worst = max(worst, diff)
        assert diff <= 1, f"box {(x, y, w, h)}: max diff {diff}"
    print("worst |oracle - cv2| =", worst)


@pytest.mark.gpu
@pytest.mark.parametrize("res", (0.005, 0.02))
def test_gpu_atlas_matches_oracle_and_cv2(res):
    import cv2
    from texturefusion_b200 import capi
    from texturefusion_b200.chisel import Chisel
    cam, kf = keyframe()
    c = Chisel(voxelResolution=res)
    o = OracleMap(res)
    pw, ph = c.atlas.PATCH_WIDTH, c.atlas.PATCH_HEIGHT
    assert (pw, ph) == o.atlas_patch_size()
    c.map.upload_frame(kf.index, kf.depth, None, kf.quality)
    c.map.upload_keyframe_rgb(kf.index, kf.rgb, kf.color_valid)
    rng = np.random.RandomState(2)
    boxes = boxes_for(rng, cam.width, cam.height, pw, ph, 300)
    ids = [(i, -i, 3) for i in range(len(boxes))]
    for cid, box in zip(ids, boxes):
        p = c.atlas.AddPatch(cid)
        assert p["texloc"] == o.atlas_alloc_slot(cid)
        c.atlas.SetPatchImage(cid, kf.index, box)
        o.atlas_update(p["texloc"], kf.rgb, box)
    c.UpdateAtlas(ids)
    last = c.atlas._patches[ids[-1]]["texloc"]
    hot_end = (last // ATLAS + ph) * ATLAS  # Chisel::GeneratePatches hot range (Chisel.cpp:184-186)
    got = c.atlas.texture_rows(0, hot_end)
    want = o.atlas_download(0, hot_end)
    assert np.array_equal(got, want), "atlas bytes differ from the oracle"
    # the same rows device to device (what a CUDA-GL pixel-unpack buffer receives, MobileFusion.h:404-427)
    import torch
    half = (hot_end // 2 // ATLAS) * ATLAS
    pbo = torch.zeros((hot_end - half) * 3 + 64, dtype=torch.uint8, device="cuda")
    c.map.atlas_copy_to_device(half, hot_end, pbo.data_ptr())
    back = pbo.cpu().numpy()
    assert np.array_equal(back[:(hot_end - half) * 3], want.reshape(-1)[half * 3:]) and not back[(hot_end - half) * 3:].any()
    with pytest.raises(capi.TexFusionError):
        c.map.atlas_copy_to_device(0, hot_end, got.ctypes.data)  # a host pointer
    img = got.reshape(-1, ATLAS, 3)
    for cid, (x, y, w, h) in list(zip(ids, boxes))[:60]:
        loc = c.atlas._patches[cid]["texloc"]
        oy, ox = loc // ATLAS, loc % ATLAS
        crop = kf.rgb[y:y + h, x:x + w]
        if w > pw or h > ph:
            ref = cv2.resize(crop, (pw, ph), interpolation=cv2.INTER_LINEAR)
            mine = img[oy:oy + ph, ox:ox + pw]
        else:
            ref, mine = crop, img[oy:oy + h, ox:ox + w]
        assert np.abs(mine.astype(int) - ref.astype(int)).max() <= 1
    # rgb upload also packs the RGBA plane on the device (MobileFusion.cpp:151-162): integrate with it
    st, *_ = c.map.integrate_frame(kf.index, True, kf.pose, cam)
    n, nupd = o.integrate_frame(kf.depth, kf.rgba(), kf.quality, kf.pose, cam, -1)
    assert (st.n_chunks, st.n_updated) == (n, nupd)
    from util import assert_maps_equal
    assert assert_maps_equal(c.map, o, what="device-packed RGBA")
    # atlas errors
    with pytest.raises(capi.TexFusionError):
        c.map.atlas_update([(0, kf.index, 600, 400, 100, 100)])  # bbox outside the image
    with pytest.raises(capi.TexFusionError):
        c.map.atlas_update([(0, 999, 0, 0, 4, 4)])  # unknown key-frame
