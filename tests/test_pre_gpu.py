"""Frame pre-processing on the device against the CPU checkers, through the C ABI:
bit-exact planes for every loop of main.cpp:117-147, at full 640x480 against the live oracle (the
reference's own sources when the host's RSQRTPS is the Intel table, else the port), and against the
fixture the reference build wrote (tests/golden/pre_room.npz)."""
import os
import sys

import numpy as np
import pytest

from oracle import pre
from texturefusion_b200 import capi, synth

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_pre_golden  # noqa: E402

pytestmark = pytest.mark.gpu


class CamF:
    """float intrinsics for the pre-processing calls (no truncation)"""

    def __init__(self, fx, fy, cx, cy, width, height):
        self.fx, self.fy, self.cx, self.cy, self.width, self.height = float(fx), float(fy), float(cx), float(cy), width, height
        self.near, self.far = 0.01, 5.0


class GpuPre:
    """oracle.pre.Pre's methods on the CUDA library: every call uploads its planes into the frame store,
    runs the tf_pre_* entry point and reads the result back."""

    def __init__(self, width, height, dot3_order=0):
        self.m = capi.Map(0.02, width=width, height=height, max_frames=8, max_chunks=1 << 15, dot3_order=dot3_order)
        self.W, self.H = width, height
        self.next = 0

    def cam(self, camf):
        return CamF(*camf, self.W, self.H)

    def put(self, depth):
        self.next += 1
        self.m.upload_frame(self.next, np.ascontiguousarray(depth, np.float32), None, None)
        return self.next

    def normal_map(self, depth, camf):
        f = self.put(depth)
        self.m.pre_normal_map(f, self.cam(camf))
        return self.m.pre_download(f, normal=True)["normal"]

    def refine_keyframe_n(self, kf_depth, new_depth, T, camf, times):
        """`times` refinements of one key-frame by the same new frame (the weights live in the library)."""
        k, n = self.put(kf_depth), self.put(new_depth)
        out = []
        for _ in range(times):
            self.m.pre_refine_keyframe(k, n, T, self.cam(camf))
            r = self.m.pre_download(k, depth=True, weight=True)
            out.append((r["depth"], r["weight"]))
        return out

    def new_frame_path(self, kf_depth, new_depth, T, camf):
        """A non-key-frame as main.cpp:117-139 treats it: normal map, outlier rejection against the key-frame,
        grazing-angle rejection.  Returns (normal, depth after the rejection, normal after grazing, depth after grazing)."""
        k, n = self.put(kf_depth), self.put(new_depth)
        c = self.cam(camf)
        self.m.pre_normal_map(n, c)
        self.m.pre_refine_newframe(k, n, T, c)
        a = self.m.pre_download(n, depth=True, normal=True)
        self.m.pre_refine_depth_by_normal(n, c)
        b = self.m.pre_download(n, depth=True, normal=True)
        return a["normal"], a["depth"], b["normal"], b["depth"]

    def color(self, depth, rgb, camf):
        f = self.put(depth)
        self.m.pre_normal_map(f, self.cam(camf))
        self.m.pre_color_quality(f, rgb, self.cam(camf))
        r = self.m.pre_download(f, normal=True, color_valid=True, quality=True)
        return r["normal"], r["color_valid"], r["quality"], f


def eq(a, b):
    return np.array_equal(np.ascontiguousarray(a).view(np.uint8), np.ascontiguousarray(b).view(np.uint8))


def live_oracle(l2r=False):
    if pre.have("ref") and pre.Pre("ref").host_rsqrt_matches():
        return pre.Pre("ref", l2r)
    return pre.Pre("port", l2r)


def chain_on_gpu(G, camf, pose_k, dk, rgb_k, pose_n, dn):
    """make_pre_golden.run_chain's protocol on the device."""
    out = {}
    out["normal_k0"] = G.normal_map(dk, camf)
    T_kn = pre.relative_transform(pose_n, pose_k)
    T_nk = pre.relative_transform(pose_k, pose_n)
    (d1, w1), (d2, w2) = G.refine_keyframe_n(dk, dn, T_kn, camf, 2)
    out["kf_depth_1"], out["kf_weight_1"], out["kf_depth_2"], out["kf_weight_2"] = d1, w1, d2, w2
    out["normal_n"], out["new_depth"], out["normal_n_refined"], out["new_depth_refined"] = G.new_frame_path(d2, dn, T_nk, camf)
    nk, valid, quality, f = G.color(d2, rgb_k, camf)
    out["normal_k2"], out["color_valid"], out["quality"] = nk, valid, quality
    return out, f


def test_gpu_reproduces_reference_fixture():
    gold = np.load(os.path.join(HERE, "golden", "pre_room.npz"))
    inp = make_pre_golden.inputs()
    H, W = inp[2].shape
    out, _ = chain_on_gpu(GpuPre(W, H), *inp)
    for k in gold.files:
        assert eq(out[k], gold[k]), k


@pytest.mark.parametrize("l2r", [False, True])
def test_gpu_equals_live_oracle_at_full_resolution(l2r):
    cam = synth.Camera()
    seq = synth.make_sequence(7, cam=cam, total=300, keyframe_every=6)
    rng = np.random.default_rng(3)
    k, n = seq.frames[0], seq.frames[6]
    dk, dn = k.depth.copy(), n.depth.copy()
    for d in (dk, dn):
        d[d > 0] += rng.normal(0, 0.002, size=int((d > 0).sum())).astype(np.float32)
    camf = (np.float32(cam.fx), np.float32(cam.fy), np.float32(cam.cx), np.float32(cam.cy))
    inp = (camf, k.pose, dk, k.rgb, n.pose, dn)
    want = make_pre_golden.run_chain(live_oracle(l2r), *inp)
    G = GpuPre(cam.width, cam.height, dot3_order=1 if l2r else 0)
    got, f = chain_on_gpu(G, *inp)
    for key, v in want.items():
        assert eq(got[key], v), key
    assert (want["kf_weight_2"] == 2).sum() > 100_000 and 0 < want["color_valid"].sum() < want["color_valid"].size
    # the planes the fusion path reads are the ones just checked: fuse the key-frame with colour straight from
    # the store and compare the map with the oracle fed the CPU planes (RGBA pack of MobileFusion.cpp:151-162)
    if l2r:
        return
    from oracle import OracleMap
    from util import assert_maps_equal
    rgba = np.zeros((cam.height, cam.width, 4), np.uint8)
    v = want["color_valid"].astype(bool)
    rgba[..., :3][v] = k.rgb[v]
    rgba[..., 3] = want["color_valid"]
    o = OracleMap(0.02)
    G.m.integrate_frame(f, True, k.pose, cam)
    o.integrate_frame(want["kf_depth_2"], rgba, want["quality"], k.pose, cam, -1)
    assert assert_maps_equal(G.m, o, what="key-frame fused from the pre-processed planes")


def test_in_place_dependency_of_the_keyframe_refinement():
    """A key-frame pixel whose nearest sample lies in an earlier 8-pixel step must see the UPDATED depth
    (BasicAPI.cpp:597-600 reads the plane it is writing).  A scene made of depth edges in the new frame and a
    sideways camera shift makes many such pixels, including chains."""
    W, H = 640, 480
    rng = np.random.default_rng(9)
    camf = tuple(np.float32(x) for x in (525.0, 525.0, 319.5, 239.5))
    dk = (1.5 + 0.3 * rng.random((H, W))).astype(np.float32)
    dn = np.where(rng.random((H, W)) < 0.5, dk, dk + 0.25).astype(np.float32)  # every other pixel is an edge
    T = np.eye(4)
    T[0, 3], T[1, 3] = 0.02, 0.03  # samples move up-left: sources are earlier in raster order
    O = live_oracle()
    G = GpuPre(W, H)
    w = np.zeros_like(dk)
    d1, w1 = O.refine_keyframe(dk, w, dn, T, camf)
    d2, w2 = O.refine_keyframe(d1, w1, dn, T, camf)
    (g1, gw1), (g2, gw2) = G.refine_keyframe_n(dk, dn, T, camf, 2)
    assert eq(g1, d1) and eq(gw1, w1) and eq(g2, d2) and eq(gw2, w2)
    assert (w2 > 0).sum() > 10_000


def test_errors():
    G = GpuPre(640, 480)
    cam = CamF(525, 525, 319.5, 239.5, 640, 480)
    with pytest.raises(capi.TexFusionError):
        G.m.pre_normal_map(77, cam)                    # frame not in the store
    f = G.put(np.ones((480, 640), np.float32))
    with pytest.raises(capi.TexFusionError):
        G.m.pre_refine_depth_by_normal(f, cam)         # no normal map yet
    with pytest.raises(capi.TexFusionError):
        G.m.pre_color_quality(f, np.zeros((480, 640, 3), np.uint8), cam)
    with pytest.raises(capi.TexFusionError):
        G.m.pre_normal_map(f, CamF(525, 525, 319.5, 239.5, 320, 240))  # camera size != map
    # a re-upload forgets the frame's normal map and weights
    G.m.pre_normal_map(f, cam)
    G.m.upload_frame(f, np.ones((480, 640), np.float32), None, None)
    with pytest.raises(capi.TexFusionError):
        G.m.pre_refine_depth_by_normal(f, cam)


def test_raw_depth_conversion_and_bilateral_filter():
    """framePreprocess: 16-bit depth -> metres with the range cut (bit-exact), then cv::bilateralFilter(9, 0.03, 10)
    against cv2 — OpenCV's own accumulation order is SIMD-width dependent, so the bar is float rounding:
    2e-6 relative (stated in include/texfusion.h)."""
    cv2 = pytest.importorskip("cv2")
    cam = synth.Camera()
    fr = synth.make_sequence(1, cam=cam, total=300, keyframe_every=10).frames[0]
    rng = np.random.default_rng(4)
    metres = fr.depth + np.where(fr.depth > 0, rng.normal(0, 0.003, fr.depth.shape), 0).astype(np.float32)
    raw = np.clip(np.round(metres * 1000.0), 0, 65535).astype(np.uint16)
    raw[100:110, 200:260] = 9000  # beyond max_depth = 5 m: dropped
    G = GpuPre(cam.width, cam.height)
    G.m.pre_upload_depth_u16(1, raw, 1000.0, 5.0)
    got = G.m.pre_download(1, depth=True)["depth"]
    r = raw.copy()
    r[r.astype(np.float32) > np.float32(5.0) * np.float32(1000.0)] = 0
    want = r.astype(np.float32) / np.float32(1000.0)
    assert eq(got, want) and (want[100:110, 200:260] == 0).all()
    for (d, sc, ss) in ((9, 0.03, 10.0), (9, 0.03, 4.5), (7, 0.03, 10.0)):
        G.m.pre_upload_depth_u16(2, raw, 1000.0, 5.0)
        G.m.pre_bilateral(2, d, sc, ss)
        g = G.m.pre_download(2, depth=True)["depth"]
        w = cv2.bilateralFilter(want, d, sc, ss)
        err = np.abs(g - w) / np.maximum(np.abs(w), 1e-3)
        assert err.max() < 2e-6, (d, sc, ss, float(err.max()))
        assert np.abs(g - want).max() > 1e-4  # the filter did something
    # a constant image is copied
    flat = np.full((cam.height, cam.width), 1.25, np.float32)
    G.m.upload_frame(3, flat)
    G.m.pre_bilateral(3)
    assert eq(G.m.pre_download(3, depth=True)["depth"], flat)


def test_empty_frames_holes_and_identity_transform():
    """Edge inputs: an all-zero frame, frames with large holes (depth 0 -> 0/0 projections), the identity
    transform (a frame refined against itself), a camera with integer principal point."""
    W, H = 640, 480
    rng = np.random.default_rng(21)
    camf = tuple(np.float32(x) for x in (525.0, 525.0, 320.0, 240.0))
    O = live_oracle()
    G = GpuPre(W, H)
    base = (1.0 + 0.5 * rng.random((H, W))).astype(np.float32)
    holes = base.copy()
    holes[rng.random((H, W)) < 0.3] = 0.0
    holes[100:200, 300:500] = 0.0
    zero = np.zeros((H, W), np.float32)
    I = np.eye(4)
    T = np.eye(4)
    T[:3, 3] = (0.01, -0.02, 0.015)
    for kd, nd, X in ((holes, base, T), (base, holes, T), (zero, base, T), (base, zero, T), (holes, holes, I)):
        assert eq(G.normal_map(kd, camf), O.normal_map(kd, camf))
        w = np.zeros_like(kd)
        d1, w1 = O.refine_keyframe(kd, w, nd, X, camf)
        (g1, gw1), = G.refine_keyframe_n(kd, nd, X, camf, 1)
        assert eq(g1, d1) and eq(gw1, w1)
        n0, a, n1, b = G.new_frame_path(d1, nd, np.linalg.inv(X), camf)
        on = O.normal_map(nd, camf)
        oa = O.refine_newframe(d1, nd, np.linalg.inv(X), camf)
        on1, ob = O.refine_depth_by_normal(on, oa, camf)
        assert eq(n0, on) and eq(a, oa) and eq(n1, on1) and eq(b, ob)
