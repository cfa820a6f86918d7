"""Frame pre-processing, CPU side: the scalar oracle port against the reference's own
BasicAPI.cpp loops (oracle/_ref/libtexfusion_ref_pre.so), the RSQRTPS emulation against the host
instruction, the restated cvtColor / Sobel against cv2, and the committed fixtures."""
import os

import numpy as np
import pytest

from oracle import pre
from texturefusion_b200 import synth

HERE = os.path.dirname(os.path.abspath(__file__))

CAM = synth.Camera()
CAMF = (CAM.fx, CAM.fy, CAM.cx, CAM.cy)
needs_ref = pytest.mark.skipif(not pre.have("ref"), reason="oracle/_ref/libtexfusion_ref_pre.so not built")


def noisy_frames(n=3, seed=5, noise=0.002):
    """Frames of the room orbit with depth noise (so that normals, thresholds and the 5 % tests all have
    borderline pixels), 6 frames apart."""
    seq = synth.make_sequence(6 * n, cam=CAM, total=300, keyframe_every=6)
    rng = np.random.default_rng(seed)
    out = []
    for fr in seq.frames[::6]:
        d = fr.depth.copy()
        d[d > 0] += rng.normal(0, noise, size=int((d > 0).sum())).astype(np.float32)
        out.append((fr, d.astype(np.float32)))
    return out


def ref_host_ok():
    return pre.have("ref") and pre.Pre("ref").host_rsqrt_matches()


@needs_ref
def test_rsqrt_table_is_the_host_instruction():
    """Every 997th positive normal float, zero, infinities and a denormal: table emulation == RSQRTSS."""
    ref, port = pre.Pre("ref"), pre.Pre("port")
    if not ref.host_rsqrt_matches():
        pytest.skip("this host's RSQRTPS is not the Intel table (AMD?)")
    for b in list(range(0x00800000, 0x7f800000, 997 * 4099)) + [0, 0x7f800000, 0x00000001, 0x3f800000, 0x40000000]:
        assert port.rsqrt_bits(b) == ref.rsqrt_bits(b), hex(b)


@needs_ref
@pytest.mark.parametrize("l2r", [False, True])
def test_port_equals_reference_sources(l2r):
    """All six loops, chained as main.cpp:117-147 chains them, bit for bit."""
    if not ref_host_ok():
        pytest.skip("host RSQRTPS differs from the Intel table")
    R, P = pre.Pre("ref", l2r), pre.Pre("port", l2r)
    frames = noisy_frames()
    (kf, kd), (nf, nd) = frames[0], frames[1]
    # normal maps
    for _, d in frames:
        a, b = R.normal_map(d, CAMF), P.normal_map(d, CAMF)
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    n_new = R.normal_map(nd, CAMF)
    # key-frame refinement from a later frame, then rejection of the new frame's outliers
    T_ref_to_new = pre.relative_transform(nf.pose, kf.pose)
    T_new_to_ref = pre.relative_transform(kf.pose, nf.pose)
    w0 = np.zeros_like(kd)
    for rep in range(2):  # twice: the second pass starts from non-zero weights
        dr, wr = R.refine_keyframe(kd, w0, nd, T_ref_to_new, CAMF)
        dp, wp = P.refine_keyframe(kd, w0, nd, T_ref_to_new, CAMF)
        assert np.array_equal(dr.view(np.uint32), dp.view(np.uint32)) and np.array_equal(wr, wp)
        assert (wr > w0).sum() > 50_000  # the test is not vacuous
        kd, w0 = dr, wr
    a, b = R.refine_newframe(kd, nd, T_new_to_ref, CAMF), P.refine_newframe(kd, nd, T_new_to_ref, CAMF)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    assert 1000 < (a == 0).sum() < a.size
    # grazing-angle rejection
    (na, da), (nb, db) = R.refine_depth_by_normal(n_new, a, CAMF), P.refine_depth_by_normal(n_new, a, CAMF)
    assert np.array_equal(na.view(np.uint32), nb.view(np.uint32)) and np.array_equal(da.view(np.uint32), db.view(np.uint32))
    # colour validity + quality of a key-frame
    nk = R.normal_map(kd, CAMF)
    fa, fb = R.color_valid(nk, CAMF), P.color_valid(nk, CAMF)
    assert np.array_equal(fa, fb) and 0 < fa.sum() < fa.size
    qa, qb = R.color_quality(kd, nk, kf.rgb, CAMF), P.color_quality(kd, nk, kf.rgb, CAMF)
    assert np.array_equal(qa.view(np.uint32), qb.view(np.uint32))


def test_gray_and_sobel_match_opencv():
    """cv::cvtColor(RGB2GRAY) and cv::Sobel(CV_32F, 1, 1) as restated in the checkers == cv2 (4.x)."""
    cv2 = pytest.importorskip("cv2")
    P = pre.Pre("port")
    rng = np.random.default_rng(0)
    rgb = rng.integers(0, 256, size=(96, 128, 3), dtype=np.uint8)
    g = P.gray(rgb)
    assert np.array_equal(g, cv2.cvtColor(rgb, cv2.COLOR_RGB2GRAY))
    assert np.array_equal(P.sobel11(g), cv2.Sobel(g, cv2.CV_32F, 1, 1))
    # every grey level of the three primaries
    ramp = np.zeros((3, 256, 3), np.uint8)
    for c in range(3):
        ramp[c, :, c] = np.arange(256)
    assert np.array_equal(P.gray(ramp), cv2.cvtColor(ramp, cv2.COLOR_RGB2GRAY))
    if pre.have("ref"):
        R = pre.Pre("ref")
        assert np.array_equal(R.gray(rgb), g) and np.array_equal(R.sobel11(g), P.sobel11(g))


def test_port_against_committed_fixtures():
    """tests/golden/pre_*.npz were written by the reference build (tests/golden/make_pre_golden.py)."""
    path = os.path.join(HERE, "golden", "pre_room.npz")
    if not os.path.exists(path):
        pytest.skip("fixture not generated")
    import sys
    sys.path.insert(0, os.path.join(HERE, "golden"))
    from make_pre_golden import run_chain, inputs
    g = np.load(path)
    out = run_chain(pre.Pre("port"), *inputs())
    for k, v in out.items():
        assert np.array_equal(np.asarray(v).view(np.uint8), g[k].view(np.uint8)), k
