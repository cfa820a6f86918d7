"""The restated oracle (oracle/tf_oracle.cpp) against the REFERENCE'S OWN SOURCES
(oracle/_ref/libtexfusion_ref.so: ProjectionIntegrator.cpp, ChunkManager.{h,cpp}, Chunk, truncator,
weighter, camera compiled from /root/reference against the Eigen stand-in; oracle/Makefile).

This is what pins the oracle: ordered chunk lists, created / updated flags, raw observation-quality
sums, every voxel (sdf, weight, colour) and meshesToUpdate must be bit-identical.  Skipped where
the library has not been built (it needs /root/reference at build time; the built .so travels).
"""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
import make_golden  # noqa: E402

from oracle import OracleMap, have_ref, set_dot3_order
from oracle.oracle import truncation_distance
from test_oracle_cpu import check_against_golden, load_inputs
from texturefusion_b200 import synth
from util import RESOLUTIONS, sort_ids

pytestmark = pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built (needs /root/reference)")


def bits(a):
    return a.view(np.uint32) if a.dtype == np.float32 else a


def assert_same_map(a, b, what):
    la, _ = sort_ids(a.list_chunks())
    lb, _ = sort_ids(b.list_chunks())
    assert np.array_equal(la, lb), f"{what}: chunk sets differ"
    for x, y, name in zip(a.download_chunks(la), b.download_chunks(lb), ("sdf", "weight", "colour")):
        assert np.array_equal(bits(x), bits(y)), f"{what}: {name} not bit-identical"
    ma, _ = sort_ids(a.meshes_to_update())
    mb, _ = sort_ids(b.meshes_to_update())
    assert np.array_equal(ma, mb), f"{what}: meshesToUpdate differ"


def drive_both(a, b, seq, cam, keyframe_every, raw_quality=True):
    """Key-frame + local-frames protocol of ReIntegrateKeyframe (GCFusion/MobileFusion.cpp:114-221) on both.
    raw_quality: compare the per-chunk quality sums too (the real chisel::Chisel does not hand them out)."""
    frames = seq.frames
    for k0 in range(0, len(frames), keyframe_every):
        kf, local = frames[k0], frames[k0 + 1:k0 + keyframe_every]
        ia, na = a.prepare(kf.depth, kf.pose, cam)
        ib, nb = b.prepare(kf.depth, kf.pose, cam)
        assert np.array_equal(ia, ib), f"frame {kf.index}: list (order included) differs"
        assert np.array_equal(na, nb)
        ua, qa = a.integrate(kf.depth, kf.rgba(), kf.quality, kf.pose, cam, ia, 1, kf.index)
        ub, qb = b.integrate(kf.depth, kf.rgba(), kf.quality, kf.pose, cam, ib, 1, kf.index)
        assert np.array_equal(ua, ub) and (not raw_quality or np.array_equal(bits(qa), bits(qb))), f"frame {kf.index}"
        for lf in local:
            ua, _ = a.integrate(lf.depth, None, None, lf.pose, cam, ia, 1, -1, ua)
            ub, _ = b.integrate(lf.depth, None, None, lf.pose, cam, ib, 1, -1, ub)
            assert np.array_equal(ua, ub), f"local frame {lf.index}"
        va, vb = a.finalize(ia, ua, na), b.finalize(ib, ub, nb)
        assert np.array_equal(va, vb)
    return va


@pytest.mark.parametrize("res", RESOLUTIONS)
def test_port_matches_reference_sources(res):
    cam = synth.Camera().scaled(0.5)
    seq = synth.make_sequence(9, cam=cam, total=300, keyframe_every=3, start=30, noise_sigma=0.001)
    a, b = OracleMap(res, impl="port"), OracleMap(res, impl="ref")
    valid = drive_both(a, b, seq, cam, 3)
    assert_same_map(a, b, f"res {res} after fusion")
    # loop closure: de-integrate the last key-frame group under its old poses over validChunks
    kf = seq.frames[6]
    ua = np.ones(len(valid), np.uint8)
    ub = ua.copy()
    ua, qa = a.integrate(kf.depth, kf.rgba(), kf.quality, kf.pose, cam, valid, 0, kf.index, ua)
    ub, qb = b.integrate(kf.depth, kf.rgba(), kf.quality, kf.pose, cam, valid, 0, kf.index, ub)
    assert np.array_equal(ua, ub) and np.array_equal(bits(qa), bits(qb))
    assert_same_map(a, b, f"res {res} after de-integration")
    for i in range(0, len(valid), max(1, len(valid) // 50)):  # chunk->observations (Structure/Chisel.h:244-247)
        for kfid in (seq.frames[0].index, kf.index):
            assert a.observation(valid[i], kfid) == b.observation(valid[i], kfid)


@pytest.mark.parametrize("res", (0.04, 0.03))
def test_restated_glue_equals_the_reference_chisel_object(res):
    """The ~60 lines of Structure/Chisel.h that oracle/ref_driver.cpp restates (PrepareIntersectChunks,
    IntegrateDepthScanColor, FinalizeIntegrateChunks, GarbageCollect, bufferIntegratorSIMDCentroids) against the
    reference's OWN chisel::Chisel compiled with the OpenCV / Sophus stand-ins (impl "ref_chisel"): lists, flags,
    voxels, observations and meshesToUpdate, on lists below the reference's threading threshold (1000 chunks)."""
    if not have_ref("ref_chisel"):
        pytest.skip("oracle/_ref/libtexfusion_ref_chisel.so not built")
    cam = synth.Camera()
    seq = synth.make_sequence(9, cam=cam, total=300, keyframe_every=3, start=30, noise_sigma=0.001)
    a, b = OracleMap(res, impl="ref"), OracleMap(res, impl="ref_chisel")
    assert len(a.prepare(seq.frames[0].depth, seq.frames[0].pose, cam)[0]) < 1000
    a.reset(), b.reset()
    valid = drive_both(a, b, seq, cam, 3, raw_quality=False)
    assert_same_map(a, b, f"res {res} after fusion")
    kf = seq.frames[6]
    u = np.ones(len(valid), np.uint8)
    ua, _ = a.integrate(kf.depth, kf.rgba(), kf.quality, kf.pose, cam, valid, 0, kf.index, u.copy())
    ub, _ = b.integrate(kf.depth, kf.rgba(), kf.quality, kf.pose, cam, valid, 0, kf.index, u.copy())
    assert np.array_equal(ua, ub)
    assert_same_map(a, b, f"res {res} after de-integration")
    n_obs = 0
    for i in range(0, len(valid), max(1, len(valid) // 80)):
        for kfid in (seq.frames[0].index, kf.index):
            oa, ob = a.observation(valid[i], kfid), b.observation(valid[i], kfid)
            assert oa == ob
            n_obs += oa is not None
    assert n_obs > 0
    # the convenience form (IntegrateFrame) on a fresh pair
    a, b = OracleMap(res, impl="ref"), OracleMap(res, impl="ref_chisel")
    for fr in seq.frames:
        rg, q = (fr.rgba(), fr.quality) if fr.is_keyframe else (None, None)
        assert a.integrate_frame(fr.depth, rg, q, fr.pose, cam, fr.index if fr.is_keyframe else -1) == \
            b.integrate_frame(fr.depth, rg, q, fr.pose, cam, fr.index if fr.is_keyframe else -1)
    assert_same_map(a, b, f"res {res} convenience form")


def test_port_matches_reference_sources_full_frames_5mm():
    cam = synth.Camera()
    seq = synth.make_sequence(4, cam=cam, total=300, keyframe_every=2, start=100)
    a, b = OracleMap(0.005, impl="port"), OracleMap(0.005, impl="ref")
    drive_both(a, b, seq, cam, 2)
    assert_same_map(a, b, "640x480 at 5 mm")


def test_convenience_form_and_thread_policy():
    """IntegrateFrame shape (Structure/Chisel.h:453-468) with the reference's own parallel_for."""
    cam = synth.Camera()
    seq = synth.make_sequence(3, cam=cam, total=300, keyframe_every=10)
    a, b = OracleMap(0.005, impl="port", threads=0), OracleMap(0.005, impl="ref", threads=0)
    for fr in seq.frames:
        rg = fr.rgba() if fr.is_keyframe else None
        q = fr.quality if fr.is_keyframe else None
        ra = a.integrate_frame(fr.depth, rg, q, fr.pose, cam, fr.index if fr.is_keyframe else -1)
        rb = b.integrate_frame(fr.depth, rg, q, fr.pose, cam, fr.index if fr.is_keyframe else -1)
        assert ra == rb
    assert_same_map(a, b, "threaded convenience form")


@pytest.mark.parametrize("name", list(make_golden.CASES))
def test_reference_sources_reproduce_golden(name):
    cam, frames, new_pose = load_inputs()
    res, _ = make_golden.CASES[name]
    out = make_golden.run_protocol(lambda: OracleMap(res, impl="ref"), cam, frames, new_pose)
    check_against_golden(out, name)


@pytest.mark.parametrize("name", list(make_golden.CASES_L2R))
def test_left_to_right_association(name):
    """Eigen 3.2 association: the restatement with the runtime switch == the reference sources built
    with -DEIGEN_STANDIN_LEFT_TO_RIGHT == the *_l2r golden, and it differs from the default."""
    cam, frames, new_pose = load_inputs()
    res = make_golden.CASES_L2R[name]
    set_dot3_order(True)
    try:
        out_p = make_golden.run_protocol(lambda: OracleMap(res, impl="port"), cam, frames, new_pose)
    finally:
        set_dot3_order(False)
    out_r = make_golden.run_protocol(lambda: OracleMap(res, impl="ref_l2r"), cam, frames, new_pose)
    check_against_golden(out_p, name)
    check_against_golden(out_r, name)
    default = np.load(os.path.join(os.path.dirname(__file__), "golden", name.replace("_l2r", "") + ".npz"))
    l2r = np.load(os.path.join(os.path.dirname(__file__), "golden", name + ".npz"))
    assert not np.array_equal(default["stage1_digest"], l2r["stage1_digest"]), "the switch changes nothing?"


def test_truncator_is_the_reference_class():
    for z in (0.0, 0.3, 1.0, 2.5, 4.99, -1.0):
        assert truncation_distance(z, impl="ref") == truncation_distance(z, impl="port")
