"""Parity at the sizes BASELINE.json names (the CUDA path through the C ABI against the CPU oracle):

  configs[1]  the full 300-frame 640x480 room sequence at 5 mm voxels, frame by frame
              (MobileFusion::IntegrateFrame, GCFusion/MobileFusion.cpp:223-250)
  configs[2]  loop closure: 50 key-frames x 7-frame groups fused under drifted poses, then
              de-integrated and re-integrated under corrected poses in ONE tf_integrate_batch call
              (GCFusion/MobileFusion.cpp:301-310); tools/bench_loopclosure.py --verify is the same check
              at 500 key-frames
  configs[3]  a 500-frame slice of the building walk with a chunk pool > 2^19 and tens of thousands
              of chunks per frame (hash load, tombstone churn, slot recycling over a long run)

Maps are compared chunk by chunk through texturefusion_b200.maphash (a 64-bit position-weighted hash of
every 8 KiB record): equal hashes for equal id sets = bit-identical sdf, weight and colour planes.
"""
import numpy as np
import pytest

from oracle import OracleMap
from texturefusion_b200 import capi, synth
from texturefusion_b200.maphash import sorted_chunk_hashes

pytestmark = pytest.mark.gpu


def assert_same_maps(g, o, what):
    gi, gh = sorted_chunk_hashes(g)
    oi, oh = sorted_chunk_hashes(o)
    assert gi.shape == oi.shape and np.array_equal(gi, oi), f"{what}: allocated chunk sets differ ({len(gi)} vs {len(oi)})"
    bad = np.nonzero(gh != oh)[0]
    assert len(bad) == 0, f"{what}: {len(bad)} of {len(gi)} chunks differ, first {gi[bad[0]]}"
    return len(gi)


def test_configs1_full_300_frame_sequence_at_5mm():
    cam = synth.Camera()
    seq = synth.make_sequence(300, cam=cam, total=300, keyframe_every=10, device="cuda")
    g = capi.Map(0.005, max_frames=8, max_chunks=1 << 19)
    o = OracleMap(0.005, threads=0)
    for fr in seq.frames:
        rg = fr.rgba() if fr.is_keyframe else None
        q = fr.quality if fr.is_keyframe else None
        g.upload_frame(fr.index, fr.depth, rg, q)
        st, ids, new, upd, qs = g.integrate_frame(fr.index, fr.is_keyframe, fr.pose, cam)
        n, n_upd = o.integrate_frame(fr.depth, rg, q, fr.pose, cam, fr.index if fr.is_keyframe else -1)
        assert (st.n_chunks, st.n_updated) == (n, n_upd), f"frame {fr.index}"
        assert int(upd.sum()) == n_upd and len(ids) == n
        if fr.index % 50 == 49:
            assert g.chunk_count() == o.chunk_count(), f"after frame {fr.index}"
    n = assert_same_maps(g, o, "300 frames at 5 mm")
    assert n > 5000
    g.close()


def _oracle_group(o, cam, group, old, flag, ids=None):
    """One ReIntegrateKeyframe call on the oracle (GCFusion/MobileFusion.cpp:114-221)."""
    kf = group[0]
    pose = lambda fr: fr.pose_old if old else fr.pose  # noqa: E731
    if flag:
        ids, new = o.prepare(kf.depth, pose(kf), cam)
        nu = np.zeros(len(ids), np.uint8)
    else:
        new = np.zeros(len(ids), np.uint8)
        nu = np.ones(len(ids), np.uint8)
    nu, _ = o.integrate(kf.depth, kf.rgba(), kf.quality, pose(kf), cam, ids, flag, kf.index, nu)
    for lf in group[1:]:
        nu, _ = o.integrate(lf.depth, None, None, pose(lf), cam, ids, flag, -1, nu)
    return o.finalize(ids, nu, new)


def run_loop_closure(res, keyframes, group, verify=True, device="cuda"):
    """Shared with tools/bench_loopclosure.py --verify.  Returns (gpu map, oracle map)."""
    cam = synth.Camera()
    n = keyframes * group
    seq = synth.make_sequence(n, cam=cam, total=max(n, 300), keyframe_every=group, device=device, with_drift=True)
    groups = [seq.frames[k:k + group] for k in range(0, n, group)]
    g = capi.Map(res, max_frames=n + 4, max_chunks=1 << 19)
    for fr in seq.frames:
        g.upload_frame(fr.index, fr.depth, fr.rgba() if fr.is_keyframe else None, fr.quality if fr.is_keyframe else None)

    def item(grp, flag, old, ids=None):
        d = {"flag": flag, "frames": [(fr.index, k == 0, fr.pose_old if old else fr.pose) for k, fr in enumerate(grp)]}
        if ids is not None:
            d["ids"] = ids
        return d

    first = g.integrate_batch([item(grp, 1, True) for grp in groups], cam)
    valid = [r[0] for r in first]
    items = []
    for grp, vl in zip(groups, valid):
        items += [item(grp, 0, True, vl), item(grp, 1, False)]
    second = g.integrate_batch(items, cam)
    o = OracleMap(res, threads=0)
    ovalid = [_oracle_group(o, cam, grp, True, 1) for grp in groups]
    for k, (a, b) in enumerate(zip(valid, ovalid)):
        assert np.array_equal(a, b), f"first fusion: validChunks of key-frame {k}"
    for k, grp in enumerate(groups):
        _oracle_group(o, cam, grp, True, 0, ovalid[k])
        v2 = _oracle_group(o, cam, grp, False, 1)
        assert np.array_equal(second[2 * k + 1][0], v2), f"re-integration: validChunks of key-frame {k}"
    return g, o


def test_configs2_loop_closure_50_keyframes_x7_batch():
    g, o = run_loop_closure(0.005, 50, 7)
    assert_same_maps(g, o, "50 key-frames x 7 de-/re-integrated")
    g.close()


def test_configs3_building_walk_slice_large_pool():
    cam = synth.Camera()
    g = capi.Map(0.005, max_frames=8, max_chunks=(1 << 19) + (1 << 18))
    o = OracleMap(0.005, threads=0)
    frames, total = 500, 5000
    removed = 0
    for k in range(frames):
        pose = synth.walk_pose(k, total)
        kf = k % 10 == 0
        depth, rgb, q = synth.render(pose, cam, color=kf, device="cuda", scene="building")
        fr = synth.Frame(k, pose, depth, rgb, np.ones(depth.shape, np.uint8) if kf else None, q, kf)
        rg = fr.rgba() if kf else None
        g.upload_frame(k, fr.depth, rg, fr.quality if kf else None)
        st, *_ = g.integrate_frame(k, kf, pose, cam, want_lists=False)
        n, n_upd = o.integrate_frame(fr.depth, rg, fr.quality if kf else None, pose, cam, k if kf else -1)
        assert (st.n_chunks, st.n_updated) == (n, n_upd), f"frame {k}"
        removed += st.n_removed
    assert removed > 100000, "the slice is meant to churn the allocator (create + garbage-collect)"
    n = assert_same_maps(g, o, "500 frames of the building walk")
    assert n > 100000
    c = g.counters()
    assert c["pool_capacity"] > (1 << 19)
    g.close()
