"""GPU parity: libtexfusion_b200.so (through the C ABI) against the CPU oracle on the same
seeded synthetic inputs.  Bar: allocated-chunk set and ordering bit-exact, weights and
colours bit-exact, TSDF within 1e-5 of the truncation distance (in practice bit-exact)."""
import numpy as np
import pytest

from oracle import OracleMap
from texturefusion_b200 import capi
from texturefusion_b200.chisel import Chisel

from util import RESOLUTIONS, assert_maps_equal, room_sequence, sort_ids

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("res", RESOLUTIONS)
def test_prepare_matches_oracle_order_and_flags(res):
    seq = room_sequence(3)
    cam = seq.cam
    g = capi.Map(res)
    o = OracleMap(res)
    for fr in seq.frames:
        g.upload_frame(fr.index, fr.depth)
        gi, gnew = g.prepare(fr.index, fr.pose, cam)
        oi, onew = o.prepare(fr.depth, fr.pose, cam)
        assert len(gi) == len(oi) and len(oi) > 0
        assert np.array_equal(gi, oi), "chunk list (reference traversal order) differs"
        assert np.array_equal(gnew, onew)
        assert g.chunk_count() == o.chunk_count()
    assert_maps_equal(g, o, what=f"prepare res={res}")


@pytest.mark.parametrize("res", RESOLUTIONS)
def test_fused_sequence_bit_exact(res):
    seq = room_sequence(6)
    cam = seq.cam
    g = capi.Map(res)
    o = OracleMap(res)
    for fr in seq.frames:
        rgba = fr.rgba() if fr.is_keyframe else None
        g.upload_frame(fr.index, fr.depth, rgba, fr.quality if fr.is_keyframe else None)
        st, ids, new, upd, q = g.integrate_frame(fr.index, fr.is_keyframe, fr.pose, cam)
        n, nupd = o.integrate_frame(fr.depth, rgba, fr.quality, fr.pose, cam, -1)
        assert st.n_chunks == n and st.n_updated == nupd
        assert g.chunk_count() == o.chunk_count()
    exact = assert_maps_equal(g, o, what=f"fused res={res}")
    assert exact, "TSDF within tolerance but not bit-exact"


@pytest.mark.parametrize("res", (0.02, 0.005))
def test_split_protocol_keyframe_with_local_frames(res):
    """ReIntegrateKeyframe(flag=1) protocol (GCFusion/MobileFusion.cpp:114-221): Prepare on the
    key-frame, integrate key-frame with colour + quality, local frames depth-only into the
    same list, Finalize."""
    seq = room_sequence(6)
    cam = seq.cam
    c = Chisel(voxelResolution=res)
    o = OracleMap(res)
    kf, local = seq.frames[0], seq.frames[1:3]
    ids, nu, new = c.PrepareIntersectChunks(kf.depth, kf.pose, cam)
    oi, onew = o.prepare(kf.depth, kf.pose, cam)
    assert np.array_equal(ids, oi) and np.array_equal(new, onew)
    c.IntegrateDepthScanColor(kf.depth, kf.rgba(), kf.pose, cam, ids, nu, 1, kf.index, kf.quality)
    onu, oq = o.integrate(kf.depth, kf.rgba(), kf.quality, kf.pose, cam, oi, 1, kf.index)
    assert np.array_equal(nu, onu)
    for lf in local:
        c.IntegrateDepthScanColor(lf.depth, None, lf.pose, cam, ids, nu, 1)
        onu, _ = o.integrate(lf.depth, None, None, lf.pose, cam, oi, 1, -1, onu)
        assert np.array_equal(nu, onu)
    valid = c.FinalizeIntegrateChunks(ids, nu, new)
    ovalid = o.finalize(oi, onu, onew)
    assert np.array_equal(valid, ovalid)
    assert assert_maps_equal(c.map, o, what="split protocol")
    # observations[keyframe] (Structure/Chisel.h:244-247), consumed by TexMap::update_datacost
    n_obs = 0
    for cid in valid[:: max(1, len(valid) // 200)]:
        want = o.observation(cid, kf.index)
        got = c.chunkManager.GetChunk(cid).observations.get(kf.index)
        assert (want is None) == (got is None)
        if want is not None:
            assert np.float32(want) == np.float32(got)
            n_obs += 1
    assert n_obs > 0
    # meshesToUpdate set
    om = {tuple(int(v) for v in r) for r in o.meshes_to_update()}
    assert om == {k for k, v in c.meshesToUpdate.items() if v}


@pytest.mark.parametrize("res", (0.02, 0.005))
def test_deintegrate_then_reintegrate(res):
    """Loop-closure step (GCFusion/MobileFusion.cpp:301-310): de-integrate a key-frame with its
    old pose over kf.validChunks, re-integrate under the corrected pose."""
    seq = room_sequence(4)
    cam = seq.cam
    c = Chisel(voxelResolution=res)
    o = OracleMap(res)
    valid_lists = []
    for fr in seq.frames[:3]:
        rgba = fr.rgba()
        ids, nu, new = c.PrepareIntersectChunks(fr.depth, fr.pose, cam)
        c.IntegrateDepthScanColor(fr.depth, rgba, fr.pose, cam, ids, nu, 1, fr.index, fr.quality)
        valid = c.FinalizeIntegrateChunks(ids, nu, new)
        oi, onew = o.prepare(fr.depth, fr.pose, cam)
        onu, _ = o.integrate(fr.depth, rgba, fr.quality, fr.pose, cam, oi, 1, fr.index)
        ovalid = o.finalize(oi, onu, onew)
        assert np.array_equal(valid, ovalid)
        valid_lists.append(valid)
    assert assert_maps_equal(c.map, o, what="before de-integration")
    fr = seq.frames[1]
    rgba = fr.rgba()
    vl = valid_lists[1]
    nu = np.ones(len(vl), np.uint8)
    c.IntegrateDepthScanColor(fr.depth, rgba, fr.pose, cam, vl, nu, 0, fr.index, fr.quality)
    c.FinalizeIntegrateChunks(vl, nu, np.zeros(len(vl), np.uint8))
    onu, _ = o.integrate(fr.depth, rgba, fr.quality, fr.pose, cam, vl, 0, fr.index, np.ones(len(vl), np.uint8))
    o.finalize(vl, onu, np.zeros(len(vl), np.uint8))
    assert assert_maps_equal(c.map, o, what="after de-integration")
    new_pose = fr.pose.copy()
    new_pose[:3, 3] += np.array([0.004, -0.003, 0.002], np.float32)
    ids, nu, new = c.PrepareIntersectChunks(fr.depth, new_pose, cam)
    c.IntegrateDepthScanColor(fr.depth, rgba, new_pose, cam, ids, nu, 1, fr.index, fr.quality)
    valid = c.FinalizeIntegrateChunks(ids, nu, new)
    oi, onew = o.prepare(fr.depth, new_pose, cam)
    onu, _ = o.integrate(fr.depth, rgba, fr.quality, new_pose, cam, oi, 1, fr.index)
    ovalid = o.finalize(oi, onu, onew)
    assert np.array_equal(valid, ovalid)
    assert assert_maps_equal(c.map, o, what="after re-integration")


def test_group_equals_sequential_calls():
    """tf_integrate_group (voxels held in registers across frames) == separate tf_integrate calls."""
    res = 0.005
    seq = room_sequence(6)
    cam = seq.cam
    a, b = capi.Map(res), capi.Map(res)
    kf = seq.frames[0]
    for m in (a, b):
        for fr in seq.frames[:4]:
            m.upload_frame(fr.index, fr.depth, fr.rgba() if fr.is_keyframe else None, fr.quality)
    ids, _ = a.prepare(kf.index, kf.pose, cam)
    ids_b, _ = b.prepare(kf.index, kf.pose, cam)
    assert np.array_equal(ids, ids_b)
    frames = [(kf.index, True, 1, kf.pose)] + [(fr.index, False, 1, fr.pose) for fr in seq.frames[1:4]]
    nu_a, q_a = a.integrate_group(frames, cam, ids)
    nu_b = np.zeros(len(ids), np.uint8)
    q_b = None
    for (fi, uc, fl, pose) in frames:
        nu_b, q = b.integrate(fi, uc, pose, cam, ids, fl, nu_b)
        q_b = q if q_b is None else q_b
    assert np.array_equal(nu_a, nu_b)
    assert np.array_equal(q_a.view(np.uint32), q_b.view(np.uint32))
    sa, wa, ca = a.download_chunks(ids)
    sb, wb, cb = b.download_chunks(ids)
    assert np.array_equal(sa.view(np.uint32), sb.view(np.uint32))
    assert np.array_equal(wa.view(np.uint32), wb.view(np.uint32))
    assert np.array_equal(ca, cb)


def test_errors_and_edge_cases():
    res = 0.02
    seq = room_sequence(2)
    cam = seq.cam
    g = capi.Map(res)
    fr = seq.frames[0]
    with pytest.raises(capi.TexFusionError) as e:
        g.prepare(123, fr.pose, cam)  # unknown frame
    assert e.value.code == capi.TF_ERR_NOT_FOUND
    g.upload_frame(fr.index, fr.depth)
    # unknown chunk in an integrate list -> NOT_FOUND (reference: unordered_map::at throws)
    with pytest.raises(capi.TexFusionError) as e:
        g.integrate(fr.index, False, fr.pose, cam, np.array([[1000, 1000, 1000]], np.int32), 1)
    assert e.value.code == capi.TF_ERR_NOT_FOUND
    # empty list is a no-op (Structure/Chisel.h:228)
    nu, q = g.integrate(fr.index, False, fr.pose, cam, np.zeros((0, 3), np.int32), 1)
    assert len(nu) == 0
    # all-invalid depth: bbox collapses to the 0.2 m shell around the camera, nothing is hit
    g.upload_frame(7, np.zeros_like(fr.depth))
    ids, new = g.prepare(7, fr.pose, cam)
    o = OracleMap(res)
    oi, _ = o.prepare(np.zeros_like(fr.depth), fr.pose, cam)
    assert np.array_equal(ids, oi)
    # colour requested but the frame has no colour plane
    with pytest.raises(capi.TexFusionError):
        g.integrate_frame(fr.index, True, fr.pose, cam)
    # remove + has_chunk
    before = g.chunk_count()
    ids, new = g.prepare(fr.index, fr.pose, cam)
    total = g.chunk_count()
    assert g.has_chunk(ids[0]) and total == before + int(new.sum())
    g.remove_chunks(ids[:5])
    assert not g.has_chunk(ids[0]) and g.chunk_count() == total - 5
    g.remove_chunks(ids[:5])  # removing again is a no-op (RemoveChunk returns false)
    assert g.chunk_count() == total - 5
    g.reset()
    assert g.chunk_count() == 0
    # a NaN or an infinity in the depth image is reported (both pipelines), and the map stays usable
    for k, bad_value in enumerate((np.nan, np.inf, -np.inf)):
        bad = fr.depth.copy()
        bad[17 + k, 33] = bad_value
        g.upload_frame(20 + k, bad)
        with pytest.raises(capi.TexFusionError) as e:
            g.integrate_frame(20 + k, False, fr.pose, cam) if k != 1 else g.prepare(20 + k, fr.pose, cam)
        assert e.value.code == capi.TF_ERR_INVALID and "NaN" in str(e.value)
    g.reset()
    g.upload_frame(fr.index, fr.depth)
    st, *_ = g.integrate_frame(fr.index, False, fr.pose, cam)
    o2 = OracleMap(res)
    n, nupd = o2.integrate_frame(fr.depth, None, None, fr.pose, cam, -1)
    assert (st.n_chunks, st.n_updated) == (n, nupd)


def test_fast_projection_never_disagrees_with_exact_path():
    """integrate_kernel evaluates the reference's (c/cz)*f + ch division-free (tf_device.cuh: project_safe
    = div.rn's own fast-path sequence without its range check) for chunks whose voxel centres pass a range
    test; inside that range it must equal the three IEEE ops bit for bit.  Random operands over the
    working range, operands engineered to land next to the rounding boundaries k + 0.5, quotients next
    to representable values' midpoints, the edges of the accepted range, tiny numerators; outside the
    range (zeros, denormals, huge, inf, NaN) the hook only has to report `not accepted`."""
    g = capi.Map(0.02)
    rng = np.random.RandomState(5)
    n = 1 << 22
    f, ch = np.float32(525.0), np.float32(319.5)
    cz = rng.uniform(0.05, 6.0, n).astype(np.float32)
    c = (rng.uniform(-1.2, 1.2, n) * cz).astype(np.float32)
    # engineered: choose c so that (c/cz)*f + ch is within a few ulps of k + 0.5
    k = rng.randint(-50, 700, n // 2)
    target = (k + 0.5 + rng.uniform(-2e-3, 2e-3, n // 2)).astype(np.float64)
    c[: n // 2] = ((target - float(ch)) / float(f) * cz[: n // 2].astype(np.float64)).astype(np.float32)
    # the whole accepted range, log-uniform, both signs of the numerator
    m = 1 << 21
    cz_w = np.exp2(rng.uniform(-17, 21, m)).astype(np.float32)
    c_w = (np.exp2(rng.uniform(-140, 21, m)) * rng.choice([-1.0, 1.0], m)).astype(np.float32)
    # divisors with all-ones / all-zeros mantissas and their neighbours (the hard cases of the Newton step)
    e = rng.randint(-16, 20, m).astype(np.float64)
    mant = rng.choice([1.0, 1.0 + 2.0 ** -23, 2.0 - 2.0 ** -23, 2.0 - 2.0 ** -22, 1.5, 1.0 + 2.0 ** -22], m)
    cz_m = (mant * np.exp2(e)).astype(np.float32)
    c_m = (rng.uniform(-2.0, 2.0, m) * cz_m).astype(np.float32)
    special = np.array([0.0, -0.0, 1e-30, -1e-30, 1e-40, 1e30, -1e30, np.inf, -np.inf, np.nan, 3e38, 1.0,
                        2.0 ** -17, 2.0 ** 21, 2.0 ** -103, 2.0 ** -126, 1e-45], np.float32)
    cs, zs = np.meshgrid(special, special)
    c = np.concatenate([c, c_w, c_m, cs.ravel()])
    cz = np.concatenate([cz, cz_w, cz_m, zs.ravel()])
    tot_acc = 0
    for (ff, cc) in ((f, ch), (np.float32(525.0), np.float32(239.5)), (np.float32(131.0), np.float32(79.5)),
                     (np.float32(2100.0), np.float32(1023.5)), (np.float32(1.0), np.float32(0.5))):
        uf, ue, acc = g.debug_project(c, cz, ff, cc)
        ue_sat = ue.copy()
        # cvtps_epi32 returns 0x80000000 where __float2int_rn saturates: both are "off the image" for the kernel
        off = (ue == np.int32(-2 ** 31)) & ((uf == np.int32(2 ** 31 - 1)) | (uf == np.int32(-2 ** 31)))
        bad = (acc != 0) & (uf != ue_sat) & ~off
        assert not bad.any(), f"{bad.sum()} in-range results differ, e.g. c={c[bad][:3]} cz={cz[bad][:3]} {uf[bad][:3]} {ue[bad][:3]}"
        tot_acc += acc[: n].mean()
    assert tot_acc / 5 == 1.0  # the working range is inside the accepted range


def test_inline_division_of_the_running_average_is_the_ieee_quotient():
    """Phase B of integrate_kernel forms (s w + sd w') / (w + w' + 1e-4) with div.rn's fast-path sequence
    inline and redoes it with the IEEE division when the result is not a number of at least 2^-60 in
    magnitude.  Whatever the operands, the value it keeps must be the IEEE quotient (bit for bit,
    signed zeros included) for every divisor the kernel can pass (> 0.5)."""
    g = capi.Map(0.02)
    rng = np.random.RandomState(11)
    n = 1 << 22
    den = np.concatenate([rng.uniform(0.5001, 400.0, n // 2), np.exp2(rng.uniform(-1, 127.9, n // 2))]).astype(np.float32)
    num = np.concatenate([rng.uniform(-0.2, 0.2, n // 4) * den[: n // 4], np.exp2(rng.uniform(-149, 127, n - n // 4)) *
                          rng.choice([-1.0, 1.0], n - n // 4)]).astype(np.float32)
    e = rng.randint(0, 30, n).astype(np.float64)
    mant = rng.choice([1.0, 1.0 + 2.0 ** -23, 2.0 - 2.0 ** -23, 2.0 - 2.0 ** -22, 1.5, 1.0 + 2.0 ** -22], n)
    den_m = (mant * np.exp2(e)).astype(np.float32)
    num_m = (rng.uniform(-3.0, 3.0, n) * den_m).astype(np.float32)
    special_n = np.array([0.0, -0.0, 1e-45, -1e-45, 1e-38, 2.0 ** -103, -2.0 ** -104, 2.0 ** -61, 2.0 ** -60, 1e-30, 1e30, -1e30,
                          3e38, np.inf, -np.inf, np.nan, 999.0, 1.0], np.float32)
    special_d = np.array([0.50001, 0.5001, 1.0, 1.0001, 3.0, 2.0 ** 60, 2.0 ** 125, 2.0 ** 126, 2.0 ** 127, 3e38, np.inf], np.float32)
    ns, ds = np.meshgrid(special_n, special_d)
    num = np.concatenate([num, num_m, ns.ravel()])
    den = np.concatenate([den, den_m, ds.ravel()])
    qk, qi, acc = g.debug_divide(num, den)
    same = (qk.view(np.uint32) == qi.view(np.uint32)) | (np.isnan(qk) & np.isnan(qi))
    assert same.all(), f"{(~same).sum()} quotients differ, e.g. num={num[~same][:4]} den={den[~same][:4]} {qk[~same][:4]} {qi[~same][:4]}"
    assert acc[: n // 4].mean() > 0.99  # the inline result is the one normally kept


def _oracle_keyframe_group(o, cam, group, poses, flag, ids=None):
    """ReIntegrateKeyframe on the oracle: group[0] is the key-frame (colour + quality), the rest
    are its local depth frames.  flag 1: prepare/integrate/finalize; flag 0: de-integrate over ids."""
    kf = group[0]
    if flag == 1:
        ids, new = o.prepare(kf.depth, poses[0], cam)
        nu = np.zeros(len(ids), np.uint8)
    else:
        new = np.zeros(len(ids), np.uint8)
        nu = np.ones(len(ids), np.uint8)
    nu, q = o.integrate(kf.depth, kf.rgba(), kf.quality, poses[0], cam, ids, flag, kf.index, nu)
    for lf, p in zip(group[1:], poses[1:]):
        nu, _ = o.integrate(lf.depth, None, None, p, cam, ids, flag, -1, nu)
    valid = o.finalize(ids, nu, new)
    keep = nu != 0
    return valid, q[keep]


@pytest.mark.parametrize("res", (0.02, 0.005))
def test_batch_loop_closure_reintegration(res):
    """tf_integrate_batch: key-frame groups fused under drifted poses, then every key-frame is
    de-integrated (old poses, its validChunks) and re-integrated (corrected poses) — the loop of
    GCFusion/MobileFusion.cpp:301-310, batched."""
    from texturefusion_b200 import synth
    cam = synth.Camera()
    seq = synth.make_sequence(9, cam=cam, total=300, keyframe_every=3, with_drift=True, start=60)
    groups = [seq.frames[k:k + 3] for k in range(0, 9, 3)]
    g = capi.Map(res, max_frames=16)
    o = OracleMap(res)
    for fr in seq.frames:
        g.upload_frame(fr.index, fr.depth, fr.rgba() if fr.is_keyframe else None, fr.quality if fr.is_keyframe else None)

    def item(group, flag, old, ids=None):
        d = {"flag": flag, "frames": [(fr.index, k == 0, fr.pose_old if old else fr.pose) for k, fr in enumerate(group)]}
        if ids is not None:
            d["ids"] = ids
        return d

    # 1. first fusion under the drifted poses
    res1 = g.integrate_batch([item(gr, 1, True) for gr in groups], cam)
    valid_lists = []
    for gr, (gv, gq) in zip(groups, res1):
        ov, oq = _oracle_keyframe_group(o, cam, gr, [fr.pose_old for fr in gr], 1)
        assert np.array_equal(gv, ov)
        assert np.array_equal(gq.view(np.uint32), oq.view(np.uint32))
        valid_lists.append(ov)
    assert assert_maps_equal(g, o, what="batch: first fusion")
    # 2. loop closure: de-integrate + re-integrate every key-frame in one batch call
    items = []
    for gr, vl in zip(groups, valid_lists):
        items += [item(gr, 0, True, vl), item(gr, 1, False)]
    res2 = g.integrate_batch(items, cam)
    for k, (gr, vl) in enumerate(zip(groups, valid_lists)):
        _oracle_keyframe_group(o, cam, gr, [fr.pose_old for fr in gr], 0, vl)
        ov, oq = _oracle_keyframe_group(o, cam, gr, [fr.pose for fr in gr], 1)
        gv, gq = res2[2 * k + 1]
        assert np.array_equal(gv, ov)
        assert np.array_equal(gq.view(np.uint32), oq.view(np.uint32))
    assert assert_maps_equal(g, o, what="batch: after loop closure")


@pytest.mark.parametrize("res,n_ranks", ((0.02, 2), (0.005, 4)))
def test_sharded_maps_union_equals_single_map(res, n_ranks):
    """Chunk sharding (one tf_map per rank): every rank culls the same broadcast frame and keeps the
    chunks it owns; the union of the per-rank maps is the single-GPU map, and the merged
    per-rank lists restore the reference's traversal order."""
    from texturefusion_b200 import sharding
    seq = room_sequence(4)
    cam = seq.cam
    single = capi.Map(res)
    ranks = [capi.Map(res, n_ranks=n_ranks, rank=r, max_chunks=1 << 16) for r in range(n_ranks)]
    o = OracleMap(res)
    for fr in seq.frames:
        rgba = fr.rgba() if fr.is_keyframe else None
        q = fr.quality if fr.is_keyframe else None
        single.upload_frame(fr.index, fr.depth, rgba, q)
        st, ids, new, upd, qs = single.integrate_frame(fr.index, fr.is_keyframe, fr.pose, cam)
        lo, _ = o.boundary_ids(fr.depth, fr.pose, cam)
        parts, pay = [], []
        for r, m in enumerate(ranks):
            m.upload_frame(fr.index, fr.depth, rgba, q)
            st_r, ids_r, new_r, upd_r, q_r = m.integrate_frame(fr.index, fr.is_keyframe, fr.pose, cam)
            assert np.all(sharding.owner_of(ids_r, n_ranks) == r)
            parts.append(ids_r)
            pay.append((new_r, upd_r, q_r))
        step = 1 if np.float32(res) > 0.01 else 4
        m_ids, m_new, m_upd, m_q = sharding.merge_rank_lists(parts, pay, min_id=lo, step=step)
        assert np.array_equal(m_ids, ids) and np.array_equal(m_new, new) and np.array_equal(m_upd, upd)
        assert np.array_equal(m_q.view(np.uint32), qs.view(np.uint32))
    all_ids = np.concatenate([m.list_chunks() for m in ranks])
    assert len(all_ids) == single.chunk_count() and sum(m.chunk_count() for m in ranks) == single.chunk_count()
    si, _ = sort_ids(single.list_chunks())
    ai, order = sort_ids(all_ids)
    assert np.array_equal(si, ai)
    ss, sw, sc = single.download_chunks(si)
    for r, m in enumerate(ranks):
        mine = si[sharding.owner_of(si, n_ranks) == r]
        rs, rw, rc = m.download_chunks(mine)
        sel = sharding.owner_of(si, n_ranks) == r
        assert np.array_equal(rs.view(np.uint32), ss[sel].view(np.uint32))
        assert np.array_equal(rw.view(np.uint32), sw[sel].view(np.uint32))
        assert np.array_equal(rc, sc[sel])


def test_capacity_errors_and_frame_store_eviction():
    """Pool exhaustion is reported (TF_ERR_CAPACITY), not silently dropped; the frame store
    replaces its least-recently-used slot."""
    seq = room_sequence(4)
    cam = seq.cam
    fr = seq.frames[0]
    tiny = capi.Map(0.005, max_chunks=1024, max_frames=2)
    tiny.upload_frame(fr.index, fr.depth)
    with pytest.raises(capi.TexFusionError) as e:
        tiny.integrate_frame(fr.index, False, fr.pose, cam)
    assert e.value.code == capi.TF_ERR_CAPACITY and "pool" in str(e.value)
    # LRU: with two slots, uploading a third frame evicts the oldest one
    g = capi.Map(0.02, max_frames=2)
    for f in seq.frames[:3]:
        g.upload_frame(f.index, f.depth)
    with pytest.raises(capi.TexFusionError) as e:
        g.prepare(seq.frames[0].index, seq.frames[0].pose, cam)
    assert e.value.code == capi.TF_ERR_NOT_FOUND
    ids, _ = g.prepare(seq.frames[2].index, seq.frames[2].pose, cam)
    assert len(ids) > 0
    g.release_frame(seq.frames[2].index)
    with pytest.raises(capi.TexFusionError):
        g.prepare(seq.frames[2].index, seq.frames[2].pose, cam)
    # output capacity: the two-call pattern of tf_prepare
    g.upload_frame(5, fr.depth)
    with pytest.raises(capi.TexFusionError) as e:
        g.prepare(5, fr.pose, cam, cap=10)
    assert e.value.code == capi.TF_ERR_CAPACITY and g.chunk_count() == len(ids)  # nothing was created
    # wrong image size
    from texturefusion_b200 import synth
    with pytest.raises(capi.TexFusionError):
        g.prepare(5, fr.pose, synth.Camera().scaled(0.5))


def test_long_sequence_chunk_recycling():
    """60 frames at 5 mm: thousands of chunks are created and garbage-collected every frame, so
    slots and hash tombstones are recycled many times; the map must still equal the oracle's."""
    seq = room_sequence(60, keyframe_every=10)
    cam = seq.cam
    g = capi.Map(0.005, max_chunks=1 << 17)
    o = OracleMap(0.005, threads=0)
    removed = 0
    for fr in seq.frames:
        rgba = fr.rgba() if fr.is_keyframe else None
        g.upload_frame(fr.index, fr.depth, rgba, fr.quality if fr.is_keyframe else None)
        st, *_ = g.integrate_frame(fr.index, fr.is_keyframe, fr.pose, cam, want_lists=False)
        n, nupd = o.integrate_frame(fr.depth, rgba, fr.quality, fr.pose, cam, -1)
        assert (st.n_chunks, st.n_updated) == (n, nupd)
        removed += st.n_removed
    assert removed > 50000
    assert assert_maps_equal(g, o, what="60-frame sequence")
    c = g.counters()
    assert c["pool_used"] == g.chunk_count() and c["frames_integrated"] == 60


def test_full_size_round_trip_properties():
    """Size-independent properties at the headline configuration (640x480, 5 mm):
    integrate followed by de-integrate of the same frame under the same pose returns every
    touched voxel to (999, 0) and every colour voxel to 0 (u16 wrap-around arithmetic);
    integrating twice doubles the weights exactly; lists are duplicate free and owned once."""
    seq = room_sequence(2)
    cam = seq.cam
    kf = seq.frames[0]
    g = capi.Map(0.005)
    g.upload_frame(kf.index, kf.depth, kf.rgba(), kf.quality)
    ids, new = g.prepare(kf.index, kf.pose, cam)
    assert len(np.unique(ids, axis=0)) == len(ids) and new.all()
    nu, q = g.integrate(kf.index, True, kf.pose, cam, ids, 1)
    s1, w1, c1 = g.download_chunks(ids)
    assert (w1 > 0).sum() > 1_000_000 and (c1.reshape(-1, 4)[:, 3] > 0).sum() > 100_000
    # twice the same observation: weights double exactly, sdf stays within rounding of itself
    g.integrate(kf.index, True, kf.pose, cam, ids, 1)
    s2, w2, c2 = g.download_chunks(ids)
    seen = w1 > 0
    assert np.array_equal(w2[seen], (w1[seen] + w1[seen]).astype(np.float32))
    assert np.allclose(s2[seen], s1[seen], rtol=0, atol=2e-5)
    assert np.array_equal(c2.reshape(-1, 4)[:, 3], 2 * c1.reshape(-1, 4)[:, 3])
    # remove both observations again
    for _ in range(2):
        g.integrate(kf.index, True, kf.pose, cam, ids, 0)
    s0, w0, c0 = g.download_chunks(ids)
    assert np.all(w0 == 0) and np.all(s0[seen] == 999.0) and not c0.any()


@pytest.mark.parametrize("res", (0.02, 0.005))
def test_fused_frame_output_paths_agree(res):
    """The fused call must give the same map and the same ordered lists whichever way the results
    travel: CUDA graph + completion stamp or plain launches (profiling level 1 forces those),
    lists exported straight into page-locked caller buffers, through the staging buffers into
    pageable arrays, or not requested at all; the next frame's upload may be in flight (copy
    stream) while a frame is fused."""
    import ctypes as C
    seq = room_sequence(6)
    cam = seq.cam
    o = OracleMap(res)
    variants = {"graph+pageable": capi.Map(res), "plain+pageable": capi.Map(res), "graph+pinned": capi.Map(res),
                "graph, no lists": capi.Map(res)}
    variants["plain+pageable"].set_profiling(1)
    cap = 1 << 16
    pins = [capi.PinnedBuffer((cap, 3), np.int32), capi.PinnedBuffer((cap,), np.uint8), capi.PinnedBuffer((cap,), np.uint8),
            capi.PinnedBuffer((cap,), np.float32)]
    vp = C.c_void_p
    frames = seq.frames

    def upload(m, fr):
        m.upload_frame(fr.index, fr.depth, fr.rgba() if fr.is_keyframe else None, fr.quality if fr.is_keyframe else None)

    for m in variants.values():
        upload(m, frames[0])
    for i, fr in enumerate(frames):
        rgba = fr.rgba() if fr.is_keyframe else None
        o_ids, o_new = o.prepare(fr.depth, fr.pose, cam)
        o_upd, _ = o.integrate(fr.depth, rgba, fr.quality if fr.is_keyframe else None, fr.pose, cam, o_ids, 1, -1)
        o.finalize(o_ids, o_upd, o_new)
        for name, m in variants.items():
            if i + 1 < len(frames):
                upload(m, frames[i + 1])  # in flight during this frame's kernels
                if name == "graph+pinned":
                    assert m.L.tf_wait_upload(m.h, frames[i + 1].index) == 0
                    assert m.L.tf_wait_upload(m.h, 987654) == capi.TF_ERR_NOT_FOUND
            if name == "graph+pinned":
                st = capi.FrameStats()
                rc = m.L.tf_integrate_frame(m.h, fr.index, int(fr.is_keyframe), C.byref(capi.make_pose(fr.pose)),
                                            C.byref(capi.make_camera(cam)), C.byref(st), vp(pins[0].ptr), vp(pins[1].ptr),
                                            vp(pins[2].ptr), vp(pins[3].ptr), cap)
                assert rc == 0, m.L.tf_last_error(m.h)
                n = st.n_chunks
                ids, new, upd = pins[0].array[:n], pins[1].array[:n], pins[2].array[:n]
            elif name == "graph, no lists":
                st, *_ = m.integrate_frame(fr.index, fr.is_keyframe, fr.pose, cam, want_lists=False)
                assert st.n_chunks == len(o_ids), name
                continue
            else:
                st, ids, new, upd, q = m.integrate_frame(fr.index, fr.is_keyframe, fr.pose, cam)
            assert np.array_equal(ids, o_ids), f"{name}: frame {i} list order"
            assert np.array_equal(new != 0, np.asarray(o_new) != 0), f"{name}: frame {i} is_new"
            assert np.array_equal(upd != 0, np.asarray(o_upd) != 0), f"{name}: frame {i} updated"
    for name, m in variants.items():
        assert assert_maps_equal(m, o, what=name)
        m.close()


def test_batch_is_order_preserving_across_sub_batches_and_reports_missing_ids():
    """tf_integrate_batch queues its items without host synchronisation (one per 32
    re-integrations): a long batch must give exactly the map and the valid lists of the same items
    issued one call at a time, and a de-integration id that is not in the map is reported for the
    batch as a whole."""
    res = 0.02
    from texturefusion_b200 import synth
    cam = synth.Camera()
    seq = synth.make_sequence(4, cam=cam, total=300, keyframe_every=2, with_drift=True, start=30)
    groups = [seq.frames[0:2], seq.frames[2:4]]
    a, b = capi.Map(res, max_frames=8), capi.Map(res, max_frames=8)
    for m in (a, b):
        for fr in seq.frames:
            m.upload_frame(fr.index, fr.depth, fr.rgba() if fr.is_keyframe else None, fr.quality if fr.is_keyframe else None)

    def item(group, flag, old, ids=None):
        d = {"flag": flag, "frames": [(fr.index, k == 0, fr.pose_old if old else fr.pose) for k, fr in enumerate(group)]}
        if ids is not None:
            d["ids"] = ids
        return d

    first = [item(g, 1, True) for g in groups]
    valid_a = [r[0] for r in a.integrate_batch(first, cam)]
    valid_b = [b.integrate_batch([it], cam)[0][0] for it in first]
    old = True
    items, n_rounds = [], 18  # 18 x 2 key-frames = 36 re-integrations + 36 de-integrations: two sub-batches
    # the batch needs every de-integration list up front, so build it from the one-at-a-time map first
    seq_results = []
    for r in range(n_rounds):
        for k, g in enumerate(groups):
            b.integrate_batch([item(g, 0, old, valid_b[k])], cam)
            valid_b[k] = b.integrate_batch([item(g, 1, not old)], cam)[0][0]
            seq_results.append(valid_b[k].copy())
        old = not old
    old = True
    lists = [v.copy() for v in valid_a]
    k_res = 0
    for r in range(n_rounds):
        for k, g in enumerate(groups):
            items += [item(g, 0, old, lists[k]), item(g, 1, not old)]
            lists[k] = seq_results[k_res]  # what the re-integration will return (checked below)
            k_res += 1
        old = not old
    out = a.integrate_batch(items, cam)
    got = [o[0] for o in out if o is not None]
    assert len(got) == len(seq_results) == 36
    for x, y in zip(got, seq_results):
        assert np.array_equal(x, y)
    ia, ib = sort_ids(a.list_chunks())[0], sort_ids(b.list_chunks())[0]
    assert np.array_equal(ia, ib)
    sa, sb = a.download_chunks(ia), b.download_chunks(ib)
    for x, y in zip(sa, sb):
        assert np.array_equal(x.view(np.uint8), y.view(np.uint8))
    # unknown id in a de-integration list
    bad = np.vstack([lists[0][:3], np.array([[9999, 9999, 9999]], np.int32)])
    with pytest.raises(capi.TexFusionError) as e:
        a.integrate_batch([item(groups[0], 0, old, bad)], cam)
    assert e.value.code == capi.TF_ERR_NOT_FOUND
    a.close()
    b.close()


def test_batch_items_without_outputs_stream_a_sequence():
    """Single-frame re-integration items that ask for nothing back (no valid list) queue a frame
    sequence without host round trips; the map must equal the one built frame by frame."""
    res = 0.02
    seq = room_sequence(8)
    cam = seq.cam
    a, b = capi.Map(res, max_frames=16), capi.Map(res, max_frames=16)
    for m in (a, b):
        for fr in seq.frames:
            m.upload_frame(fr.index, fr.depth, fr.rgba() if fr.is_keyframe else None, fr.quality if fr.is_keyframe else None)
    items = [{"flag": 1, "frames": [(fr.index, fr.is_keyframe, fr.pose)]} for fr in seq.frames]
    out = a.run_batch(a.marshal_batch(items, want_lists=False), cam)
    assert all(o is None for o in out)
    for fr in seq.frames:
        b.integrate_frame(fr.index, fr.is_keyframe, fr.pose, cam, want_lists=False)
    ia, ib = sort_ids(a.list_chunks())[0], sort_ids(b.list_chunks())[0]
    assert np.array_equal(ia, ib) and len(ia) > 0
    for x, y in zip(a.download_chunks(ia), b.download_chunks(ib)):
        assert np.array_equal(x.view(np.uint8), y.view(np.uint8))
    assert a.counters()["voxel_updates"] == b.counters()["voxel_updates"]
    a.close()
    b.close()


@pytest.mark.parametrize("seed", (0, 1, 2))
def test_random_protocol_mix_matches_oracle(seed):
    """State-machine fuzz: one map driven by a random mix of the entry points — fused frames with and
    without lists, the split Prepare / Integrate / Finalize protocol, queued batches and
    de-/re-integration of earlier key-frames, with tf_reset in between — must stay bit-identical to
    the oracle driven by the same calls (allocator, free stack, lazy chunks, garbage collection and
    the frame store are shared by all of these paths)."""
    rng = np.random.RandomState(seed)
    res = (0.02, 0.01, 0.005)[seed]
    from texturefusion_b200 import synth
    cam = synth.Camera()
    seq = synth.make_sequence(12, cam=cam, total=300, keyframe_every=3, with_drift=True, start=90 + 21 * seed)
    groups = [seq.frames[k:k + 3] for k in range(0, 12, 3)]
    g = capi.Map(res, max_frames=16)
    o = OracleMap(res)
    for fr in seq.frames:
        g.upload_frame(fr.index, fr.depth, fr.rgba() if fr.is_keyframe else None, fr.quality if fr.is_keyframe else None)
    fused_in = {}  # key-frame group index -> (poses used, valid list) for groups currently in the map

    def item(group, flag, poses, ids=None):
        d = {"flag": flag, "frames": [(fr.index, k == 0, poses[k]) for k, fr in enumerate(group)]}
        if ids is not None:
            d["ids"] = ids
        return d

    for step in range(14):
        free = [k for k in range(len(groups)) if k not in fused_in]
        op = rng.choice(["fused", "fused_nolists", "split", "batch_in", "batch_out", "reset"],
                        p=[0.2, 0.15, 0.15, 0.25, 0.2, 0.05])
        if op in ("fused", "fused_nolists"):
            fr = seq.frames[rng.randint(len(seq.frames))]
            rgba = fr.rgba() if fr.is_keyframe else None
            st, ids, new, upd, q = g.integrate_frame(fr.index, fr.is_keyframe, fr.pose, cam, want_lists=(op == "fused"))
            oi, onew = o.prepare(fr.depth, fr.pose, cam)
            onu, _ = o.integrate(fr.depth, rgba, fr.quality if fr.is_keyframe else None, fr.pose, cam, oi, 1, -1)
            o.finalize(oi, onu, onew)
            assert st.n_chunks == len(oi) and st.n_updated == int(np.count_nonzero(onu))
            if op == "fused":
                assert np.array_equal(ids, oi) and np.array_equal(upd != 0, np.asarray(onu) != 0)
        elif op == "split":
            fr = seq.frames[rng.randint(len(seq.frames))]
            ids, new = g.prepare(fr.index, fr.pose, cam)
            oi, onew = o.prepare(fr.depth, fr.pose, cam)
            assert np.array_equal(ids, oi) and np.array_equal(new, onew)
            nu, _ = g.integrate(fr.index, False, fr.pose, cam, ids, 1)
            onu, _ = o.integrate(fr.depth, None, None, fr.pose, cam, oi, 1, -1)
            assert np.array_equal(nu != 0, np.asarray(onu) != 0)
            garbage = ids[(nu == 0) & (new != 0)]
            if len(garbage):
                g.remove_chunks(garbage)
            o.finalize(oi, onu, onew)
        elif op == "batch_in" and free:
            ks = [int(k) for k in rng.choice(free, size=min(len(free), 1 + rng.randint(2)), replace=False)]
            old = bool(rng.randint(2))
            poses = {k: [fr.pose_old if old else fr.pose for fr in groups[k]] for k in ks}
            out = g.integrate_batch([item(groups[k], 1, poses[k]) for k in ks], cam)
            for k, (gv, gq) in zip(ks, out):
                ov, oq = _oracle_keyframe_group(o, cam, groups[k], poses[k], 1)
                assert np.array_equal(gv, ov) and np.array_equal(gq.view(np.uint32), oq.view(np.uint32))
                fused_in[k] = (poses[k], ov)
        elif op == "batch_out" and fused_in:
            k = int(rng.choice(list(fused_in)))
            poses, vl = fused_in.pop(k)
            if len(vl):
                g.integrate_batch([item(groups[k], 0, poses, vl)], cam)
                _oracle_keyframe_group(o, cam, groups[k], poses, 0, vl)
        elif op == "reset":
            g.reset()
            o.reset()
            fused_in.clear()
        assert g.chunk_count() == o.chunk_count(), f"step {step} ({op})"
    assert assert_maps_equal(g, o, what=f"protocol mix seed {seed}")
    g.close()


@pytest.mark.parametrize("scale,res", ((0.35, 0.02), (0.1, 0.04), (1.7, 0.01)))
def test_other_image_sizes(scale, res):
    """Images other than 640x480 (224x168, 64x48, 1088x816; the width must stay a multiple of 8 like
    the reference's 8-pixel loops): plane strides in the frame-store slabs, grid sizes derived from
    the pixel count, and a principal point that changes the rounding slack of the projection."""
    seq = room_sequence(4, scale=scale)
    cam = seq.cam
    assert cam.width % 8 == 0
    with pytest.raises(capi.TexFusionError):
        capi.Map(res, width=cam.width + 3, height=cam.height)
    g = capi.Map(res, width=cam.width, height=cam.height, max_frames=8)
    o = OracleMap(res)
    for fr in seq.frames:
        rgba = fr.rgba() if fr.is_keyframe else None
        g.upload_frame(fr.index, fr.depth, rgba, fr.quality if fr.is_keyframe else None)
        st, ids, new, upd, q = g.integrate_frame(fr.index, fr.is_keyframe, fr.pose, cam)
        oi, onew = o.prepare(fr.depth, fr.pose, cam)
        onu, _ = o.integrate(fr.depth, rgba, fr.quality if fr.is_keyframe else None, fr.pose, cam, oi, 1, -1)
        o.finalize(oi, onu, onew)
        assert np.array_equal(ids, oi) and np.array_equal(upd != 0, np.asarray(onu) != 0)
    assert assert_maps_equal(g, o, what=f"image {cam.width}x{cam.height}")
    g.close()


def test_colour_count_overflow_shift():
    """ProjectionIntegrator.cpp:274-292: once a voxel's colour count exceeds 120 all four u16
    accumulators are shifted right by two.  130 integrations of the same key-frame drive every
    observed voxel through that branch (and 130 de-integrations through the wrapping subtract)."""
    res = 0.04
    seq = room_sequence(1, keyframe_every=1, scale=0.25)
    cam, fr = seq.cam, seq.frames[0]
    g = capi.Map(res, width=cam.width, height=cam.height, max_frames=4)
    o = OracleMap(res)
    g.upload_frame(fr.index, fr.depth, fr.rgba(), fr.quality)
    for _ in range(130):
        g.integrate_frame(fr.index, True, fr.pose, cam, want_lists=False)
        o.integrate_frame(fr.depth, fr.rgba(), fr.quality, fr.pose, cam, -1)
    ids = sort_ids(g.list_chunks())[0]
    _, _, col = g.download_chunks(ids)
    counts = col.reshape(len(ids), 512, 4)[:, :, 3]
    assert counts.max() <= 120 + 1 and (counts > 30).any(), "the shift branch was not reached"
    assert assert_maps_equal(g, o, what="130 colour integrations")
    for _ in range(3):  # wrapping de-integration of shifted accumulators, over every chunk of the map
        g.integrate(fr.index, True, fr.pose, cam, ids, 0)
        o.integrate(fr.depth, fr.rgba(), fr.quality, fr.pose, cam, ids, 0, -1)
    assert assert_maps_equal(g, o, what="de-integration after the shift")
    g.close()


@pytest.mark.parametrize("res", (0.02, 0.005))
def test_near_and_far_planes_inside_the_depth_range(res):
    """Depth values outside (near, far) of the camera handed to the fusion call: the culling gates
    (block origin depth, ChunkManager.h:598-602) and the TSDF gate (ProjectionIntegrator.cpp:310-317)
    must cut exactly where the reference cuts."""
    import dataclasses
    seq = room_sequence(3)
    cam = dataclasses.replace(seq.cam, near=0.9, far=2.2)
    g = capi.Map(res)
    o = OracleMap(res)
    for fr in seq.frames:
        rgba = fr.rgba() if fr.is_keyframe else None
        g.upload_frame(fr.index, fr.depth, rgba, fr.quality if fr.is_keyframe else None)
        st, ids, new, upd, q = g.integrate_frame(fr.index, fr.is_keyframe, fr.pose, cam)
        oi, onew = o.prepare(fr.depth, fr.pose, cam)
        onu, _ = o.integrate(fr.depth, rgba, fr.quality if fr.is_keyframe else None, fr.pose, cam, oi, 1, -1)
        o.finalize(oi, onu, onew)
        assert np.array_equal(ids, oi) and np.array_equal(upd != 0, np.asarray(onu) != 0)
        assert 0 < int(np.count_nonzero(onu)) < len(oi)
    assert assert_maps_equal(g, o, what="near/far inside the depth range")
    g.close()


@pytest.mark.parametrize("res", (0.02, 0.005))
def test_chunks_at_and_behind_the_camera_plane(res):
    """integrate_kernel projects division-free only for chunks whose 512 voxel centres lie in a safe depth
    range (origin depth above FrameDev::z_safe, about 13 voxels); every other chunk takes the reference's own
    three ops in a compact loop.  The culling never lists such chunks at fine resolutions, the list form
    (de-integration, ReIntegrateKeyframe over validChunks) can: a camera moved INTO the fused surface
    integrates over every chunk of the map — chunks behind it (negative depths), cut by its plane (zero and
    tiny depths, quotients of either sign and any size) and in front of it — single frames with and without
    colour and a three-frame group, against the oracle bit for bit."""
    seq = room_sequence(4, start=40)
    cam = seq.cam
    g = capi.Map(res, max_frames=16)
    o = OracleMap(res)
    first = seq.frames[0]
    g.upload_frame(first.index, first.depth, None, None)
    g.integrate_frame(first.index, False, first.pose, cam)
    o.integrate_frame(first.depth, None, None, first.pose, cam, -1)
    ids = g.list_chunks()
    assert len(ids) > 100
    # a camera on the first frame's central ray, a few voxels in front of / behind the surface it saw
    d_c = float(first.depth[cam.height // 2, cam.width // 2])
    yy, xx = np.mgrid[0:cam.height, 0:cam.width]
    n_slow = n_upd = 0
    for k, back in enumerate((6 * res, 20 * res, -4 * res)):
        pose = np.array(first.pose, np.float32).copy()
        pose[:3, 3] += pose[:3, 2] * np.float32(d_c - back)
        z = ((ids.astype(np.float64) * 8 * res - pose[:3, 3].astype(np.float64)) @ pose[:3, :3].astype(np.float64))[:, 2]
        n_slow += int(np.count_nonzero(z < 13 * res))
        depth = (abs(back) + 3 * res * (xx / cam.width) + 2 * res * (yy / cam.height)).astype(np.float32)
        color = k != 1
        rgba = seq.frames[2].rgba() if color else None
        quality = seq.frames[2].quality if color else None
        fi = 100 + k
        g.upload_frame(fi, depth, rgba, quality)
        nu, q = g.integrate(fi, color, pose, cam, ids, 1)
        onu, oq = o.integrate(depth, rgba, quality, pose, cam, ids, 1, fi if color else -1)
        assert np.array_equal(nu, onu), f"needsUpdate differs for camera offset {back}"
        assert np.array_equal(q.view(np.uint32)[nu != 0], np.asarray(oq, np.float32).view(np.uint32)[nu != 0])
        n_upd += int(np.count_nonzero(nu))
    assert n_slow > 0 and n_upd > 0
    # the same three frames as one group (the group kernel), de-integrated
    poses = []
    for k, back in enumerate((6 * res, 20 * res, -4 * res)):
        pose = np.array(first.pose, np.float32).copy()
        pose[:3, 3] += pose[:3, 2] * np.float32(d_c - back)
        poses.append(pose)
    nu, q = g.integrate_group([(100, True, 0, poses[0]), (101, False, 0, poses[1]), (102, False, 0, poses[2])], cam, ids)
    depths = [(abs(back) + 3 * res * (xx / cam.width) + 2 * res * (yy / cam.height)).astype(np.float32) for back in (6 * res, 20 * res, -4 * res)]
    onu, _ = o.integrate(depths[0], seq.frames[2].rgba(), seq.frames[2].quality, poses[0], cam, ids, 0, 100)
    onu, _ = o.integrate(depths[1], None, None, poses[1], cam, ids, 0, -1, onu)
    onu, _ = o.integrate(depths[2], None, None, poses[2], cam, ids, 0, -1, onu)
    assert np.array_equal(nu != 0, np.asarray(onu) != 0)
    assert assert_maps_equal(g, o, what="chunks at the camera plane")
    g.close()


def test_noisy_depth_sequence_bit_exact():
    """Gaussian depth noise (sigma 2 mm) makes pixel roundings, band edges and the early-exit rows
    irregular; fused frames and a three-frame group against the oracle at 5 mm."""
    res = 0.005
    seq = room_sequence(6, noise=0.002, start=123)
    cam = seq.cam
    g = capi.Map(res)
    o = OracleMap(res)
    for fr in seq.frames[:3]:
        rgba = fr.rgba() if fr.is_keyframe else None
        g.upload_frame(fr.index, fr.depth, rgba, fr.quality if fr.is_keyframe else None)
        st, ids, new, upd, q = g.integrate_frame(fr.index, fr.is_keyframe, fr.pose, cam)
        oi, onew = o.prepare(fr.depth, fr.pose, cam)
        onu, _ = o.integrate(fr.depth, rgba, fr.quality if fr.is_keyframe else None, fr.pose, cam, oi, 1, -1)
        o.finalize(oi, onu, onew)
        assert np.array_equal(ids, oi) and np.array_equal(upd != 0, np.asarray(onu) != 0)
    group = seq.frames[3:6]
    for fr in group:
        g.upload_frame(fr.index, fr.depth, fr.rgba() if fr.is_keyframe else None, fr.quality if fr.is_keyframe else None)
    assert group[0].is_keyframe
    out = g.integrate_batch([{"flag": 1, "frames": [(fr.index, k == 0, fr.pose) for k, fr in enumerate(group)]}], cam)
    ov, oq = _oracle_keyframe_group(o, cam, group, [fr.pose for fr in group], 1)
    assert np.array_equal(out[0][0], ov) and np.array_equal(out[0][1].view(np.uint32), oq.view(np.uint32))
    assert assert_maps_equal(g, o, what="noisy depth")
    g.close()


def test_maps_on_two_devices_in_one_process():
    """Every entry point selects its map's device: two maps on different GPUs driven alternately from
    one thread (whatever the thread's current device is) must both match the oracle."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    res = 0.02
    seq = room_sequence(4)
    cam = seq.cam
    maps = [capi.Map(res, device=0), capi.Map(res, device=1)]
    o = OracleMap(res)
    for fr in seq.frames:
        rgba = fr.rgba() if fr.is_keyframe else None
        o.integrate_frame(fr.depth, rgba, fr.quality, fr.pose, cam, -1)
        for k, m in enumerate(maps):
            torch.cuda.set_device(1 - k)  # deliberately the other one
            m.upload_frame(fr.index, fr.depth, rgba, fr.quality if fr.is_keyframe else None)
            m.integrate_frame(fr.index, fr.is_keyframe, fr.pose, cam)
    torch.cuda.set_device(0)
    for k, m in enumerate(maps):
        assert assert_maps_equal(m, o, what=f"map on device {k}")
        m.close()


def test_reupload_replaces_the_slot_and_keyframes_stay_pinned():
    """(a) re-uploading a frame index WITHOUT colour / quality must not leave the previous
    upload's planes marked valid (the reference reads NULL as "no colour" / quality 0); (b) key-frames
    uploaded with tf_upload_keyframe_rgb are not evicted by later uploads; (c) an upload that overwrites
    a slot is ordered behind queued kernels that read it."""
    seq = room_sequence(6)
    cam = seq.cam
    kf, lf = seq.frames[0], seq.frames[1]
    g = capi.Map(0.02, max_frames=3)
    o = OracleMap(0.02)
    SCR = 77
    # key-frame with colour + quality, then the same index re-uploaded depth-only
    g.upload_frame(SCR, kf.depth, kf.rgba(), kf.quality)
    ids, new = g.prepare(SCR, kf.pose, cam)
    oi, on = o.prepare(kf.depth, kf.pose, cam)
    assert np.array_equal(ids, oi)
    nu, q = g.integrate(SCR, True, kf.pose, cam, ids, 1)
    ou, oq = o.integrate(kf.depth, kf.rgba(), kf.quality, kf.pose, cam, oi, 1, kf.index)
    assert np.array_equal(nu, ou) and np.array_equal(q.view(np.uint32), oq.view(np.uint32))
    g.upload_frame(SCR, lf.depth)  # no colour, no quality
    with pytest.raises(capi.TexFusionError) as e:  # asking for colour now must fail, not use the stale plane
        g.integrate(SCR, True, lf.pose, cam, ids, 1)
    assert e.value.code == capi.TF_ERR_NOT_FOUND
    nu, q = g.integrate(SCR, False, lf.pose, cam, ids, 1, nu)
    ou, _ = o.integrate(lf.depth, None, None, lf.pose, cam, oi, 1, -1, ou)
    assert np.array_equal(nu, ou)
    # colour without a quality plane: the sums must be those of a NULL observationQualityPointer (0 / sentinel), not stale data
    g.upload_frame(SCR, kf.depth, kf.rgba(), None)
    nu2, q2 = g.integrate(SCR, True, kf.pose, cam, ids, 1)
    ou2, oq2 = o.integrate(kf.depth, kf.rgba(), None, kf.pose, cam, oi, 1, kf.index)
    assert np.array_equal(q2.view(np.uint32), oq2.view(np.uint32))
    assert_maps_equal(g, o, what="re-upload protocol")
    # (b) pinned key-frames survive later uploads; the store refuses when only pinned slots remain
    p = capi.Map(0.02, max_frames=3)
    for k in (0, 1):
        p.upload_frame(100 + k, kf.depth)
        p.upload_keyframe_rgb(100 + k, kf.rgb)
    for k in range(5):  # all of these share the one evictable slot
        p.upload_frame(k, lf.depth)
    loc = p.atlas_alloc_slot((0, 0, 0))
    p.atlas_update([(loc, 100, 10, 10, 8, 8), (p.atlas_alloc_slot((1, 0, 0)), 101, 20, 20, 8, 8)])  # both key-frames still there
    p.upload_frame(102, kf.depth)
    p.upload_keyframe_rgb(102, kf.rgb)  # third pinned slot: nothing evictable left
    with pytest.raises(capi.TexFusionError) as e:
        p.upload_frame(7, lf.depth)
    assert e.value.code == capi.TF_ERR_CAPACITY
    p.release_frame(100)
    p.upload_frame(7, lf.depth)
    # (c) overwrite ordering: queue a fused frame asynchronously, overwrite its slot at once, collect
    s = capi.Map(0.02, max_frames=2)
    s2 = OracleMap(0.02)
    import ctypes as C
    s.upload_frame(1, kf.depth)
    st = capi.FrameStats()
    assert s.L.tf_integrate_frame_begin(s.h, 1, 0, C.byref(capi.make_pose(kf.pose)), C.byref(capi.make_camera(cam)), None, None,
                                        None, None, 0) == 0
    s.upload_frame(1, np.zeros_like(kf.depth))  # must wait for the kernels that read slot 1
    assert s.L.tf_integrate_frame_end(s.h, C.byref(st)) == 0
    n, n_upd = s2.integrate_frame(kf.depth, None, None, kf.pose, cam, -1)
    assert (st.n_chunks, st.n_updated) == (n, n_upd)
    assert_maps_equal(s, s2, what="overwrite during an in-flight frame")
