"""Host-side logic of the chunk-sharded multi-GPU path, on CPU: the ownership function, the
order-restoring merge, and a world_size-2 gloo run of broadcast -> per-rank work -> gather."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import OracleMap
from texturefusion_b200 import sharding, synth
from texturefusion_b200.distributed import broadcast_frame, frame_header, gather_lists


def _frame():
    cam = synth.Camera().scaled(0.25)
    fr = synth.make_sequence(1, cam=cam, total=300, keyframe_every=1, start=30).frames[0]
    return cam, fr


def test_owner_is_constant_inside_4x4x4_blocks_and_balanced():
    rng = np.random.RandomState(0)
    ids = rng.randint(-500, 500, size=(20000, 3)).astype(np.int32)
    for n in (2, 4, 8):
        own = sharding.owner_of(ids, n)
        base = (ids >> 2) << 2
        assert np.array_equal(own, sharding.owner_of(base, n))
        assert np.array_equal(own, sharding.owner_of(base + 3, n))
        counts = np.bincount(own, minlength=n)
        assert counts.min() > 0.8 * len(ids) / n
    # negative coordinates use floor division (arithmetic shift), like the device code
    assert sharding.owner_of([[-1, -1, -1]], 8)[0] == sharding.owner_of([[-4, -4, -4]], 8)[0]


@pytest.mark.parametrize("res", (0.02, 0.01))
def test_merge_restores_reference_order(res):
    cam, fr = _frame()
    o = OracleMap(res)
    ids = o.observed_ids(fr.depth, fr.pose, cam)
    lo, _ = o.boundary_ids(fr.depth, fr.pose, cam)
    step = 1 if np.float32(res) > 0.01 else 4
    for n in (2, 4, 8):
        parts, own = sharding.split_by_owner(ids, n)
        assert sum(len(p) for p in parts) == len(ids)
        payload = [(np.arange(len(ids))[own == r],) for r in range(n)]
        merged, idx = sharding.merge_rank_lists(parts, payload, min_id=lo, step=step)
        assert np.array_equal(merged, ids)
        assert np.array_equal(idx, np.arange(len(ids)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, res, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cam, fr = _frame()
    H, W = cam.height, cam.width
    if rank == 0:
        depth = torch.from_numpy(fr.depth.copy())
        header = frame_header(fr.index, True, fr.pose)
    else:  # every other rank receives the frame: no data of its own
        depth = torch.zeros((H, W), dtype=torch.float32)
        header = torch.zeros(18, dtype=torch.float64)
    header = broadcast_frame([depth], header, src=0)
    pose = header[2:].reshape(4, 4).numpy().astype(np.float32)
    assert int(header[0]) == fr.index
    # each rank runs the full per-frame path on the chunks it owns (CPU stand-in for the kernels)
    o = OracleMap(res)
    ids, new = o.prepare(depth.numpy(), pose, cam)
    mine = sharding.owner_of(ids, world) == rank
    garbage = ids[~mine]
    o.finalize(garbage, np.zeros(len(garbage), np.uint8), np.ones(len(garbage), np.uint8))  # drop foreign chunks
    nu, _ = o.integrate(depth.numpy(), None, None, pose, cam, ids[mine], 1)
    valid = o.finalize(ids[mine], nu, new[mine])
    sdf, w, _ = o.download_chunks(valid)
    out = gather_lists(valid, (w.sum(axis=1),), dst=0)
    if rank == 0:
        q.put(out)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_broadcast_shard_gather():
    res, world = 0.02, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, res, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process result
    cam, fr = _frame()
    o = OracleMap(res)
    ids, new = o.prepare(fr.depth, fr.pose, cam)
    nu, _ = o.integrate(fr.depth, None, None, fr.pose, cam, ids, 1)
    valid = o.finalize(ids, nu, new)
    w = o.download_chunks(valid)[1].sum(axis=1)
    merged, mw = sharding.merge_rank_lists([o_[0] for o_ in out], [o_[1] for o_ in out])
    order = np.lexsort((valid[:, 2], valid[:, 1], valid[:, 0]))
    assert np.array_equal(merged, valid[order])
    assert np.array_equal(mw, w[order])
    assert all(len(o_[0]) > 0 for o_ in out)
