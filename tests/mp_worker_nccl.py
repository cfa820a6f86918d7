"""Worker of tests/test_multiprocess_gpu.py: one process per GPU (torch.distributed.run), NCCL.

Drives exactly bench.py's multi-GPU end-to-end sequence (texturefusion_b200.streaming.FrameStreamer:
rank 0 uploads from page-locked memory, tf_broadcast_frame on the copy stream, tf_integrate_frame with
ordered lists on every rank's shard, tf_wait_upload) over a run that contains key-frames, then gathers
every rank's per-frame lists and final chunk hashes on rank 0, which fuses the same frames with the CPU
oracle and asserts: per-frame lists (ids in reference order, created / updated flags, quality sums) and
the union of the ranks' maps are identical to the single-process CPU result, bit for bit."""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from texturefusion_b200 import capi, sharding, synth  # noqa: E402
from texturefusion_b200.maphash import sorted_chunk_hashes  # noqa: E402
from texturefusion_b200.streaming import FrameStreamer  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=32)
    ap.add_argument("--res", type=float, default=0.005)
    ap.add_argument("--first-frame", type=int, default=0)
    args = ap.parse_args()
    rank, local_rank, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local_rank)
    dev = f"cuda:{local_rank}"
    dist.init_process_group("nccl", device_id=torch.device(dev))
    cam = synth.Camera()
    seq = synth.make_sequence(args.frames, cam=cam, total=300, keyframe_every=10, device=dev, start=args.first_frame)
    frames = seq.frames
    m = capi.Map(args.res, device=local_rank, n_ranks=world, rank=rank, max_frames=8, max_chunks=1 << 18)
    m.comm_init(capi.share_unique_id(dist, dev))
    fs = FrameStreamer(m, frames, cam, rank=rank, world=world, cap=1 << 17)
    fs.prime(0)
    per_frame = []
    for i in range(args.frames):
        fs.step(i)  # (the frames staged past the end wrap around to the first ones and are never fused)
        per_frame.append(fs.lists())
    m.sync()
    ids, hashes = sorted_chunk_hashes(m)
    gathered = [None] * world if rank == 0 else None
    dist.gather_object((per_frame, ids, hashes), gathered, dst=0)
    ok = True
    if rank == 0:
        from oracle import OracleMap
        o = OracleMap(args.res, threads=0)
        for i, fr in enumerate(frames):
            rg = fr.rgba() if fr.is_keyframe else None
            q = fr.quality if fr.is_keyframe else None
            oi, on = o.prepare(fr.depth, fr.pose, cam)
            ou, oq = o.integrate(fr.depth, rg, q, fr.pose, cam, oi, 1, fr.index if fr.is_keyframe else -1)
            o.finalize(oi, ou, on)
            own = sharding.owner_of(oi, world)
            for r in range(world):
                gi, gn, gu, gq = gathered[r][0][i]
                sel = own == r
                assert np.array_equal(gi, oi[sel]), f"frame {i} rank {r}: list differs from the reference order restricted to its shard"
                assert np.array_equal(gn, on[sel]) and np.array_equal(gu, ou[sel]), f"frame {i} rank {r}: flags"
                if fr.is_keyframe:
                    assert np.array_equal(gq.view(np.uint32), oq[sel].view(np.uint32)), f"frame {i} rank {r}: quality sums"
        oi, oh = sorted_chunk_hashes(o)
        ui = np.concatenate([g[1] for g in gathered])
        uh = np.concatenate([g[2] for g in gathered])
        order = np.lexsort((ui[:, 2], ui[:, 1], ui[:, 0]))
        ui, uh = ui[order], uh[order]
        assert ui.shape == oi.shape and np.array_equal(ui, oi), f"union of the shards: {len(ui)} chunks vs {len(oi)}"
        assert np.array_equal(uh, oh), f"{int((uh != oh).sum())} chunks differ from the CPU result"
        for r in range(world):
            assert np.all(sharding.owner_of(gathered[r][1], world) == r)
        print(f"OK: {args.frames} frames, {world} ranks, {len(oi)} chunks, shards {[len(g[1]) for g in gathered]}", flush=True)
    m.close()
    dist.barrier()
    dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
