"""BASELINE.json configs[3] (C4): a long walk through the 40 x 25 x 3 m building at 5 mm voxels, on one GPU
or chunk-sharded over N GPUs (one process per GPU, frames broadcast over NVLink by the library).

Frames are rendered on the fly on rank 0 (not timed), uploaded from page-locked host memory and fused one
call at a time, double buffered like bench.py's e2e leg (upload + tf_broadcast_frame of frame k+1 in flight
during the kernels of frame k).  Reports throughput per window of frames next to the size of the map —
i.e. whether the per-frame cost depends on how much has been mapped — and at the end the chunks and HBM
held per rank (load balance of the ownership hash) and the map checksum (texturefusion_b200.maphash).

  python tools/bench_building.py [--frames 5000] [--res 0.005] [--max-chunks 4194304]
  python -m torch.distributed.run --nproc-per-node N ... tools/bench_building.py --gpus N"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from texturefusion_b200 import capi, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=5000)
    ap.add_argument("--total", type=int, default=5000, help="poses of the whole walk (frames are its first --frames)")
    ap.add_argument("--res", type=float, default=0.005)
    ap.add_argument("--max-chunks", type=int, default=1 << 22, help="chunk pool of the whole job (split over the ranks, + 50 %% slack)")
    ap.add_argument("--window", type=int, default=500)
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--hash", action="store_true", help="also compute the checksum of the final map (downloads it)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = f"cuda:{local_rank}"
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device(dev))
    cam = synth.Camera()
    per_rank = args.max_chunks if world == 1 else int(args.max_chunks * 1.5 / world)
    m = capi.Map(args.res, device=local_rank, n_ranks=world, rank=rank, max_frames=8, max_chunks=per_rank)
    if dist is not None:
        m.comm_init(capi.share_unique_id(dist, dev))
    L = m.L
    vp = C.c_void_p
    camc = capi.make_camera(cam)
    st = capi.FrameStats()
    npix = cam.width * cam.height
    pins = [(capi.PinnedBuffer((npix,), np.float32), capi.PinnedBuffer((npix * 4,), np.uint8), capi.PinnedBuffer((npix,), np.float32))
            for _ in range(3)] if rank == 0 else None
    # sharded maps ingest two frames ahead (see texturefusion_b200/streaming.py: the broadcast needs an SM slot
    # and gets one in the bbox / culling phase of the frame in between)
    ahead = 2 if world > 1 else 1

    def render(k):  # rank 0 only
        pose = synth.walk_pose(k, args.total)
        kf = k % 10 == 0
        depth, rgb, q = synth.render(pose, cam, color=kf, device=dev, scene="building")
        fr = synth.Frame(k, pose, depth, rgb, np.ones(depth.shape, np.uint8) if kf else None, q, kf)
        d, c, qq = pins[k % 3]
        d.array[:] = np.asarray(fr.depth, np.float32).ravel()
        if kf:
            c.array[:] = np.asarray(fr.rgba(), np.uint8).ravel()
            qq.array[:] = np.asarray(fr.quality, np.float32).ravel()

    def stage(k):
        kf = k % 10 == 0
        if rank == 0:
            d, c, qq = pins[k % 3]
            rc = L.tf_upload_frame(m.h, k, vp(d.ptr), vp(c.ptr) if kf else None, vp(qq.ptr) if kf else None)
            assert rc == 0, L.tf_last_error(m.h)
        if world > 1:
            rc = L.tf_broadcast_frame(m.h, k, int(kf), 0)
            assert rc == 0, L.tf_last_error(m.h)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    rows = []
    for j in range(min(ahead, args.frames)):
        if rank == 0:
            render(j)
        stage(j)
    t_win, n_win, vox_win = 0.0, 0, 0
    for k in range(args.frames):
        has_next = k + 1 < args.frames
        has_ahead = k + ahead < args.frames
        if has_ahead and rank == 0:
            render(k + ahead)  # (not timed; buffer (k + ahead) % 3 was last used by frame k + ahead - 3, uploaded long ago)
        pose = capi.make_pose(synth.walk_pose(k, args.total))
        barrier()
        t0 = time.perf_counter()
        if has_ahead:
            stage(k + ahead)
        rc = L.tf_integrate_frame(m.h, k, int(k % 10 == 0), C.byref(pose), C.byref(camc), C.byref(st), None, None, None, None, 0)
        assert rc == 0, L.tf_last_error(m.h)
        if has_next:
            assert L.tf_wait_upload(m.h, k + 1) == 0
        t_win += time.perf_counter() - t0
        n_win += 1
        vox_win += st.voxel_updates
        if n_win == args.window or not has_next:
            stats = torch.tensor([t_win, float(vox_win), float(m.chunk_count())], dtype=torch.float64, device=dev)
            if dist is not None:
                tmax = stats.clone()
                dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
                dist.all_reduce(stats)
                t_job, vox_job, live = float(tmax[0]), float(stats[1]), float(stats[2])
            else:
                t_job, vox_job, live = t_win, float(vox_win), float(m.chunk_count())
            if rank == 0:
                rows.append({"frames": k + 1, "fps_e2e": n_win / t_job, "chunks_per_frame": vox_job / 512 / n_win,
                             "voxel_updates_per_s": vox_job / t_job, "live_chunks": int(live), "map_GB": live * 8192 / 1e9})
                print(json.dumps(rows[-1]), flush=True)
            t_win, n_win, vox_win = 0.0, 0, 0
    m.sync()
    free_b, total_b = torch.cuda.mem_get_info()
    mine = {"rank": rank, "chunks": int(m.chunk_count()), "map_GB": m.chunk_count() * 8192 / 1e9,
            "hbm_used_GB": (total_b - free_b) / 1e9, "pool_capacity": per_rank}
    if args.hash:
        from texturefusion_b200.maphash import map_hash
        mine["map_hash_part"] = map_hash(m)[1]
    parts = [mine]
    if dist is not None:
        parts = [None] * world if rank == 0 else None
        dist.gather_object(mine, parts, dst=0)
    if rank == 0:
        chunks = [p["chunks"] for p in parts]
        total_frames = sum(1 for _ in range(args.frames))
        total_t = sum((r["frames"] - (rows[i - 1]["frames"] if i else 0)) / r["fps_e2e"] for i, r in enumerate(rows))
        line = {"metric": "building-scale fusion: frames/s end to end (upload + broadcast + fuse)", "n_gpus": world,
                "value": total_frames / total_t, "unit": "frames/s",
                "config": {"workload": f"configs[3]: {args.frames} frames of the {args.total}-pose walk through the 40x25x3 m building, "
                                       f"{args.res} m voxels, chunk-sharded x{world}"},
                "map_chunks": int(sum(chunks)), "active_voxels": int(sum(chunks)) * 512, "map_GB": sum(chunks) * 8192 / 1e9,
                "per_rank": parts, "load_balance_max_over_mean": max(chunks) / (sum(chunks) / len(chunks)),
                "windows": rows}
        if args.hash:
            line["map_hash"] = f"{sum(p['map_hash_part'] for p in parts) & ((1 << 64) - 1):016x}"
        print(json.dumps(line), flush=True)
    m.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
