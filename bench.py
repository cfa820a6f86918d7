#!/usr/bin/env python
"""bench.py — fused RGB-D frames/s of the fusion hot path (BASELINE.json metric).

Workload (config.workload): BASELINE.json configs[1] — the seeded synthetic 640x480 room
sequence (300 poses, every 10th frame a key-frame with colour + quality, the others
depth-only) fused at 0.005 m voxels.  One "step" = one frame through
Prepare + Integrate + Finalize (MobileFusion::IntegrateFrame, GCFusion/MobileFusion.cpp:223-250).

  value     frames/s with the frames already resident in HBM (CUDA events around the step's kernels)
  e2e       frames/s through the public C ABI with HOST (pinned) buffers: H2D of the frame's
            planes and D2H of the chunk list / flags / quality sums inside the timed region
  roofline  integrate_kernel: algorithmic bytes / CUDA-event kernel time vs measured HBM peak
  cpu_baseline  the CPU oracle (AVX2 restatement of the reference) on a bounded sample

`--impl reference` times the reference's CPU algorithm (oracle port; the reference itself
cannot be built here) with the reference's thread policy on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SEQ_FRAMES = 300
KEYFRAME_EVERY = 10
METRIC = "fused RGB-D frames/s (640x480, 5 mm voxels)"


def load_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def load_traffic(steps, warmup):
    """dram__bytes_read.sum + dram__bytes_write.sum per integrate_kernel launch from the ncu capture
    of THIS command line (profiles/r2_integrate_traffic.json records the --steps/--warmup it was taken
    with and the launches it averaged); None when no capture matches the current arguments — a
    figure from a different workload would not be comparable with `achieved`."""
    p = os.path.join(ROOT, "profiles", "r2_integrate_traffic.json")
    try:
        d = json.load(open(p))
        if int(d["steps"]) == steps and int(d["warmup"]) == warmup:
            return float(d["dram_bytes_per_launch"]), d["note"]
        return None, f"capture in profiles/ is for --steps {d['steps']} --warmup {d['warmup']}, not this run"
    except Exception:
        return None, "no ncu capture committed"


class ClockSampler:
    """Samples the SM clock and the throttle reasons during the timed regions: NVML in-process every 10 ms
    (the timed regions are fractions of a second), nvidia-smi as the fallback."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu=0):
        self.gpu, self.samples, self.stop_flag, self.th = gpu, [], False, None
        self.nvml = self.handle = None
        try:
            import pynvml
            pynvml.nvmlInit()
            # (CUDA_VISIBLE_DEVICES re-numbers CUDA devices, NVML does not)
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[gpu]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else gpu
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n = self.nvml
        sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
        mx = n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)
        r = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
        flags = ["Active" if r & bit else "Not Active" for bit in
                 (n.nvmlClocksEventReasonHwSlowdown, n.nvmlClocksEventReasonHwThermalSlowdown,
                  n.nvmlClocksEventReasonSwThermalSlowdown, n.nvmlClocksEventReasonSwPowerCap)]
        return [str(sm), str(mx)] + flags

    def _run(self):
        while not self.stop_flag:
            try:
                if self.nvml is not None:
                    self.samples.append(self._sample_nvml())
                    time.sleep(0.01)
                    continue
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.gpu)], capture_output=True, text=True, timeout=5).stdout
                self.samples.append([x.strip() for x in out.strip().split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def start(self):
        self.th = threading.Thread(target=self._run, daemon=True)
        self.th.start()

    def stop(self):
        self.stop_flag = True
        if self.th:
            self.th.join(timeout=6)
        sm, mx, reasons = [], [], set()
        for s in self.samples:
            try:
                sm.append(float(s[0]))
                mx.append(float(s[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def make_data(n_frames, device):
    from texturefusion_b200 import synth
    cam = synth.Camera()
    t0 = time.time()
    seq = synth.make_sequence(n_frames, cam=cam, total=SEQ_FRAMES, keyframe_every=KEYFRAME_EVERY, device=device)
    return seq, time.time() - t0


def cpu_impl():
    """("ref", "reference") when oracle/_ref — the reference's own sources — is built, else the restated port."""
    from oracle import have_ref
    return ("ref", "reference") if have_ref() else ("port", "port")


def run_cpu_reference(seq, res, steps, warmup, budget_s, threads, want_hash=False):
    """The reference's CPU path: Prepare + Integrate + Finalize per frame (IntegrateFrame shape)."""
    from oracle import OracleMap
    o = OracleMap(res, threads=threads, impl=cpu_impl()[0])
    frames = seq.frames
    rgba = {fr.index: fr.rgba() for fr in frames if fr.is_keyframe}
    k = 0
    for _ in range(warmup):
        fr = frames[k % len(frames)]
        o.integrate_frame(fr.depth, rgba.get(fr.index), fr.quality, fr.pose, seq.cam, fr.index if fr.is_keyframe else -1)
        k += 1
    t_total, done, vox = 0.0, 0, 0
    for _ in range(steps):
        fr = frames[k % len(frames)]
        t0 = time.perf_counter()
        n, _ = o.integrate_frame(fr.depth, rgba.get(fr.index), fr.quality, fr.pose, seq.cam,
                                 fr.index if fr.is_keyframe else -1)
        t_total += time.perf_counter() - t0
        vox += n * 512
        done += 1
        k += 1
        if t_total > budget_s:
            break
    out = {"fps": done / t_total, "frames": done, "seconds": t_total, "voxel_updates_per_s": vox / t_total,
           "cores": o.threads_used}
    if want_hash:
        from texturefusion_b200.maphash import map_hash
        out["map_chunks"], out["map_hash"] = map_hash(o)
    return out


def flush_l2(torch, buf):
    buf.fill_(1.0)  # 512 MiB write > 126 MB L2


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=SEQ_FRAMES)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--res", type=float, default=0.005)
    ap.add_argument("--cpu-budget", type=float, default=10.0, help="seconds of CPU work for cpu_baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)  # both arms: they fuse the same frames

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    import torch
    have_cuda = torch.cuda.is_available()
    n_need = min(SEQ_FRAMES, args.steps + args.warmup)
    config = {"workload": "configs[1]: synthetic 640x480 room sequence (300 poses, key-frame every 10th frame with "
                          "colour+quality, others depth-only), TSDF + voxel colour fusion, Prepare+Integrate+Finalize per frame",
              "frames_fused": f"frames 0..{n_need - 1} of the sequence into an empty map: {args.warmup} warm-up + {args.steps} "
                              "timed (past frame 299 the sequence repeats)",
              "voxel_res_m": args.res, "frames_in_sequence": SEQ_FRAMES, "image": "640x480",
              "l2": "flushed between timed steps (512 MiB write), flush excluded from the timed region",
              "parallelism": f"chunk-sharded x{args.gpus}" if args.gpus > 1 else "single GPU"}

    # ---------------- reference arm: CPU implementation on the host cores -----------------
    if args.impl == "reference":
        if rank != 0:
            return
        seq, _ = make_data(n_need, "cuda" if have_cuda else "cpu")
        # the same frames as our arm (same warm-up, same steps, empty map); bounded by a time budget only
        r = run_cpu_reference(seq, args.res, args.steps, args.warmup, 120.0, threads=0, want_hash=True)
        kind = cpu_impl()[1]
        what = ("the reference's own sources (oracle/_ref: ProjectionIntegrator.cpp + ChunkManager.{h,cpp} built against the "
                "Eigen stand-in)" if kind == "reference" else "oracle port")
        line = {"impl": "reference", "metric": METRIC, "value": r["fps"], "unit": "frames/s", "n_gpus": args.gpus,
                "steps": r["frames"], "warmup": args.warmup, "ms_per_step": 1e3 / r["fps"],
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": config,
                "cpu_baseline": {"value": r["fps"], "unit": "frames/s", "cores": r["cores"], "kind": kind,
                                 "sample": f"{r['frames']} frames of the workload after {args.warmup} warm-up frames, "
                                           f"{what}, the reference's parallel_for policy (hardware_concurrency-2 threads)"},
                "e2e": {"value": r["fps"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "voxel_updates_per_s": r["voxel_updates_per_s"],
                "map_chunks": r["map_chunks"], "map_hash": f"{r['map_hash']:016x}"}
        print(json.dumps(line))
        return

    # ---------------- our arm ------------------------------------------------------------------
    if not have_cuda:
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    from texturefusion_b200 import capi
    import ctypes as C

    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    dev = f"cuda:{local_rank}"
    seq, gen_s = make_data(n_need, dev)
    cam = seq.cam
    frames = seq.frames
    nf = len(frames)
    peak_gbs, peak_src = load_peak()

    def new_map():
        return capi.Map(args.res, device=local_rank, n_ranks=world, rank=rank, max_frames=nf + 8,
                        max_chunks=1 << 19)

    flush_buf = torch.empty(128 * 1024 * 1024, dtype=torch.float32, device=dev)
    rgba = {fr.index: fr.rgba() for fr in frames if fr.is_keyframe}

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # arguments marshalled once: the timed region contains the C-ABI call, not Python glue
    camc = capi.make_camera(cam)
    poses = [capi.make_pose(fr.pose) for fr in frames]
    st = capi.FrameStats()

    # (one C call per step with its arguments in a struct built once: tf_stream_step without an ingest is
    #  tf_integrate_frame; from Python a call with four arguments costs about a microsecond less than one with eleven)
    steps_resident = []
    for i, fr in enumerate(frames):
        a = capi.StreamStepArgs()
        a.frame_index, a.use_color, a.pose, a.next_index, a.wait_index = fr.index, int(fr.is_keyframe), poses[i], -1, -1
        steps_resident.append(a)

    def fuse_resident(mp, i):
        rc = mp.L.tf_stream_step(mp.h, C.byref(camc), C.byref(steps_resident[i]), C.byref(st))
        if rc != 0:
            raise RuntimeError(mp.L.tf_last_error(mp.h))
        return st

    # ===== pass 1: frames resident in HBM, device-event timing =====================================
    m = new_map()
    for fr in frames:
        m.upload_frame(fr.index, fr.depth, rgba.get(fr.index), fr.quality if fr.is_keyframe else None)
    m.sync()
    ext = torch.cuda.ExternalStream(m.stream(), device=dev)
    ev_a = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ev_b = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    k = 0
    for _ in range(args.warmup):
        fuse_resident(m, k % nf)
        k += 1
    c0 = m.counters()
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    vox = 0
    chunks = 0
    # The step's kernels are queued between the two events (tf_integrate_frame_begin), the closing event is
    # recorded behind them on the same stream, then the host collects the frame (tf_integrate_frame_end): the
    # timed region is the device time of the step from an idle GPU — launch latency included, the host's
    # completion polling not (that is part of `e2e`).
    L, vp0 = m.L, C.c_void_p(None)
    for s in range(args.steps):
        fr = frames[k % nf]
        flush_l2(torch, flush_buf)
        torch.cuda.synchronize()
        ev_a[s].record(ext)
        rc = L.tf_integrate_frame_begin(m.h, fr.index, int(fr.is_keyframe), C.byref(poses[k % nf]), C.byref(camc), vp0, vp0, vp0, vp0, 0)
        ev_b[s].record(ext)
        rc = rc or L.tf_integrate_frame_end(m.h, C.byref(st))
        if rc != 0:
            raise RuntimeError(L.tf_last_error(m.h))
        vox += st.voxel_updates
        chunks += st.n_chunks
        k += 1
    barrier()
    c1 = m.counters()
    dev_ms = sum(a.elapsed_time(b) for a, b in zip(ev_a, ev_b))
    dev_ms = max_over_ranks(dev_ms)
    if dist is not None:  # whole-job totals: every rank fused its own share of the chunks
        tot = torch.tensor([float(vox), float(chunks)], dtype=torch.float64, device=dev)
        dist.all_reduce(tot)
        vox, chunks = int(tot[0].item()), int(tot[1].item())
    launches = c1["kernel_launches"] - c0["kernel_launches"]
    value = args.steps / (dev_ms * 1e-3)
    live_chunks = m.chunk_count()
    m.close()

    # ===== pass 2: kernel-level timing of integrate_kernel (roofline) ================================
    m = new_map()
    for fr in frames:
        m.upload_frame(fr.index, fr.depth, rgba.get(fr.index), fr.quality if fr.is_keyframe else None)
    m.sync()
    k = 0
    for _ in range(args.warmup):
        fuse_resident(m, k % nf)
        k += 1
    m.set_profiling(2)
    m.kernel_time(reset=True)
    m.stage_times(reset=True)
    for s in range(args.steps):
        fr = frames[k % nf]
        flush_l2(torch, flush_buf)
        torch.cuda.synchronize()
        fuse_resident(m, k % nf)
        k += 1
    k_ms, k_n, k_bytes = m.kernel_time(reset=True)
    stage_us = {k: 1e3 * v / args.steps for k, v in m.stage_times().items()}
    m.close()
    achieved = (k_bytes / 1e9) / (k_ms * 1e-3) if k_ms > 0 else 0.0
    traffic, traffic_note = load_traffic(args.steps, args.warmup)
    roofline = {"bound": "hbm", "kernel": "integrate_kernel", "achieved": achieved, "peak": peak_gbs, "unit": "GB/s",
                "frac": achieved / peak_gbs, "traffic": traffic, "traffic_note": traffic_note,
                # real DRAM bytes / kernel time / peak (only when the capture is of this command)
                "frac_dram": (traffic / 1e9) / (1e-3 * k_ms / max(k_n, 1)) / peak_gbs if traffic and k_ms > 0 else None,
                "peak_source": peak_src,
                "launches": k_n, "avg_launch_us": 1e3 * k_ms / max(k_n, 1),
                "algorithmic_bytes_per_launch": k_bytes / max(k_n, 1),
                "kernel_share_of_step": k_ms / dev_ms if dev_ms > 0 else None}

    # ===== pass 3: end to end through the C ABI with host buffers ======================================
    # Every timed step contains the H2D copy of one frame from page-locked memory, the fused call
    # and the D2H read of its results (frame statistics + the ordered chunk lists, written into
    # page-locked output buffers).  The loop is double buffered the way a streaming caller would
    # write it: tf_upload_frame(i+1) is issued before tf_integrate_frame(i), so the copy of the next
    # frame overlaps the kernels of the current one; tf_wait_upload at the end of the step makes sure that
    # copy has finished inside the timed region.
    from texturefusion_b200.streaming import FrameStreamer
    m = capi.Map(args.res, device=local_rank, n_ranks=world, rank=rank, max_frames=32, max_chunks=1 << 19)
    if dist is not None:  # NCCL communicator of the library itself (the id travels through torch.distributed, once)
        m.comm_init(capi.share_unique_id(dist, dev))
    # Multi-GPU: rank 0 ingests frame i+1 (H2D on the map's copy stream) and tf_broadcast_frame moves its
    # planes over NVLink into every rank's frame store with one ncclBroadcast ON THAT STREAM, i.e. behind
    # the copy and next to the kernels of frame i; the compute stream waits for the slot's event.
    fs = FrameStreamer(m, frames, cam, rank=rank, world=world, cap=1 << 16)
    e2e_step = fs.step

    k = 0
    fs.prime(0)
    m.sync()
    for _ in range(args.warmup):
        e2e_step(k % nf)
        k += 1
    c0 = m.counters()
    barrier()
    e2e_s = 0.0
    step_s = []
    for s in range(args.steps):
        flush_l2(torch, flush_buf)
        barrier() if dist is not None else torch.cuda.synchronize()
        t0 = time.perf_counter()
        e2e_step(k % nf)
        dt = time.perf_counter() - t0
        e2e_s += dt
        step_s.append(dt)
        k += 1
    barrier()
    clocks = sampler.stop()  # (sampled over all three timed passes: resident, kernel-level, end to end)
    c1 = m.counters()
    e2e_s = max_over_ranks(e2e_s)
    e2e = {"value": args.steps / e2e_s, "unit": "frames/s",
           "h2d_bytes_per_step": (c1["h2d_bytes"] - c0["h2d_bytes"]) / args.steps,
           "d2h_bytes_per_step": (c1["d2h_bytes"] - c0["d2h_bytes"]) / args.steps,
           "ms_per_step": 1e3 * e2e_s / args.steps, "median_ms_per_step": 1e3 * float(np.median(step_s)),
           "mode": ("one tf_stream_step call per step = tf_integrate_frame_begin(i) | tf_upload_frame(i+1) | "
                    "tf_integrate_frame_end(i) | tf_wait_upload(i+1) inside the library, host buffers page-locked" if dist is None else
                    "one tf_stream_step call per step and rank = tf_integrate_frame_begin(i) on the rank's shard | rank 0 uploads "
                    "frame i+2, tf_broadcast_frame (one ncclBroadcast inside the library, on the copy stream) | "
                    "tf_integrate_frame_end(i) | tf_wait_upload(i+1)")}
    # The map the e2e pass built (frames 0..warmup+steps-1 from empty), as a shard-independent checksum:
    # the same at every N and in the reference arm's line if and only if the fused maps are identical.
    from texturefusion_b200.maphash import map_hash
    m.sync()
    map_chunks, mh = map_hash(m)
    if dist is not None:
        parts = torch.tensor([map_chunks, mh & 0xffffffff, mh >> 32], dtype=torch.int64, device=dev)
        allp = [torch.zeros_like(parts) for _ in range(world)]
        dist.all_gather(allp, parts)
        map_chunks = int(sum(int(p[0]) for p in allp))
        mh = sum((int(p[2]) << 32) | int(p[1]) for p in allp) & ((1 << 64) - 1)
    m.close()

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ===== CPU baseline (rank 0, N=1 only): oracle port on a bounded sample ================================
    # The same frames as the GPU arm, from an empty map; repeated (fresh map each time) until about
    # --cpu-budget seconds of CPU work have been timed, mean over the repetitions.
    cpu_baseline = None
    if args.gpus == 1 and not args.no_cpu_baseline:
        runs, spent = [], 0.0
        while spent < args.cpu_budget and len(runs) < 8:
            r = run_cpu_reference(seq, args.res, min(args.steps, nf), args.warmup, args.cpu_budget, threads=0)
            runs.append(r)
            spent += r["seconds"]
        fps = sum(r["frames"] for r in runs) / sum(r["seconds"] for r in runs)
        vps = sum(r["voxel_updates_per_s"] * r["seconds"] for r in runs) / sum(r["seconds"] for r in runs)
        cpu_baseline = {"value": fps, "unit": "frames/s", "cores": runs[0]["cores"], "kind": cpu_impl()[1],
                        "sample": f"{len(runs)} x {runs[0]['frames']} frames of the same workload after {args.warmup} warm-up frames, "
                                  f"from an empty map ({spent:.1f} s of CPU work in total), "
                                  + ("the reference's own sources (oracle/_ref)" if cpu_impl()[1] == "reference" else "oracle port")
                                  + ", reference parallel_for policy (hardware_concurrency-2 threads, >=1000 chunks per group)",
                        "voxel_updates_per_s": vps}

    line = {"metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
            "voxel_updates_per_s": vox / (dev_ms * 1e-3), "chunks_per_frame": chunks / args.steps,
            "live_chunks": live_chunks, "map_chunks": map_chunks, "map_hash": f"{mh:016x}", "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e,
            "gpu_launches": int(launches), "clocks": clocks, "data_gen_s": gen_s,
            "stage_us_per_frame": stage_us}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
