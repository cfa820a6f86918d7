"""Micro-benchmark of integrate_kernel alone: time vs chunk count on a fixed map state.
Usage (GPU box): python tools/microbench_integrate.py [--res 0.005] [--frames 24]"""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from texturefusion_b200 import capi, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--res", type=float, default=0.005)
    ap.add_argument("--frames", type=int, default=24)
    ap.add_argument("--reps", type=int, default=20)
    args = ap.parse_args()
    cam = synth.Camera()
    seq = synth.make_sequence(args.frames, cam=cam, total=300, keyframe_every=10, device="cuda")
    m = capi.Map(args.res, max_frames=args.frames + 4)
    lists = {}
    for fr in seq.frames:
        m.upload_frame(fr.index, fr.depth, fr.rgba() if fr.is_keyframe else None, fr.quality if fr.is_keyframe else None)
        st, ids, new, upd, q = m.integrate_frame(fr.index, fr.is_keyframe, fr.pose, cam)
        lists[fr.index] = ids[upd != 0]
    print("live chunks", m.chunk_count())
    flush = torch.empty(128 * 1024 * 1024, dtype=torch.float32, device="cuda")
    fr = seq.frames[-1]
    kf = seq.frames[(len(seq.frames) - 1) // 10 * 10]
    for name, frame, color in (("depth-only", fr, False), ("colour", kf, True)):
        full = lists[frame.index]
        print(f"== {name}: frame {frame.index}, list {len(full)} chunks (all exist, all updated)")
        for n in (256, 1024, 2048, 4096, 8192, len(full)):
            n = min(n, len(full))
            ids = full[:: max(1, len(full) // n)][:n]
            for do_flush in (True, False):
                m.set_profiling(1)
                m.kernel_time(reset=True)
                t_wall = 0.0
                for _ in range(args.reps):
                    if do_flush:
                        flush.fill_(1.0)
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    m.integrate(frame.index, color, frame.pose, cam, ids, 1)
                    t_wall += time.perf_counter() - t0
                ms, k, b = m.kernel_time(reset=True)
                us = 1e3 * ms / k
                gbs = b / k / (us * 1e-6) / 1e9
                print(f"  n={len(ids):6d} flush={int(do_flush)} kernel {us:7.2f} us  ({us * 1e3 / len(ids):6.2f} ns/chunk, "
                      f"{gbs:7.1f} GB/s algorithmic)  call wall {1e6 * t_wall / args.reps:7.1f} us")
    m.close()


if __name__ == "__main__":
    main()
