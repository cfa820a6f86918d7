"""Device timeline of the fused per-frame chain (needs a -DTF_TIMELINE build of the library):

  python -m texturefusion_b200.build timeline          # -> build/variants/timeline.so
  TEXFUSION_B200_LIB=build/variants/timeline.so python tools/timeline.py [--steps 100] [--lists] [--res 0.04]

Prints, per kernel, the mean (over frames) of: first block start, first / … return from the
programmatic-dependent-launch wait, end of the last block's main work and end of its tail, in
microseconds relative to the first block of bbox_kernel, plus the host-side wall time of the call."""
import argparse
import ctypes as C
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
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--no-flush", action="store_true")
    ap.add_argument("--lists", action="store_true", help="request the ordered chunk lists (as the e2e leg of bench.py does)")
    ap.add_argument("--flush-read", action="store_true", help="after the write flush, read a second buffer (clean L2 lines)")
    args = ap.parse_args()
    cam = synth.Camera()
    n = args.steps + args.warmup
    seq = synth.make_sequence(n, cam=cam, total=300, keyframe_every=10, device="cuda")
    m = capi.Map(args.res, max_frames=n + 4)
    L = m.L
    L.tf_debug_timeline.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.c_int]
    flush = torch.empty(128 * 1024 * 1024, dtype=torch.float32, device="cuda")
    flush2 = torch.zeros(128 * 1024 * 1024, dtype=torch.float32, device="cuda")
    for fr in seq.frames:
        m.upload_frame(fr.index, fr.depth, fr.rgba() if fr.is_keyframe else None, fr.quality if fr.is_keyframe else None)
    camc = capi.make_camera(cam)
    st = capi.FrameStats()
    out = (C.c_uint64 * 32)()
    L.tf_debug_trace.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
    trace = (C.c_uint64 * 1024)()
    host = (C.c_double * 16)()
    hosts = []
    traces2 = []
    traces = []
    rows = []
    o_ids = np.empty((m.list_cap, 3), np.int32)
    o_new = np.empty(m.list_cap, np.uint8)
    o_upd = np.empty(m.list_cap, np.uint8)
    o_q = np.empty(m.list_cap, np.float32)
    cnt = []
    walls = []
    for i, fr in enumerate(seq.frames):
        pose = capi.make_pose(fr.pose)
        if not args.no_flush:
            flush.fill_(1.0)
            if args.flush_read:
                sink = flush2.sum()
        torch.cuda.synchronize()
        L.tf_debug_timeline(m.h, None, 1)
        t0 = time.perf_counter()
        if args.lists:
            vp = C.c_void_p
            rc = L.tf_integrate_frame(m.h, fr.index, int(fr.is_keyframe), C.byref(pose), C.byref(camc), C.byref(st),
                                      o_ids.ctypes.data_as(vp), o_new.ctypes.data_as(vp), o_upd.ctypes.data_as(vp),
                                      o_q.ctypes.data_as(vp), m.list_cap)
        else:
            rc = L.tf_integrate_frame(m.h, fr.index, int(fr.is_keyframe), C.byref(pose), C.byref(camc), C.byref(st),
                                      None, None, None, None, 0)
        t1 = time.perf_counter()
        assert rc == 0
        L.tf_debug_timeline(m.h, out, 0)
        if i < args.warmup:
            continue
        L.tf_debug_host(host)
        hosts.append([host[k] - host[0] for k in range(7)] + [(t1 - t0) * 1e6])
        L.tf_debug_trace(m.h, trace)
        t_zero = float((~int(out[0])) & 0xFFFFFFFFFFFFFFFF)
        tr = np.array(trace[:512], dtype=np.float64).reshape(16, 32)
        tr[tr == 0] = np.nan
        traces.append((tr - t_zero) / 1e3)
        tr = np.array(trace[512:], dtype=np.float64).reshape(16, 32)
        tr[tr == 0] = np.nan
        traces2.append((tr - t_zero) / 1e3)
        a = np.array(out[:24], dtype=np.uint64).reshape(6, 4)
        cnt.append([int(out[28]), int(out[29]), int(out[30]), int(out[31])])
        t = np.empty((6, 4))
        for k in range(6):
            for j in range(4):
                v = int(a[k, j])
                if v == 0:
                    t[k, j] = np.nan
                else:
                    t[k, j] = float((~v & 0xFFFFFFFFFFFFFFFF) if (j < 2 and k < 4) else v)
        t -= t[0, 0]
        rows.append(t / 1e3)
        walls.append((t1 - t0) * 1e6)
    r = np.nanmean(np.stack(rows), axis=0)
    names = ["bbox", "cull", "export", "integrate"]
    print(f"{'kernel':10s} {'start':>8s} {'waited':>8s} {'work end':>9s} {'end':>8s}   (us after bbox start, mean of {len(rows)} frames)")
    for k in range(4):
        print(f"{names[k]:10s} {r[k,0]:8.2f} {r[k,1]:8.2f} {r[k,2]:9.2f} {r[k,3]:8.2f}")
    print(f"cull internals (latest block): grid known {r[4,0]:.2f}, round-1 coarse done {r[4,1]:.2f}, round-1 fine done {r[4,2]:.2f}")
    print(f"integrate internals (latest block, warp 0): list length known {r[5,0]:.2f}, first chunk arrived {r[5,1]:.2f}, first chunk done {r[5,2]:.2f}")
    c = np.mean(np.array(cnt, dtype=np.float64), axis=0)
    print(f"mean per frame: coarse hits {c[0]:.0f}, coarse candidates {c[1]:.0f}, hit candidates {c[2]:.0f}, list {c[3]:.0f}")
    tr = np.nanmean(np.stack(traces2), axis=0)
    print("cull per-warp trace (warp 0 of blocks 0,37,...)")
    for k, nm in enumerate(["waited", "grid known", "coarse tested", "coarse sync", "fine tested", "chunks resolved",
                            "list pos known", "task done", "round done", "bbox words loaded", "block synced"]):
        print(f"  {nm:18s} " + " ".join(f"{tr[b, k]:6.1f}" for b in range(0, 16)))
    tr = np.nanmean(np.stack(traces), axis=0)
    names_t = ["kernel n known", "loop top", "bulk issued", "p0 projected", "p0 gathered", "p0 chunk arrived", "p0 updated",
               "p1 projected", "p1 gathered", "p1 (arrived)", "p1 updated", "frames done", "written back", "ord shuffled",
               "finalized", "while top"]
    print("per-warp trace (warp 0 of blocks 0,37,...; us after bbox start; mean over frames); chunk 0 | chunk 1")
    for k, nm in enumerate(names_t):
        a = " ".join(f"{tr[b, k]:6.1f}" for b in range(0, 16, 3))
        bb = " ".join(f"{tr[b, 16 + k]:6.1f}" for b in range(0, 16, 3))
        print(f"  {nm:18s} {a}  | {bb}")
    hm = np.median(np.array(hosts), axis=0)
    print("host (us after entry, median): params ready %.1f | bbox launch %.1f..%.1f | cull launched %.1f | integrate launched %.1f | sync returned %.1f | call %.1f"
          % (hm[1], hm[2], hm[3], hm[4], hm[5], hm[6], hm[7]))
    print(f"host call wall: mean {np.mean(walls):.1f} us, median {np.median(walls):.1f} us")
    m.close()


if __name__ == "__main__":
    main()
