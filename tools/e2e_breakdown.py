"""Where does the end-to-end frame time go?  Times, per frame: the H2D upload alone (stream
synchronised), the fused call with / without the ordered lists, and both back to back as bench.py's
e2e leg does.  Usage (GPU box): python tools/e2e_breakdown.py [--steps 100]"""
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
    ap.add_argument("--warmup", type=int, default=10)
    args = ap.parse_args()
    cam = synth.Camera()
    n = args.steps + args.warmup
    seq = synth.make_sequence(n, cam=cam, total=300, keyframe_every=10, device="cuda")
    L = capi.load()
    vp = C.c_void_p
    camc = capi.make_camera(cam)
    st = capi.FrameStats()
    flush = torch.empty(128 * 1024 * 1024, dtype=torch.float32, device="cuda")
    npix = cam.width * cam.height
    pins = []
    for fr in seq.frames:
        d = capi.PinnedBuffer((npix,), np.float32)
        d.array[:] = np.asarray(fr.depth, np.float32).ravel()
        c = q = None
        if fr.is_keyframe:
            c = capi.PinnedBuffer((npix * 4,), np.uint8)
            c.array[:] = np.asarray(fr.rgba(), np.uint8).ravel()
            q = capi.PinnedBuffer((npix,), np.float32)
            q.array[:] = np.asarray(fr.quality, np.float32).ravel()
        pins.append((d, c, q))
    for mode in ("split: upload | fused call with lists", "split: upload | fused call, no lists", "back to back with lists"):
        m = capi.Map(args.res, max_frames=n + 4)
        cap = m.list_cap
        ids = np.empty((cap, 3), np.int32)
        new = np.empty(cap, np.uint8)
        upd = np.empty(cap, np.uint8)
        q_ = np.empty(cap, np.float32)
        t_up = {False: [], True: []}
        t_in = {False: [], True: []}
        for i, fr in enumerate(seq.frames):
            pose = capi.make_pose(fr.pose)
            d, c, q = pins[i]
            flush.fill_(1.0)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            rc = L.tf_upload_frame(m.h, fr.index, vp(d.ptr), vp(c.ptr) if c else None, vp(q.ptr) if q else None)
            assert rc == 0
            if not mode.startswith("back"):
                L.tf_sync(m.h)
            t1 = time.perf_counter()
            if "no lists" in mode:
                rc = L.tf_integrate_frame(m.h, fr.index, int(fr.is_keyframe), C.byref(pose), C.byref(camc), C.byref(st),
                                          None, None, None, None, 0)
            else:
                rc = L.tf_integrate_frame(m.h, fr.index, int(fr.is_keyframe), C.byref(pose), C.byref(camc), C.byref(st),
                                          ids.ctypes.data_as(vp), new.ctypes.data_as(vp), upd.ctypes.data_as(vp),
                                          q_.ctypes.data_as(vp), cap)
            t2 = time.perf_counter()
            assert rc == 0
            if i >= args.warmup:
                t_up[fr.is_keyframe].append((t1 - t0) * 1e6)
                t_in[fr.is_keyframe].append((t2 - t1) * 1e6)
        print(mode)
        for kf in (False, True):
            print(f"   {'key-frames ' if kf else 'depth-only '} upload {np.median(t_up[kf]):7.1f} us   fused call {np.median(t_in[kf]):7.1f} us"
                  f"   (n={len(t_up[kf])})")
        tot = sum(t_up[False]) + sum(t_up[True]) + sum(t_in[False]) + sum(t_in[True])
        print(f"   mean per frame {tot / args.steps:7.1f} us -> {args.steps / tot * 1e6:7.0f} frames/s")
        m.close()


if __name__ == "__main__":
    main()
