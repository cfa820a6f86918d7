"""Texture-atlas update (Structure/Atlas.cpp:71-91): N patches of one key-frame are
copied (crop fits the slot) or shrunk with the fixed-point bilinear resize (crop larger than the
slot) into their atlas slots by one tf_atlas_update call.  Reports patches/s and the algorithmic
bytes (crop read + slot written, 3 B per texel) per second, next to the CPU oracle.

  python tools/bench_atlas.py [--patches 4096] [--res 0.005]"""
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
    ap.add_argument("--patches", type=int, default=4096)
    ap.add_argument("--res", type=float, default=0.005)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--cpu", action="store_true", help="also time the CPU oracle (test infrastructure) on the same patches")
    args = ap.parse_args()
    cam = synth.Camera()
    kf = synth.make_sequence(1, cam=cam, total=300, keyframe_every=1, start=50).frames[0]
    m = capi.Map(args.res)
    m.upload_frame(kf.index, kf.depth)
    m.upload_keyframe_rgb(kf.index, kf.rgb)
    pw, ph = m.atlas_patch_size()
    rng = np.random.RandomState(0)
    out = {}
    for kind in ("copy", "resize"):
        patches, bytes_alg = [], 0
        for i in range(args.patches):
            if kind == "copy":
                w, h = int(rng.randint(pw // 2, pw + 1)), int(rng.randint(ph // 2, ph + 1))
            else:
                w, h = int(rng.randint(pw + 1, 3 * pw)), int(rng.randint(ph + 1, 3 * ph))
            x, y = int(rng.randint(0, cam.width - w)), int(rng.randint(0, cam.height - h))
            loc = m.atlas_alloc_slot((i, 0, 0 if kind == "copy" else 1))
            patches.append((loc, kf.index, x, y, w, h))
            bytes_alg += 3 * (w * h + (pw * ph if kind == "resize" else w * h))
        arr = (capi.PatchDesc * len(patches))()
        for i, p in enumerate(patches):
            arr[i] = capi.PatchDesc(*[int(v) for v in p])
        for _ in range(3):
            assert m.L.tf_atlas_update(m.h, arr, len(patches)) == 0
        m.sync()
        t0 = time.perf_counter()
        for _ in range(args.reps):
            assert m.L.tf_atlas_update(m.h, arr, len(patches)) == 0
        m.sync()
        dt = (time.perf_counter() - t0) / args.reps
        out[kind] = {"patches": len(patches), "us_per_call": dt * 1e6, "patches_per_s": len(patches) / dt,
                     "algorithmic_GBps": bytes_alg / dt / 1e9}
        if args.cpu:
            from oracle import OracleMap
            o = OracleMap(args.res)
            locs = [o.atlas_alloc_slot((i, 0, 0)) for i in range(len(patches))]
            t0 = time.perf_counter()
            for loc, (_, _, x, y, w, h) in zip(locs, patches):
                o.atlas_update(loc, kf.rgb, (x, y, w, h))
            out[kind]["cpu_oracle_patches_per_s"] = len(patches) / (time.perf_counter() - t0)
    print(json.dumps({"metric": "atlas update", "slot": [pw, ph], "res": args.res, **out}))
    m.close()


if __name__ == "__main__":
    main()
