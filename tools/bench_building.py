"""BASELINE.json configs[3] (C4), single GPU: a long walk through the 40 x 25 x 3 m building at 5 mm
voxels.  Frames are rendered on the fly, uploaded from page-locked host memory and fused one call
at a time (double buffered like bench.py's e2e leg); reports throughput per window of frames next
to the size of the map, i.e. whether the per-frame cost depends on how much has been mapped.

  python tools/bench_building.py [--frames 2000] [--res 0.005] [--max-chunks 1048576]"""
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
    ap.add_argument("--frames", type=int, default=2000)
    ap.add_argument("--total", type=int, default=5000, help="poses of the whole walk (frames are its first --frames)")
    ap.add_argument("--res", type=float, default=0.005)
    ap.add_argument("--max-chunks", type=int, default=1 << 20)
    ap.add_argument("--window", type=int, default=250)
    args = ap.parse_args()
    cam = synth.Camera()
    m = capi.Map(args.res, max_frames=8, max_chunks=args.max_chunks)
    L = m.L
    vp = C.c_void_p
    camc = capi.make_camera(cam)
    st = capi.FrameStats()
    npix = cam.width * cam.height
    pins = [(capi.PinnedBuffer((npix,), np.float32), capi.PinnedBuffer((npix * 4,), np.uint8), capi.PinnedBuffer((npix,), np.float32))
            for _ in range(2)]

    def render(k):
        pose = synth.walk_pose(k, args.total)
        kf = k % 10 == 0
        depth, rgb, q = synth.render(pose, cam, color=kf, device="cuda", scene="building")
        fr = synth.Frame(k, pose, depth, rgb, np.ones(depth.shape, np.uint8) if kf else None, q, kf)
        d, c, qq = pins[k & 1]
        d.array[:] = np.asarray(fr.depth, np.float32).ravel()
        if kf:
            c.array[:] = np.asarray(fr.rgba(), np.uint8).ravel()
            qq.array[:] = np.asarray(fr.quality, np.float32).ravel()
        return fr

    def upload(fr):
        d, c, qq = pins[fr.index & 1]
        rc = L.tf_upload_frame(m.h, fr.index, vp(d.ptr), vp(c.ptr) if fr.is_keyframe else None, vp(qq.ptr) if fr.is_keyframe else None)
        assert rc == 0, L.tf_last_error(m.h)

    rows = []
    cur = render(0)
    upload(cur)
    t_win, n_win, vox_win = 0.0, 0, 0
    for k in range(args.frames):
        nxt = render(k + 1) if k + 1 < args.frames else None  # (rendering is not timed)
        pose = capi.make_pose(cur.pose)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if nxt is not None:
            upload(nxt)
        rc = L.tf_integrate_frame(m.h, cur.index, int(cur.is_keyframe), C.byref(pose), C.byref(camc), C.byref(st), None, None,
                                  None, None, 0)
        assert rc == 0, L.tf_last_error(m.h)
        if nxt is not None:
            assert L.tf_wait_upload(m.h, nxt.index) == 0
        t_win += time.perf_counter() - t0
        n_win += 1
        vox_win += st.voxel_updates
        if n_win == args.window or k + 1 == args.frames:
            rows.append({"frames": k + 1, "fps_e2e": n_win / t_win, "chunks_per_frame": vox_win / 512 / n_win,
                         "live_chunks": m.chunk_count(), "map_GB": m.chunk_count() * 8192 / 1e9})
            print(json.dumps(rows[-1]), flush=True)
            t_win, n_win, vox_win = 0.0, 0, 0
        cur = nxt
    m.close()


if __name__ == "__main__":
    main()
