"""Meshing: ChunkManager::GenerateMeshEfficient for every chunk of a fused room
on the device (tf_mesh_chunks) next to the reference's own CPU mesher (oracle/_ref, when present).

  python tools/bench_mesh.py [--frames 60] [--res 0.005] [--cpu]
Prints one JSON line: chunks, vertices, triangles, device ms per call (two kernel passes + the mesh
download), chunks/s, and — with --cpu — the reference mesher's time on the host cores and whether the
two meshes are bit-identical."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from texturefusion_b200 import capi, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=60)
    ap.add_argument("--res", type=float, default=0.005)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--cpu", action="store_true")
    args = ap.parse_args()
    cam = synth.Camera()
    seq = synth.make_sequence(args.frames, cam=cam, total=300, keyframe_every=10, device="cuda")
    m = capi.Map(args.res, max_frames=8)
    o = None
    if args.cpu:
        from oracle import OracleMap
        o = OracleMap(args.res, impl="ref", threads=0)
    for rep in range(3):  # three passes over the frames: weights above the mesher's threshold of 50
        for fr in seq.frames:
            rg = fr.rgba() if fr.is_keyframe else None
            q = fr.quality if fr.is_keyframe else None
            m.upload_frame(fr.index, fr.depth, rg, q)
            m.integrate_frame(fr.index, fr.is_keyframe, fr.pose, cam, want_lists=False)
            if o is not None:
                o.integrate_frame(fr.depth, rg, q, fr.pose, cam, fr.index if fr.is_keyframe else -1)
    ids = m.list_chunks()
    ids = ids[np.lexsort((ids[:, 2], ids[:, 1], ids[:, 0]))]
    out = m.mesh_chunks(ids)
    m.sync()
    t0 = time.perf_counter()
    for _ in range(args.reps):
        out = m.mesh_chunks(ids)
    dt = (time.perf_counter() - t0) / args.reps
    line = {"metric": "marching-cubes meshing of a fused room (tf_mesh_chunks, size query + mesh: 3 kernel passes + download)",
            "res": args.res, "chunks": int(len(ids)), "vertices": int(out[0][-1]), "triangles": int(out[1][-1]) // 3,
            "ms_per_call": dt * 1e3, "chunks_per_s": len(ids) / dt, "d2h_MB": (out[0][-1] * 36 + out[1][-1] * 4) / 1e6,
            "voxel_planes_MB_not_downloaded": len(ids) * 8192 / 1e6}
    if o is not None:
        t0 = time.perf_counter()
        ref = o.mesh_chunks(ids)
        line["cpu_reference_ms"] = (time.perf_counter() - t0) * 1e3 / 2  # (mesh_chunks runs the mesher twice: sizes, then data)
        line["bit_identical"] = all(np.array_equal(np.asarray(a).view(np.uint32) if np.asarray(a).dtype == np.float32 else a,
                                                   np.asarray(b).view(np.uint32) if np.asarray(b).dtype == np.float32 else b)
                                    for a, b in zip(out, ref))
    print(json.dumps(line))
    m.close()


if __name__ == "__main__":
    main()
