"""Generates tests/golden/mesh_*.npz: meshes of ChunkManager::GenerateMeshEfficient
(Structure/ChunkManager.cpp:595-1002) as computed by the REFERENCE'S OWN SOURCES (oracle/_ref) on maps
fused from tests/golden/inputs.npz.  Run from the repo root: `python tests/golden/make_mesh_golden.py`.

Protocol: the three frames of inputs.npz are fused one by one (IntegrateFrame shape: Prepare +
Integrate + Finalize, the key-frame with colour), 12 times over with small pose offsets so that voxel
weights pass the mesher's threshold of 50 (Structure/ChunkManager.cpp:793); then every chunk of
meshesToUpdate is meshed, in lexicographic id order.  Stored: the id list, the per-chunk vertex / index
offsets and SHA-256 digests of the vertex, normal, colour and index arrays (20 mm also in full).
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = {"mesh_20mm": (0.02, True), "mesh_5mm": (0.005, False)}
REPEATS = 12


def fuse(m, cam, frames):
    """m: anything with the oracle-style integrate_frame (OracleMap or the GPU adapter of tests/test_mesh.py)."""
    for rep in range(REPEATS):
        for fr in frames:
            pose = fr.pose.copy()
            pose[:3, 3] += np.float32(0.0007 * rep)
            m.integrate_frame(fr.depth, fr.rgba() if fr.is_keyframe else None, fr.quality if fr.is_keyframe else None, pose, cam,
                              fr.index if fr.is_keyframe else -1)


def mesh_ids(m):
    ids = m.meshes_to_update()
    return ids[np.lexsort((ids[:, 2], ids[:, 1], ids[:, 0]))]


def digest(a) -> np.ndarray:
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), np.uint8)


def main():
    from oracle import OracleMap
    from test_oracle_cpu import load_inputs
    cam, frames, _ = load_inputs()
    for name, (res, full) in CASES.items():
        o = OracleMap(res, impl="ref")
        fuse(o, cam, frames)
        ids = mesh_ids(o)
        voff, ioff, vert, norm, col, idx = o.mesh_chunks(ids)
        d = {"ids": ids, "vert_off": voff, "idx_off": ioff, "vert_digest": digest(vert), "norm_digest": digest(norm),
             "col_digest": digest(col), "idx_digest": digest(idx), "generator": np.frombuffer(o.L.tfo_impl(), np.uint8)}
        if full:
            d.update(vert=vert, norm=norm, col=col, idx=idx)
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **d)
        print(name, len(ids), "chunks,", int(voff[-1]), "vertices,", int(ioff[-1]) // 3, "triangles, NaN normals:",
              int(np.isnan(norm).any(axis=1).sum()))


if __name__ == "__main__":
    main()
