"""Generates tests/golden/*.npz from the CPU oracle (run from the repo root:
`python tests/golden/make_golden.py`).  The reference ships no golden vectors and cannot be
built here, so these pin the ORACLE's behaviour at the commit they were made;
tests/test_oracle_cpu.py re-checks the oracle against them on every run and
tests/test_golden_gpu.py checks the CUDA path against them without running the oracle.

Each case: a quarter-resolution (160x120) view of the seeded synthetic room, fused with the
ReIntegrateKeyframe protocol (GCFusion/MobileFusion.cpp:114-221): Prepare on the key-frame,
key-frame with colour + quality, two local depth frames into the same list, Finalize; then
the key-frame is de-integrated and re-integrated under a corrected pose.
Coarse cases store the full voxel dumps, fine cases store SHA-256 digests.
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import OracleMap  # noqa: E402
from texturefusion_b200 import synth  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = {"res40mm": (0.04, True), "res20mm": (0.02, True), "res10mm": (0.01, False), "res5mm": (0.005, False)}


def digest(*arrays) -> str:
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def sorted_dump(m):
    ids = m.list_chunks()
    ids = ids[np.lexsort((ids[:, 2], ids[:, 1], ids[:, 0]))]
    sdf, w, col = m.download_chunks(ids)
    return ids, sdf, w, col


def run_protocol(make_map, cam, frames, new_pose):
    """Drives any map exposing the oracle-style calls; returns the observable outputs."""
    m = make_map()
    kf, local = frames[0], frames[1:]
    out = {}
    ids, new = m.prepare(kf.depth, kf.pose, cam)
    nu, q = m.integrate(kf.depth, kf.rgba(), kf.quality, kf.pose, cam, ids, 1, kf.index)
    for lf in local:
        nu, _ = m.integrate(lf.depth, None, None, lf.pose, cam, ids, 1, -1, nu)
    valid = m.finalize(ids, nu, new)
    out.update(ids0=ids, new0=new, nu0=nu, q0=q, valid0=valid)
    out["stage1"] = sorted_dump(m)
    # loop closure: de-integrate with the old pose over validChunks, re-integrate with the new one
    nu1 = np.ones(len(valid), np.uint8)
    nu1, _ = m.integrate(kf.depth, kf.rgba(), kf.quality, kf.pose, cam, valid, 0, kf.index, nu1)
    for lf in local:
        nu1, _ = m.integrate(lf.depth, None, None, lf.pose, cam, valid, 0, -1, nu1)
    m.finalize(valid, nu1, np.zeros(len(valid), np.uint8))
    out["stage2"] = sorted_dump(m)
    ids2, new2 = m.prepare(kf.depth, new_pose, cam)
    nu2, q2 = m.integrate(kf.depth, kf.rgba(), kf.quality, new_pose, cam, ids2, 1, kf.index)
    for lf in local:
        nu2, _ = m.integrate(lf.depth, None, None, lf.pose, cam, ids2, 1, -1, nu2)
    valid2 = m.finalize(ids2, nu2, new2)
    out.update(ids2=ids2, new2=new2, nu2=nu2, q2=q2, valid2=valid2)
    out["stage3"] = sorted_dump(m)
    return out


def case_inputs():
    cam = synth.Camera().scaled(0.25)
    seq = synth.make_sequence(3, cam=cam, total=300, keyframe_every=3, start=21, noise_sigma=0.001)
    new_pose = seq.frames[0].pose.copy()
    new_pose[:3, 3] += np.array([0.006, -0.004, 0.003], np.float32)
    return cam, seq.frames, new_pose


def pack(out, full: bool) -> dict:
    d = {k: out[k] for k in ("ids0", "new0", "nu0", "q0", "valid0", "ids2", "new2", "nu2", "q2", "valid2")}
    for st in ("stage1", "stage2", "stage3"):
        ids, sdf, w, col = out[st]
        d[f"{st}_ids"] = ids
        d[f"{st}_digest"] = np.frombuffer(bytes.fromhex(digest(sdf, w, col)), np.uint8)
        d[f"{st}_stats"] = np.array([int((w > 0).sum()), int((col.reshape(-1, 4)[:, 3] > 0).sum())], np.int64)
        if full:
            d[f"{st}_sdf"], d[f"{st}_w"], d[f"{st}_col"] = sdf, w, col
    return d


def main():
    cam, frames, new_pose = case_inputs()
    inputs = {"cam": np.array([cam.width, cam.height, cam.fx, cam.fy, cam.cx, cam.cy, cam.near, cam.far], np.float64),
              "new_pose": new_pose}
    for i, fr in enumerate(frames):
        inputs[f"depth{i}"], inputs[f"pose{i}"], inputs[f"index{i}"] = fr.depth, fr.pose, np.int32(fr.index)
    inputs["rgb0"], inputs["valid0"], inputs["quality0"] = frames[0].rgb, frames[0].color_valid, frames[0].quality
    np.savez_compressed(os.path.join(HERE, "inputs.npz"), **inputs)
    for name, (res, full) in CASES.items():
        out = run_protocol(lambda: OracleMap(res), cam, frames, new_pose)
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **pack(out, full))
        print(name, "chunks per stage:", [len(out[s][0]) for s in ("stage1", "stage2", "stage3")],
              "list", len(out["ids0"]), len(out["ids2"]))


if __name__ == "__main__":
    main()
