"""Generates tests/golden/pre_room.npz: outputs of the reference's OWN pre-processing loops
(BasicAPI.cpp, compiled as oracle/_ref/libtexfusion_ref_pre.so) on seeded synthetic frames.
Run from the repo root, on an Intel host (RSQRTPS!): `python tests/golden/make_pre_golden.py`.

Chain (main.cpp:117-147, without the tracker between the steps), at 160x120:
  key-frame K, later frame N, both with 2 mm depth noise
  normal_map(K), normal_map(N)
  refine_keyframe(K <- N) twice (weights 0 -> 1 -> 2), refine_newframe(N | K)
  refine_depth_by_normal(N)
  color_valid(K), color_quality(K)
Stored in full (float32 / uint8 planes, compressed).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))


def inputs():
    from texturefusion_b200 import synth
    cam = synth.Camera().scaled(0.25)
    seq = synth.make_sequence(7, cam=cam, total=300, keyframe_every=6)
    rng = np.random.default_rng(11)
    k, n = seq.frames[0], seq.frames[6]
    dk, dn = k.depth.copy(), n.depth.copy()
    for d in (dk, dn):
        d[d > 0] += rng.normal(0, 0.002, size=int((d > 0).sum())).astype(np.float32)
    camf = (np.float32(cam.fx), np.float32(cam.fy), np.float32(cam.cx), np.float32(cam.cy))
    return camf, k.pose, dk.astype(np.float32), k.rgb, n.pose, dn.astype(np.float32)


def run_chain(P, camf, pose_k, dk, rgb_k, pose_n, dn):
    """P: oracle.pre.Pre, or anything with the same methods (the GPU adapter of tests/test_pre_gpu.py)."""
    from oracle.pre import relative_transform
    out = {}
    out["normal_k0"] = P.normal_map(dk, camf)
    out["normal_n"] = P.normal_map(dn, camf)
    T_kn = relative_transform(pose_n, pose_k)   # key-frame camera -> new camera
    T_nk = relative_transform(pose_k, pose_n)
    w = np.zeros_like(dk)
    d1, w1 = P.refine_keyframe(dk, w, dn, T_kn, camf)
    d2, w2 = P.refine_keyframe(d1, w1, dn, T_kn, camf)
    out["kf_depth_1"], out["kf_weight_1"], out["kf_depth_2"], out["kf_weight_2"] = d1, w1, d2, w2
    out["new_depth"] = P.refine_newframe(d2, dn, T_nk, camf)
    out["normal_n_refined"], out["new_depth_refined"] = P.refine_depth_by_normal(out["normal_n"], out["new_depth"], camf)
    nk = P.normal_map(d2, camf)
    out["normal_k2"] = nk
    out["color_valid"] = P.color_valid(nk, camf)
    out["quality"] = P.color_quality(d2, nk, rgb_k, camf)
    return out


def main():
    from oracle import pre
    R = pre.Pre("ref")
    assert R.host_rsqrt_matches(), "generate on a host whose RSQRTPS is the Intel table"
    out = run_chain(R, *inputs())
    np.savez_compressed(os.path.join(HERE, "pre_room.npz"), **out)
    print({k: (v.shape, str(v.dtype)) for k, v in out.items()})
    print("refined key-frame pixels:", int((out["kf_weight_2"] > 0).sum()), " rejected new pixels:",
          int(((out["new_depth"] == 0)).sum()), " valid colour pixels:", int(out["color_valid"].sum()))


if __name__ == "__main__":
    main()
