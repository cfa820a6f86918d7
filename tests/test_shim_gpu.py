"""The C++ drop-in shim (texturefusion_b200/host/chisel_b200.h).

  * compile checks (no GPU): the header alone (Eigen-free mode), the round-1 driver, and
    tests/cpp/mobilefusion_excerpt.cpp — MobileFusion's ReIntegrateKeyframe / IntegrateFrame /
    RetractObservations / tsdfFusion tail written against the shim exactly like the reference's call
    sites, with -DTF_WITH_EIGEN against the Eigen stand-in;
  * GPU: that excerpt run on a multi-key-frame sequence (fusion, meshing, patches, atlas, GL buffers,
    loop closure, IntegrateFrame) and every observable compared with the CPU oracle / the reference
    mesher; the shim's per-key-frame time against the same calls on the raw C ABI.
"""
import os
import re
import struct
import subprocess

import numpy as np
import pytest

from oracle import OracleMap, have_ref, patch_texcoords
from texturefusion_b200 import synth
from texturefusion_b200.build import LIB_PATH

from util import room_sequence

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INC = ["-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "texturefusion_b200", "host")]
LINK = [LIB_PATH, f"-Wl,-rpath,{os.path.dirname(LIB_PATH)}"]


def build_driver(tmp) -> str:
    exe = os.path.join(tmp, "shim_driver")
    subprocess.check_call(["g++", "-O2", "-std=c++14", *INC, os.path.join(ROOT, "tests", "cpp", "shim_driver.cpp"), "-o", exe, *LINK])
    return exe


def build_excerpt(tmp) -> str:
    exe = os.path.join(tmp, "mobilefusion_excerpt")
    subprocess.check_call(["g++", "-O2", "-std=gnu++14", "-DTF_WITH_EIGEN", "-I", os.path.join(ROOT, "oracle", "eigen_standin"), *INC,
                           os.path.join(ROOT, "tests", "cpp", "mobilefusion_excerpt.cpp"), "-o", exe, *LINK])
    return exe


def wsum(words: np.ndarray) -> int:
    """position-weighted checksum over 32-bit words: sum(word * index) mod 2^64 (index from 1)."""
    w = np.ascontiguousarray(words).view(np.uint32).astype(np.uint64).ravel()
    with np.errstate(over="ignore"):
        return int((w * np.arange(1, len(w) + 1, dtype=np.uint64)).sum(dtype=np.uint64))


def test_shim_header_compiles_standalone(tmp_path):
    """No GPU needed: the header is self-contained C++14 (Eigen-free mode)."""
    src = tmp_path / "t.cpp"
    src.write_text('#include "chisel_b200.h"\nvoid chisel::Chisel::CompensateColor() {}\n'
                   'int main(){ chisel::PinholeCamera c; return c.GetCx() == 319 ? 0 : 1; }\n')
    exe = tmp_path / "t"
    subprocess.check_call(["g++", "-std=c++14", *INC, str(src), "-o", str(exe), *LINK])
    assert subprocess.call([str(exe)]) == 0  # int-returning getter: cx 319.5 -> 319 (PinholeCamera.h:46-49)


def test_mobilefusion_excerpt_compiles_against_the_shim(tmp_path):
    """MobileFusion's call sites (ReIntegrateKeyframe, IntegrateFrame, RetractObservations, the tsdfFusion tail with
    GetCentroids / UpdateMeshes / CompressMeshes / GeneratePatches(UniGraph&, vector<Frame>&) / UpdateAtlas / DrawMeshes /
    atlas.texture_buffer / hot_start / hot_end) compile unchanged against the shim's classes."""
    build_excerpt(str(tmp_path))
    build_driver(str(tmp_path))


# ---- the excerpt on the GPU against the oracle ---------------------------------------------------------

class Reader:
    def __init__(self, path):
        self.b = open(path, "rb").read()
        self.p = 0

    def arr(self, dtype, n):
        a = np.frombuffer(self.b, dtype, n, self.p)
        self.p += a.nbytes
        return a

    def i64(self):
        return int(self.arr(np.int64, 1)[0])


def read_state(r):
    st = {"valid": [], "chunks": {}, "meshes": {}}
    for _ in range(r.i64()):
        st["valid"].append(r.arr(np.int32, 3 * r.i64()).reshape(-1, 3))
    n = r.i64()
    st["chunk_count"] = r.i64()
    for _ in range(n):
        cid = tuple(r.arr(np.int32, 3))
        sdf, w, col = r.arr(np.float32, 512), r.arr(np.float32, 512), r.arr(np.uint16, 2048)
        obs = {}
        for _ in range(r.i64()):
            k = int(r.arr(np.int32, 1)[0])
            obs[k] = r.arr(np.float32, 1)[0]
        st["chunks"][cid] = (sdf, w, col, obs)
    for _ in range(r.i64()):
        cid = tuple(r.arr(np.int32, 3))
        nv, ni = r.i64(), r.i64()
        vnc = r.arr(np.float32, 9 * nv).reshape(nv, 9)
        idx = r.arr(np.int32, ni)
        adj = int(r.arr(np.int32, 1)[0])
        m = {"vert": vnc[:, 0:3], "norm": vnc[:, 3:6], "col": vnc[:, 6:9], "idx": idx, "adj": adj, "patch": None}
        if int(r.arr(np.int32, 1)[0]):
            texloc = r.i64()
            pr = r.arr(np.int32, 6)
            ntc = r.i64()
            tcc = r.arr(np.float32, 5 * ntc).reshape(ntc, 5)
            m["patch"] = {"texloc": texloc, "frameid": int(pr[0]), "box": tuple(int(v) for v in pr[1:5]), "wrong": int(pr[5]),
                          "tc": tcc[:, 0:2], "tcol": tcc[:, 2:5]}
        st["meshes"][cid] = m
    st["hot"] = (r.i64(), r.i64())
    st["hot_bytes"] = r.arr(np.uint8, 3 * max(0, st["hot"][1] - st["hot"][0]))
    st["n_vert"], st["n_idx"] = r.i64(), r.i64()
    return st


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def rigid_inverse(pose):
    """Sophus::SE3d::inverse() of a camera->world pose, then .cast<float>()."""
    P = np.asarray(pose, np.float64)
    T = np.eye(4)
    T[:3, :3] = P[:3, :3].T
    for i in range(3):
        T[i, 3] = -(P[0, i] * P[0, 3] + P[1, i] * P[1, 3] + P[2, i] * P[2, 3])
    return T.astype(np.float32)


class OracleFlow:
    """The same map-thread flow on the CPU: fusion by the oracle, meshes by the reference mesher."""

    def __init__(self, res, cam, frames, group):
        self.o = OracleMap(res, impl="ref" if have_ref() else "port")
        self.atlas = OracleMap(res)  # (slot copy / resize: the restatement's atlas)
        self.res, self.cam, self.frames, self.group = np.float32(res), cam, frames, group
        self.kfs = [k for k in range(len(frames)) if k % group == 0]
        self.valid = {k: np.zeros((0, 3), np.int32) for k in self.kfs}
        self.meshes = {}   # id -> dict(vert, norm, col, idx, simplified, adj)
        self.labels = {}

    def reintegrate(self, k, flag, poses):
        kf, local = self.frames[k], self.frames[k + 1:k + self.group]
        o, cam = self.o, self.cam
        if flag:
            ids, new = o.prepare(kf.depth, poses[k], cam)
            nu = np.zeros(len(ids), np.uint8)
        else:
            ids = self.valid[k]
            new, nu = np.zeros(len(ids), np.uint8), np.ones(len(ids), np.uint8)
        if len(ids):
            nu, _ = o.integrate(kf.depth, kf.rgba(), kf.quality, poses[k], cam, ids, flag, k, nu)
            for j, lf in enumerate(local):
                nu, _ = o.integrate(lf.depth, None, None, poses[k + 1 + j], cam, ids, flag, -1, nu)
        v = o.finalize(ids, nu, new)
        for cid in np.asarray(ids).reshape(-1, 3)[(np.asarray(new) != 0) & (np.asarray(nu) == 0)]:
            self.meshes.pop(tuple(cid), None)  # RemoveChunk drops the mesh too (Structure/ChunkManager.h:155-157)
        self.valid[k] = v if flag else np.zeros((0, 3), np.int32)

    def mesh_and_texture(self):
        o = self.o
        upd = [tuple(i) for i in o.meshes_to_update()]
        in_map = [i for i in upd if o.has_chunk(i)]
        if in_map and have_ref():
            voff, ioff, vert, norm, col, idx = o.mesh_chunks(np.array(in_map, np.int32))
            for n, cid in enumerate(in_map):
                a, b, c, d = voff[n], voff[n + 1], ioff[n], ioff[n + 1]
                m = self.meshes.get(cid)
                if m is None and b == a:
                    continue  # empty meshes are not inserted (Structure/ChunkManager.cpp:261-263)
                self.meshes[cid] = {"vert": vert[a:b], "norm": norm[a:b], "col": col[a:b], "idx": idx[c:d], "simplified": False,
                                    "adj": 0}
        to_update = [i for i in upd if i in self.meshes]
        # CompressMeshes (Structure/Chisel.cpp:112-147)
        for cid in to_update:
            m = self.meshes[cid]
            if m["simplified"]:
                continue
            origin = (np.float32(8) * np.array(cid, np.float32)) * self.res if False else \
                np.array([np.float32(8 * c) * self.res for c in cid], np.float32)
            pos = np.floor((m["vert"] - origin[None, :]) / self.res).astype(np.int64)
            adj = 0
            for ax in range(3):
                if len(pos) and (pos[:, ax] >= 8).any():
                    adj |= 1 << (2 * ax + 1)
                if len(pos) and (pos[:, ax] <= 0).any():
                    adj |= 1 << (2 * ax)
            m["adj"], m["simplified"] = adj, True
        nb = [(-1, 0, 0), (1, 0, 0), (0, -1, 0), (0, 1, 0), (0, 0, -1), (0, 0, 1)]
        for cid in to_update:
            m = self.meshes[cid]
            for k, d in enumerate(nb):
                other = self.meshes.get((cid[0] + d[0], cid[1] + d[1], cid[2] + d[2]))
                if other is None or not other["simplified"]:
                    continue
                mm = k + 1 if k % 2 == 0 else k - 1
                if (m["adj"] >> k) & 1 or (other["adj"] >> mm) & 1:
                    m["adj"] |= 1 << k
                    other["adj"] |= 1 << mm
        # labels: the key-frame with the largest observation quality
        textured = []
        for cid in to_update:
            if not o.has_chunk(cid):
                continue
            best, bq = -1, -1.0
            for k in self.kfs:
                q = o.observation(cid, k)
                if q is not None and q > bq:
                    best, bq = k, q
            if best >= 0:
                self.labels[cid] = best
                textured.append(cid)
        o.clear_meshes_to_update()  # CompressMeshes: chunksToUpdate.clear() (Structure/Chisel.cpp:146)
        return textured


def check_stage(st, flow, textured, corrected, stage):
    o, cam, frames = flow.o, flow.cam, flow.frames
    # validChunks of every key-frame, in order
    for n, k in enumerate(flow.kfs):
        assert np.array_equal(st["valid"][n], flow.valid[k]), f"{stage}: validChunks of key-frame {k}"
    # the map: id set (the shim's host mirror == the device == the oracle), voxels, observations
    oi = o.list_chunks()
    oi = oi[np.lexsort((oi[:, 2], oi[:, 1], oi[:, 0]))]
    assert st["chunk_count"] == len(oi) == len(st["chunks"]), f"{stage}: chunk count"
    assert [tuple(i) for i in oi] == list(st["chunks"].keys()), f"{stage}: id set of the host mirror"
    sdf, w, col = o.download_chunks(oi)
    for n, cid in enumerate(st["chunks"]):
        gs, gw, gc, gobs = st["chunks"][cid]
        assert np.array_equal(bits(gs), bits(sdf[n])) and np.array_equal(bits(gw), bits(w[n])) and np.array_equal(gc, col[n]), \
            f"{stage}: voxels of {cid}"
        want = {k: o.observation(cid, k) for k in flow.kfs}
        want = {k: v for k, v in want.items() if v is not None}
        assert set(gobs) == set(want) and all(np.float32(gobs[k]) == np.float32(want[k]) for k in want), f"{stage}: observations of {cid}"
    if not have_ref():
        return
    # meshes: bit-identical to the reference mesher, adjacency flags of CompressMeshes
    assert set(st["meshes"]) == set(flow.meshes), f"{stage}: set of meshed chunks"
    n_complete = 0
    for cid, gm in st["meshes"].items():
        om = flow.meshes[cid]
        for key in ("vert", "norm", "col"):
            assert np.array_equal(bits(gm[key]), bits(om[key])), f"{stage}: mesh {key} of {cid}"
        assert np.array_equal(gm["idx"], om["idx"]), f"{stage}: mesh indices of {cid}"
        assert gm["adj"] == om["adj"], f"{stage}: adj flags of {cid}: {gm['adj']:06b} vs {om['adj']:06b}"
    # patches of the chunks textured in this stage: texcoords / texcolor / bbox / wrong_mapping per key-frame
    by_frame = {}
    for cid in textured:
        by_frame.setdefault(flow.labels[cid], []).append(cid)
    seen_loc = set()
    for k, cids in by_frame.items():
        off = np.cumsum([0] + [len(flow.meshes[c]["vert"]) for c in cids]).astype(np.int64)
        verts = np.concatenate([flow.meshes[c]["vert"] for c in cids])
        cols = np.concatenate([flow.meshes[c]["col"] for c in cids])
        tc, tcol, res = patch_texcoords(frames[k].rgb, frames[k].depth, rigid_inverse(corrected[k]), cam, off, verts, cols)
        for n, cid in enumerate(cids):
            p = st["meshes"][cid]["patch"]
            if off[n + 1] == off[n]:  # the chunk's mesh became empty: its patch cannot be complete (Patch.cpp:191-196)
                assert p is None
                continue
            assert p is not None, f"{stage}: patch of {cid} incomplete"
            n_complete += 1
            assert p["frameid"] == k and p["box"] == tuple(int(v) for v in res[n][:4]) and p["wrong"] == int(res[n][4]), f"{stage}: patch of {cid}"
            assert np.array_equal(bits(p["tc"]), bits(tc[off[n]:off[n + 1]])), f"{stage}: texcoords of {cid}"
            assert np.array_equal(bits(p["tcol"]), bits(tcol[off[n]:off[n + 1]])), f"{stage}: texcolor of {cid}"
            assert p["texloc"] not in seen_loc
            seen_loc.add(p["texloc"])
            if p["box"][2] > 0 and p["box"][3] > 0:
                flow.atlas.atlas_update(p["texloc"], frames[k].rgb, p["box"])
    assert n_complete > 0, f"{stage}: nothing was textured"
    # atlas: the hot rows of the host mirror == the oracle's atlas; hot range as Structure/Chisel.cpp:184-186
    locs = [st["meshes"][c]["patch"]["texloc"] for c in textured if st["meshes"][c]["patch"] is not None]
    pw, ph = flow.atlas.atlas_patch_size()
    W = 13824
    # (patches of chunks whose mesh became empty also count in the reference's range, but are not dumped)
    assert st["hot"][0] % W == 0 and st["hot"][1] % W == 0 and st["hot"][0] <= (min(locs) // W) * W and \
        st["hot"][1] >= (max(locs) // W + ph) * W, f"{stage}: hot range"
    if all(st["meshes"][c]["patch"] is not None for c in textured):
        assert st["hot"] == ((min(locs) // W) * W, (max(locs) // W + ph) * W), f"{stage}: hot range"
    want = flow.atlas.atlas_download(st["hot"][0], min(st["hot"][1], W * W))
    assert np.array_equal(st["hot_bytes"][:len(want)], want), f"{stage}: atlas hot rows differ from the oracle"
    # GL buffers: every complete patch contributes its mesh
    complete = [m for m in st["meshes"].values() if m["patch"] is not None and len(m["patch"]["tc"]) == len(m["vert"])]
    assert st["n_vert"] == sum(len(m["vert"]) for m in complete) and st["n_idx"] == sum(len(m["idx"]) for m in complete)


def check_exports(folder, st):
    """SaveAllMeshesToPLY / SaveTexturedModel (Structure/Chisel.cpp:357-379, io/PLY.cpp, Structure/Atlas.cpp:93-179):
    file layouts as the reference writes them, contents consistent with the dumped state."""
    n_idx_all = sum(len(m["idx"]) for m in st["meshes"].values())
    ply = open(os.path.join(folder, "model.ply")).read().split("\n")
    assert ply[0] == "ply" and ply[1] == "format ascii 1.0" and ply[2] == f"element vertex {n_idx_all}"
    end = ply.index("end_header")
    assert f"element face {n_idx_all // 3}" in ply[:end] and "property uchar red" in ply[:end]
    body = [ln for ln in ply[end + 1:] if ln.strip()]
    assert len(body) == n_idx_all + n_idx_all // 3
    assert len(body[0].split()) == 6 and body[n_idx_all].split()[0] == "3"
    complete = [m for m in st["meshes"].values() if m["patch"] is not None and len(m["patch"]["tc"]) == len(m["vert"])]
    nv, nf = sum(len(m["vert"]) for m in complete), sum(len(m["idx"]) for m in complete) // 3
    obj = open(os.path.join(folder, "texture_model.obj")).read().split("\n")
    assert obj[0] == "mtllib texture_model.mtl"
    assert sum(ln.startswith("v ") for ln in obj) == nv and sum(ln.startswith("vt ") for ln in obj) == nv
    assert sum(ln.startswith("vn ") for ln in obj) == nv and sum(ln.startswith("f ") for ln in obj) == nf
    vt = np.array([[float(x) for x in ln.split()[1:]] for ln in obj if ln.startswith("vt ")])
    assert len(vt) == 0 or (vt.min() >= -1e-3 and vt.max() <= 1 + 1e-3)
    mtl = open(os.path.join(folder, "texture_model.mtl")).read()
    assert "newmtl demo_texture" in mtl and "map_Kd texture_material.ppm" in mtl
    with open(os.path.join(folder, "texture_material.ppm"), "rb") as f:
        assert f.readline() == b"P6\n"
        w, h = (int(x) for x in f.readline().split())
        assert f.readline() == b"255\n" and w == 13824 and h * w == st["hot"][1]
        img = np.frombuffer(f.read(), np.uint8)
    assert np.array_equal(img[3 * st["hot"][0]:], st["hot_bytes"]), "exported texture differs from the atlas hot rows"


@pytest.mark.gpu
@pytest.mark.parametrize("res,scale", ((0.02, 1.0), (0.005, 0.25)))
def test_mobilefusion_flow_through_the_shim_matches_oracle(tmp_path, res, scale):
    cam = synth.Camera()
    if scale != 1.0:
        cam = cam.scaled(scale)
    group, n_kf = 3, 3
    seq = synth.make_sequence(group * n_kf, cam=cam, total=300, keyframe_every=group, start=12, with_drift=True)
    frames = seq.frames
    for k, fr in enumerate(frames):
        fr.index = k
    path, outp = tmp_path / "in.bin", tmp_path / "out.bin"
    with open(path, "wb") as f:
        f.write(struct.pack("<4i", cam.width, cam.height, len(frames), group))
        f.write(struct.pack("<6f", cam.fx, cam.fy, cam.cx, cam.cy, cam.near, cam.far))
        for k, fr in enumerate(frames):
            f.write(np.ascontiguousarray(fr.pose.T, np.float32).tobytes())      # corrected pose (column-major)
            f.write(np.ascontiguousarray(fr.pose_old.T, np.float32).tobytes())  # drifted pose
            f.write(np.ascontiguousarray(fr.depth, np.float32).tobytes())
            if k % group == 0:
                f.write(np.ascontiguousarray(fr.rgb, np.uint8).tobytes())
                f.write(np.ascontiguousarray(fr.color_valid, np.uint8).tobytes())
                f.write(np.ascontiguousarray(fr.quality, np.float32).tobytes())
    exe = build_excerpt(str(tmp_path))
    exp = tmp_path / "export"
    exp.mkdir()
    out = subprocess.run([exe, str(path), repr(res), str(outp), "--export", str(exp)], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr + out.stdout
    r = Reader(str(outp))
    drift = [fr.pose_old for fr in frames]
    corrected = [fr.pose for fr in frames]
    flow = OracleFlow(res, cam, frames, group)
    # stage 1: first fusion under the drifted poses, mesh + texture
    for k in flow.kfs:
        flow.reintegrate(k, 1, drift)
    st1 = read_state(r)
    textured = flow.mesh_and_texture()
    check_stage(st1, flow, textured, drift, "stage 1")
    # stage 2: loop closure of every key-frame (observations retracted, de- and re-integration)
    for k in flow.kfs:
        flow.o.retract_observations(flow.valid[k], k)
        flow.reintegrate(k, 0, drift)
        flow.reintegrate(k, 1, corrected)
    st2 = read_state(r)
    textured = flow.mesh_and_texture()
    check_stage(st2, flow, textured, corrected, "stage 2")
    check_exports(str(exp), st2)
    # stage 3: IntegrateFrame (fused convenience form) of a depth-only frame
    n, n_upd = flow.o.integrate_frame(frames[1].depth, None, None, corrected[1], cam, -1)
    assert (r.i64(), r.i64()) == (n, n_upd)
    assert r.i64() == flow.o.chunk_count() == r.i64(), "host id mirror after the fused frame"
    assert int(r.arr(np.int32, 1)[0]) == 1, "GetChunk on an unknown id must throw std::out_of_range"


@pytest.mark.gpu
def test_shim_per_keyframe_time_close_to_raw_c_abi(tmp_path):
    cam = synth.Camera()
    group = 7
    seq = synth.make_sequence(group, cam=cam, total=300, keyframe_every=group, start=42, with_drift=True)
    path = tmp_path / "in.bin"
    with open(path, "wb") as f:
        f.write(struct.pack("<4i", cam.width, cam.height, group, group))
        f.write(struct.pack("<6f", cam.fx, cam.fy, cam.cx, cam.cy, cam.near, cam.far))
        for k, fr in enumerate(seq.frames):
            f.write(np.ascontiguousarray(fr.pose.T, np.float32).tobytes())
            f.write(np.ascontiguousarray(fr.pose_old.T, np.float32).tobytes())
            f.write(np.ascontiguousarray(fr.depth, np.float32).tobytes())
            if k == 0:
                f.write(np.ascontiguousarray(fr.rgb, np.uint8).tobytes())
                f.write(np.ascontiguousarray(fr.color_valid, np.uint8).tobytes())
                f.write(np.ascontiguousarray(fr.quality, np.float32).tobytes())
    exe = build_excerpt(str(tmp_path))
    out = subprocess.run([exe, str(path), "0.005", str(tmp_path / "unused.bin"), "--bench", "20"], capture_output=True, text=True,
                         timeout=600)
    assert out.returncode == 0, out.stderr + out.stdout
    m = re.search(r"shim ([0-9.]+) us .* raw C ABI ([0-9.]+) us, ratio ([0-9.]+)", out.stdout)
    assert m, out.stdout
    print(out.stdout.strip())
    assert float(m.group(3)) < 1.3, out.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("res", (0.02, 0.005))
def test_shim_matches_oracle(tmp_path, res):
    seq = room_sequence(6)
    cam = seq.cam
    frames = seq.frames[:3]  # key-frame + two local depth frames
    path = tmp_path / "in.bin"
    with open(path, "wb") as f:
        f.write(struct.pack("<3i", cam.width, cam.height, len(frames)))
        f.write(struct.pack("<6f", cam.fx, cam.fy, cam.cx, cam.cy, cam.near, cam.far))
        for k, fr in enumerate(frames):
            f.write(np.ascontiguousarray(fr.pose.T, np.float32).tobytes())  # column-major
            f.write(struct.pack("<i", fr.index if k == 0 else -1))
            f.write(np.ascontiguousarray(fr.depth, np.float32).tobytes())
            if k == 0:
                f.write(fr.rgba().tobytes())
                f.write(np.ascontiguousarray(fr.quality, np.float32).tobytes())
    exe = build_driver(str(tmp_path))
    out = subprocess.run([exe, str(path), repr(res)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    m = re.search(r"ids=(\d+) valid=(\d+) chunks=(\d+) voxel_hash=([0-9a-f]+) obs=(\d+) obs_hash=([0-9a-f]+) meshes=(\d+)", out.stdout)
    assert m, out.stdout
    assert "out_of_range ok" in out.stdout
    # oracle
    o = OracleMap(res)
    kf = frames[0]
    ids, new = o.prepare(kf.depth, kf.pose, cam)
    nu, q = o.integrate(kf.depth, kf.rgba(), kf.quality, kf.pose, cam, ids, 1, kf.index)
    for lf in frames[1:]:
        nu, _ = o.integrate(lf.depth, None, None, lf.pose, cam, ids, 1, -1, nu)
    valid = o.finalize(ids, nu, new)
    sdf, w, col = o.download_chunks(valid)
    per_chunk = np.concatenate([sdf.view(np.uint32), w.view(np.uint32), col.view(np.uint32)], axis=1)  # n x 2048 words
    h = wsum(per_chunk)
    obs = []
    for cid in valid:
        v = o.observation(cid, kf.index)
        if v is not None:
            obs += [np.int32(kf.index).view(np.uint32), np.float32(v).view(np.uint32)]
    nobs = len(obs) // 2
    hobs = wsum(np.array(obs, np.uint32)) if obs else 0
    assert int(m.group(1)) == len(ids) and int(m.group(2)) == len(valid) and int(m.group(3)) == o.chunk_count()
    assert int(m.group(4), 16) == h, "voxel planes differ"
    assert int(m.group(5)) == nobs and int(m.group(6), 16) == hobs, "observations differ"
    assert int(m.group(7)) == len(o.meshes_to_update())
