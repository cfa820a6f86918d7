"""The C++ drop-in shim (texturefusion_b200/host/chisel_b200.h): compiled with g++ against
include/texfusion.h, linked to libtexfusion_b200.so, driven like ReIntegrateKeyframe, and
compared with the CPU oracle."""
import os
import re
import struct
import subprocess

import numpy as np
import pytest

from oracle import OracleMap
from texturefusion_b200.build import LIB_PATH

from util import room_sequence

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build_driver(tmp) -> str:
    exe = os.path.join(tmp, "shim_driver")
    cmd = ["g++", "-O2", "-std=c++14", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "texturefusion_b200", "host"),
           os.path.join(ROOT, "tests", "cpp", "shim_driver.cpp"), "-o", exe, LIB_PATH, f"-Wl,-rpath,{os.path.dirname(LIB_PATH)}"]
    subprocess.check_call(cmd)
    return exe


def wsum(words: np.ndarray) -> int:
    """position-weighted checksum over 32-bit words: sum(word * index) mod 2^64 (index from 1)."""
    w = np.ascontiguousarray(words).view(np.uint32).astype(np.uint64).ravel()
    with np.errstate(over="ignore"):
        return int((w * np.arange(1, len(w) + 1, dtype=np.uint64)).sum(dtype=np.uint64))


def test_shim_header_compiles_standalone(tmp_path):
    """No GPU needed: the header is self-contained C++14 (Eigen-free mode)."""
    src = tmp_path / "t.cpp"
    src.write_text('#include "chisel_b200.h"\nint main(){ chisel::PinholeCamera c; return c.GetCx() == 319 ? 0 : 1; }\n')
    exe = tmp_path / "t"
    subprocess.check_call(["g++", "-std=c++14", "-I", os.path.join(ROOT, "include"), "-I",
                           os.path.join(ROOT, "texturefusion_b200", "host"), str(src), "-o", str(exe), LIB_PATH,
                           f"-Wl,-rpath,{os.path.dirname(LIB_PATH)}"])
    assert subprocess.call([str(exe)]) == 0  # int-returning getter: cx 319.5 -> 319 (PinholeCamera.h:46-49)


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
