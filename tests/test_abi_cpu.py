"""C-ABI checks that need no GPU: the library loads, exports every symbol the header declares,
the ctypes structs match the header's layout, and the product path fails loudly without CUDA."""
import ctypes as C
import os
import re
import subprocess

import pytest

from texturefusion_b200 import capi
from texturefusion_b200.build import LIB_PATH

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "texfusion.h")


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tf_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = capi.load()
    names = header_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/texfusion.h but not exported"
    assert sorted(capi.EXPORTS) == names, "capi.EXPORTS out of sync with the header"


def test_exports_are_plain_c_symbols():
    out = subprocess.run(["nm", "-D", "--defined-only", LIB_PATH], capture_output=True, text=True).stdout
    syms = {l.split()[-1] for l in out.splitlines() if l.strip()}
    for n in header_functions():
        assert n in syms


def test_struct_layouts_match_header():
    assert C.sizeof(capi.ChunkId) == 12
    assert C.sizeof(capi.Camera) == 32
    assert C.sizeof(capi.Truncation) == 20
    assert C.sizeof(capi.Pose) == 64
    assert C.sizeof(capi.GroupFrame) == 80
    assert C.sizeof(capi.FrameStats) == 40
    assert C.sizeof(capi.PatchDesc) == 32
    assert C.sizeof(capi.Config) == 72
    assert C.sizeof(capi.BatchItem) == 64
    # the header is valid C (not only C++)
    r = subprocess.run(["gcc", "-std=c99", "-fsyntax-only", "-x", "c", HEADER], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_sizeof_agrees_with_the_c_compiler(tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "texfusion.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu\\n",'
                   "sizeof(tf_chunk_id),sizeof(tf_camera),sizeof(tf_pose),sizeof(tf_config),sizeof(tf_group_frame),"
                   "sizeof(tf_frame_stats),sizeof(tf_batch_item),sizeof(tf_patch_desc));return 0;}\n")
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(v) for v in subprocess.check_output([str(exe)], text=True).split()]
    want = [C.sizeof(t) for t in (capi.ChunkId, capi.Camera, capi.Pose, capi.Config, capi.GroupFrame,
                                  capi.FrameStats, capi.BatchItem, capi.PatchDesc)]
    assert got == want


def test_no_cuda_device_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(capi.TexFusionError) as e:
        capi.Map(0.02)
    assert e.value.code == capi.TF_ERR_CUDA and "no CPU fallback" in str(e.value)


def test_create_argument_validation():
    lib = capi.load()
    h = C.c_void_p()
    cfg = capi.Config(4, 0.02, 1, capi.Truncation(*capi.DEFAULT_TRUNC), 0, 1, 0, 0, 0, 640, 480)
    assert lib.tf_create(C.byref(h), C.byref(cfg)) == capi.TF_ERR_INVALID  # chunk_dim != 8
    cfg.chunk_dim, cfg.voxel_res = 8, 0.0
    assert lib.tf_create(C.byref(h), C.byref(cfg)) == capi.TF_ERR_INVALID
    cfg.voxel_res, cfg.rank, cfg.n_ranks = 0.02, 2, 2
    assert lib.tf_create(C.byref(h), C.byref(cfg)) == capi.TF_ERR_INVALID
    assert b"rank" in lib.tf_last_error(None)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "texturefusion_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import oracle" not in txt and "from oracle" not in txt and "tf_oracle" not in txt, f
