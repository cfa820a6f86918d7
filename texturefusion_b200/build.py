"""Builds libtexfusion_b200.so (sm_100a only) in-tree with nvcc."""
from __future__ import annotations

import os
import shutil
import subprocess

_PKG = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_PKG, "csrc")
LIB_PATH = os.path.join(_PKG, "libtexfusion_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false",  # the reference is un-fused AVX2 (no -mfma): keep every float op separately rounded
    "-Xcompiler", "-fPIC,-ffp-contract=off,-Wall",
    "-shared",
]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def sources() -> list[str]:
    return [os.path.join(_CSRC, f) for f in sorted(os.listdir(_CSRC))] + [
        os.path.join(os.path.dirname(_PKG), "include", "texfusion.h")]


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(s) > t for s in sources())


def build_library(force: bool = False, verbose: bool = False) -> str:
    if force or needs_build():
        cmd = [_nvcc(), *NVCC_FLAGS, "-o", LIB_PATH, os.path.join(_CSRC, "tf_capi.cu")]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        subprocess.check_call(cmd)
    return LIB_PATH


def build_timeline_variant() -> str:
    """The instrumented library tools/timeline.py needs (device + host time stamps, -DTF_TIMELINE);
    written to build/variants/timeline.so, selected with TEXFUSION_B200_LIB."""
    out = os.path.join(os.path.dirname(_PKG), "build", "variants", "timeline.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    subprocess.check_call([_nvcc(), *NVCC_FLAGS, "-DTF_TIMELINE", "-o", out, os.path.join(_CSRC, "tf_capi.cu")])
    return out


def build_variant(name: str, defines: list[str]) -> str:
    """An experimental build with extra -D flags (A/B runs: TEXFUSION_B200_LIB=<path> python bench.py ...);
    written to build/variants/<name>.so (git-ignored, travels to the GPU box)."""
    out = os.path.join(os.path.dirname(_PKG), "build", "variants", f"{name}.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    subprocess.check_call([_nvcc(), *NVCC_FLAGS, *[f"-D{d}" for d in defines], "-o", out, os.path.join(_CSRC, "tf_capi.cu")])
    return out


if __name__ == "__main__":
    import sys
    if len(sys.argv) > 1 and sys.argv[1] == "timeline":
        print(build_timeline_variant())
    elif len(sys.argv) > 2 and sys.argv[1] == "variant":
        print(build_variant(sys.argv[2], sys.argv[3:]))
    else:
        print(build_library(force=True, verbose=True))
