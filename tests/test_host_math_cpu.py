"""Host-side arithmetic of the library that needs no GPU (compiled with g++ from the product's own header)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cuda_include():
    for cand in ("/usr/local/cuda/include", os.path.join(os.path.dirname(shutil.which("nvcc") or ""), "..", "include")):
        if os.path.exists(os.path.join(cand, "cuda_runtime.h")):
            return cand
    return None


@pytest.mark.skipif(_cuda_include() is None, reason="CUDA headers not found")
def test_projection_range_test_covers_every_voxel_centre(tmp_path):
    """FrameDev::z_safe (tf_host_math.h) is the per-chunk limit above which integrate_kernel evaluates the
    reference's division with div.rn's fast-path sequence and no range check: for random rigid poses, seven
    voxel sizes and both product associations every voxel centre of a chunk that passes the test lies in
    2^-17 < cz < 2^21, |c| < 2^21; non-rigid transforms and out-of-range intrinsics widen or disable it."""
    exe = tmp_path / "host_math_check"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-I", _cuda_include(),
                           os.path.join(ROOT, "tests", "cpp", "host_math_check.cpp"), "-o", str(exe)])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.startswith("ok "), out.stdout + out.stderr
    assert int(out.stdout.split()[1]) > 10_000_000
