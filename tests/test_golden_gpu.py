"""GPU vs the committed golden fixtures (tests/golden/*.npz) — no oracle in the loop."""
import os
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
import make_golden  # noqa: E402

from test_oracle_cpu import check_against_golden, load_inputs
from util import GpuAsOracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", list(make_golden.CASES))
def test_gpu_reproduces_golden(name):
    cam, frames, new_pose = load_inputs()
    res, _ = make_golden.CASES[name]
    out = make_golden.run_protocol(lambda: GpuAsOracle(res, width=cam.width, height=cam.height), cam, frames, new_pose)
    check_against_golden(out, name)
