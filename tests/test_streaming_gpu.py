"""tf_stream_step (one C call per streaming step) against the separate calls it stands for, and against the
CPU oracle: per-frame ordered lists, flags, quality sums and the final map."""
import numpy as np
import pytest

from oracle import OracleMap
from texturefusion_b200 import capi
from texturefusion_b200.maphash import map_hash
from texturefusion_b200.streaming import FrameStreamer

from util import assert_maps_equal, room_sequence

pytestmark = pytest.mark.gpu


def test_stream_step_equals_separate_calls_and_oracle():
    seq = room_sequence(24)
    cam, frames = seq.cam, seq.frames
    res = 0.01
    maps = [capi.Map(res, max_frames=8), capi.Map(res, max_frames=8)]
    fs = [FrameStreamer(m, frames, cam, cap=1 << 16) for m in maps]
    o = OracleMap(res)
    for f in fs:
        f.prime(0)
    for i, fr in enumerate(frames[:-1]):  # (the ring would wrap on the last frame)
        n0 = fs[0].step(i)
        n1 = fs[1].step_calls(i)
        rgba = fr.rgba() if fr.is_keyframe else None
        n, nupd = o.integrate_frame(fr.depth, rgba, fr.quality, fr.pose, cam, fr.index if fr.is_keyframe else -1)
        assert n0 == n1 == n
        a, b = fs[0].lists(), fs[1].lists()
        for x, y in zip(a, b):
            assert np.array_equal(x, y)
        assert int(a[2].sum()) == nupd
    assert map_hash(maps[0]) == map_hash(maps[1])
    assert assert_maps_equal(maps[0], o, what="tf_stream_step")


def test_stream_step_reports_errors_of_either_half():
    seq = room_sequence(3)
    m = capi.Map(0.02, max_frames=4)
    f = FrameStreamer(m, seq.frames, seq.cam)
    with pytest.raises(capi.TexFusionError):  # frame 0 has not been staged
        f.step(0)
    f.prime(0)
    f.step(0)
    a = f._step_args(1)
    a.next_index, a.wait_index = -5, -1  # no ingest this step: fine
    f.step(1)
    f.stage(2)
    a = f._step_args(2)
    a.next_index = -1
    a.wait_index = 999  # a frame that is not in the store
    with pytest.raises(capi.TexFusionError):
        f.step(2)
