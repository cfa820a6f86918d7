import sys, time, ctypes as C, numpy as np
sys.path.insert(0, '/root/repo')
from texturefusion_b200 import capi, synth
cam = synth.Camera()
seq = synth.make_sequence(120, cam=cam, total=300, keyframe_every=10, device="cuda")
m = capi.Map(0.005, max_frames=128)
for fr in seq.frames:
    m.upload_frame(fr.index, fr.depth, fr.rgba() if fr.is_keyframe else None, fr.quality if fr.is_keyframe else None)
m.sync()
L, h = m.L, m.h
camc = capi.make_camera(cam); st = capi.FrameStats()
poses = [capi.make_pose(fr.pose) for fr in seq.frames]
tb, te = [], []
for i, fr in enumerate(seq.frames):
    t0 = time.perf_counter()
    rc = L.tf_integrate_frame_begin(h, fr.index, int(fr.is_keyframe), C.byref(poses[i]), C.byref(camc), None, None, None, None, 0)
    t1 = time.perf_counter()
    L.tf_integrate_frame_end(h, C.byref(st))
    t2 = time.perf_counter()
    if i >= 20: tb.append(t1 - t0); te.append(t2 - t1)
print("begin median %.2f us, end median %.2f us" % (np.median(tb) * 1e6, np.median(te) * 1e6))
