"""The streaming loop of the fusion path as a caller writes it against the C ABI — shared by
bench.py's end-to-end leg and the multi-process parity test, so that what is timed is what is tested.

Single GPU:   fuse_begin(i) | upload(i+1) | fuse_end(i) | wait_upload(i+1)   (double buffered, page-locked sources)
N GPUs:       every rank queues frame i on its shard (tf_integrate_frame_begin); rank 0 uploads frame
              i+1; every rank calls tf_broadcast_frame(i+1) — ONE ncclBroadcast inside the library, queued
              on the upload stream behind the copy —; tf_integrate_frame_end(i) collects the frame's
              lists; tf_wait_upload(i+1) closes the step.
              With N > 1 the ingest runs TWO frames ahead (`lookahead`): NCCL's broadcast kernel needs an
              SM slot, and integrate_kernel occupies every SM while it runs, so a broadcast queued one
              frame ahead would only start when the frame's kernels end and then sit on the critical path
              of the next frame; queued two ahead it runs in the bbox / culling phase of the frame between.
All arguments are marshalled once.  `step` is ONE C call (tf_stream_step = the sequence above inside the library:
a foreign call from Python costs 1-3 us, four or five of them would be a tenth of a step); the separate
calls stay available (`fuse_begin` / `stage` / `fuse_end` / `wait`, `step_calls`) and give the same results.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi


class FrameStreamer:
    def __init__(self, m: capi.Map, frames, cam, *, rank=0, world=1, cap=1 << 16, want_lists=True, lookahead=None):
        self.m, self.frames, self.cam, self.rank, self.world = m, frames, cam, rank, world
        self.lookahead = lookahead or (2 if world > 1 else 1)
        self.L = m.L
        self.nf = len(frames)
        self.camc = capi.make_camera(cam)
        self.poses = [capi.make_pose(fr.pose) for fr in frames]
        self.st = capi.FrameStats()
        self.cap = cap
        self.want_lists = want_lists
        self.pin_d, self.pin_c, self.pin_q = {}, {}, {}
        if rank == 0:  # the ingest rank holds the frames in page-locked host memory
            for i, fr in enumerate(frames):
                self.pin_d[i] = capi.PinnedBuffer((cam.height, cam.width), np.float32)
                self.pin_d[i].array[...] = fr.depth
                if fr.is_keyframe:
                    self.pin_c[i] = capi.PinnedBuffer((cam.height, cam.width, 4), np.uint8)
                    self.pin_c[i].array[...] = fr.rgba()
                    self.pin_q[i] = capi.PinnedBuffer((cam.height, cam.width), np.float32)
                    self.pin_q[i].array[...] = fr.quality
        self.out_ids = capi.PinnedBuffer((cap, 3), np.int32)
        self.out_new = capi.PinnedBuffer((cap,), np.uint8)
        self.out_upd = capi.PinnedBuffer((cap,), np.uint8)
        self.out_q = capi.PinnedBuffer((cap,), np.float32)

        self._steps = {}

    def _step_args(self, i):
        """tf_stream_step arguments of step(i), built once per frame of the ring."""
        a = self._steps.get(i)
        if a is None:
            fr, nx = self.frames[i], (i + self.lookahead) % self.nf
            fn = self.frames[nx]
            a = capi.StreamStepArgs()
            a.frame_index, a.use_color, a.pose = fr.index, int(fr.is_keyframe), self.poses[i]
            if self.want_lists:
                a.ids_out, a.is_new_out, a.updated_out, a.quality_out = self.out_ids.ptr, self.out_new.ptr, self.out_upd.ptr, self.out_q.ptr
                a.cap = self.cap
            a.next_index, a.next_has_color = fn.index, int(fn.is_keyframe)
            if self.rank == 0:
                a.next_depth = self.pin_d[nx].ptr
                if fn.is_keyframe:
                    a.next_rgba, a.next_quality = self.pin_c[nx].ptr, self.pin_q[nx].ptr
            a.broadcast_root = 0
            a.wait_index = self.frames[(i + 1) % self.nf].index
            self._steps[i] = a
        return a

    def _ok(self, rc):
        if rc != 0:
            raise capi.TexFusionError(rc, self.L.tf_last_error(self.m.h).decode())

    def upload(self, i):
        fr = self.frames[i]
        vp = C.c_void_p
        self._ok(self.L.tf_upload_frame(self.m.h, fr.index, vp(self.pin_d[i].ptr),
                                        vp(self.pin_c[i].ptr) if fr.is_keyframe else None,
                                        vp(self.pin_q[i].ptr) if fr.is_keyframe else None))

    def stage(self, i):
        """Frame i on its way into every rank's frame store (asynchronous)."""
        fr = self.frames[i]
        if self.rank == 0:
            self.upload(i)
        if self.world > 1:
            self._ok(self.L.tf_broadcast_frame(self.m.h, fr.index, int(fr.is_keyframe), 0))

    def fuse_begin(self, i):
        fr = self.frames[i]
        vp = C.c_void_p
        if self.want_lists:
            rc = self.L.tf_integrate_frame_begin(self.m.h, fr.index, int(fr.is_keyframe), C.byref(self.poses[i]), C.byref(self.camc),
                                                 vp(self.out_ids.ptr), vp(self.out_new.ptr), vp(self.out_upd.ptr),
                                                 vp(self.out_q.ptr), self.cap)
        else:
            rc = self.L.tf_integrate_frame_begin(self.m.h, fr.index, int(fr.is_keyframe), C.byref(self.poses[i]), C.byref(self.camc),
                                                 None, None, None, None, 0)
        self._ok(rc)

    def fuse_end(self):
        self._ok(self.L.tf_integrate_frame_end(self.m.h, C.byref(self.st)))
        return self.st.n_chunks

    def fuse(self, i):
        self.fuse_begin(i)
        return self.fuse_end()

    def wait(self, i):
        self._ok(self.L.tf_wait_upload(self.m.h, self.frames[i].index))

    def prime(self, i):
        """Stage the frames a first step(i) expects to be on their way: i .. i + lookahead - 1."""
        for k in range(self.lookahead):
            self.stage((i + k) % self.nf)

    def step(self, i):
        """One end-to-end step for frame i (frames i .. i+lookahead-1 have been staged): one frame's ingest
        is queued while this frame's kernels run, and the NEXT frame to be fused has arrived when the step ends."""
        self._ok(self.L.tf_stream_step(self.m.h, C.byref(self.camc), C.byref(self._step_args(i)), C.byref(self.st)))
        return self.st.n_chunks

    def step_calls(self, i):
        """The same step as separate C calls."""
        self.fuse_begin(i)                           # frame i's kernels are running ...
        self.stage((i + self.lookahead) % self.nf)   # ... while the host queues a later frame's copy (+ broadcast)
        n = self.fuse_end()
        self.wait((i + 1) % self.nf)
        return n

    def lists(self):
        """The last fused frame's ordered outputs (copies)."""
        n = min(self.st.n_chunks, self.cap)
        return (self.out_ids.array[:n].copy(), self.out_new.array[:n].copy(), self.out_upd.array[:n].copy(),
                self.out_q.array[:n].copy())
