"""ctypes binding of libtexfusion_b200.so — the C ABI declared in include/texfusion.h.

This is the same call surface the C++ shim (texturefusion_b200/host/) uses.  There is no
fallback: if the library is missing, or no CUDA device is present, the calls fail loudly.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .build import LIB_PATH

TF_OK, TF_ERR_INVALID, TF_ERR_CUDA, TF_ERR_CAPACITY, TF_ERR_NOT_FOUND, TF_ERR_ATLAS_FULL = 0, -1, -2, -3, -4, -5

# QuadraticTruncator(0.0019, 0.00152, 0.001504, 6) + ConstantWeighter(1): GCFusion/MobileFusion.h:215-228
DEFAULT_TRUNC = (0.0019, 0.00152, 0.001504, 6.0, 1.0)

EXPORTS = [
    "tf_create", "tf_destroy", "tf_last_error", "tf_reset", "tf_set_truncation", "tf_host_alloc", "tf_host_free",
    "tf_upload_frame", "tf_upload_keyframe_rgb", "tf_release_frame", "tf_frame_device_ptrs",
    "tf_comm_unique_id", "tf_comm_init", "tf_broadcast_frame",
    "tf_prepare", "tf_integrate", "tf_integrate_group", "tf_remove_chunks", "tf_integrate_frame", "tf_integrate_frame_begin", "tf_integrate_frame_end",
    "tf_stream_step", "tf_integrate_batch", "tf_mesh_chunks", "tf_has_chunk", "tf_chunk_count", "tf_list_chunks", "tf_download_chunks",
    "tf_atlas_alloc_slot", "tf_atlas_update", "tf_atlas_download", "tf_atlas_copy_to_device", "tf_atlas_patch_size", "tf_patch_texcoords", "tf_sync", "tf_wait_upload",
    "tf_pre_upload_depth_u16", "tf_pre_bilateral", "tf_pre_normal_map", "tf_pre_refine_keyframe", "tf_pre_refine_newframe", "tf_pre_refine_depth_by_normal", "tf_pre_color_quality",
    "tf_pre_download",
    "tf_get_counters", "tf_stream", "tf_copy_stream", "tf_set_profiling", "tf_get_kernel_time", "tf_get_stage_times", "tf_debug_project", "tf_debug_divide",
]


class ChunkId(C.Structure):
    _fields_ = [("x", C.c_int32), ("y", C.c_int32), ("z", C.c_int32)]


class Camera(C.Structure):
    _fields_ = [("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float),
                ("width", C.c_int32), ("height", C.c_int32), ("near_plane", C.c_float),
                ("far_plane", C.c_float)]


class Truncation(C.Structure):
    _fields_ = [("quad", C.c_float), ("lin", C.c_float), ("cst", C.c_float), ("scale", C.c_float),
                ("weight", C.c_float)]


class Pose(C.Structure):
    _fields_ = [("m", C.c_float * 16)]


class StreamStepArgs(C.Structure):
    _fields_ = [("frame_index", C.c_int32), ("use_color", C.c_int32), ("pose", Pose),
                ("ids_out", C.c_void_p), ("is_new_out", C.c_void_p), ("updated_out", C.c_void_p), ("quality_out", C.c_void_p),
                ("cap", C.c_int64), ("next_index", C.c_int32), ("next_has_color", C.c_int32),
                ("next_depth", C.c_void_p), ("next_rgba", C.c_void_p), ("next_quality", C.c_void_p),
                ("broadcast_root", C.c_int32), ("wait_index", C.c_int32)]


class Config(C.Structure):
    _fields_ = [("chunk_dim", C.c_int32), ("voxel_res", C.c_float), ("use_color", C.c_int32),
                ("trunc", Truncation), ("device", C.c_int32), ("n_ranks", C.c_int32), ("rank", C.c_int32),
                ("max_chunks", C.c_int64), ("max_frames", C.c_int32), ("width", C.c_int32),
                ("height", C.c_int32), ("dot3_order", C.c_int32)]


class GroupFrame(C.Structure):
    _fields_ = [("frame_index", C.c_int32), ("use_color", C.c_int32), ("flag", C.c_int32),
                ("reserved", C.c_int32), ("pose", Pose)]


class FrameStats(C.Structure):
    _fields_ = [("n_chunks", C.c_int64), ("n_new", C.c_int64), ("n_updated", C.c_int64),
                ("n_removed", C.c_int64), ("voxel_updates", C.c_int64)]


class BatchItem(C.Structure):
    _fields_ = [("flag", C.c_int32), ("n_frames", C.c_int32), ("frames", C.POINTER(GroupFrame)),
                ("ids", C.c_void_p), ("n_ids", C.c_int64), ("valid_out", C.c_void_p),
                ("quality_out", C.c_void_p), ("cap", C.c_int64), ("n_valid_out", C.POINTER(C.c_int64))]


class PatchDesc(C.Structure):
    _fields_ = [("texloc", C.c_uint64), ("frame_index", C.c_int32), ("x", C.c_int32), ("y", C.c_int32),
                ("w", C.c_int32), ("h", C.c_int32)]


class Counters(C.Structure):
    _fields_ = [("kernel_launches", C.c_int64), ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64),
                ("frames_integrated", C.c_int64), ("voxel_updates", C.c_int64),
                ("pool_capacity", C.c_int64), ("pool_used", C.c_int64)]


_LIB = None


def load() -> C.CDLL:
    """dlopen the in-tree library and declare the prototypes of include/texfusion.h."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = os.environ.get("TEXFUSION_B200_LIB", LIB_PATH)  # override: A/B builds of the same library
    if not os.path.exists(path):
        raise RuntimeError(f"{path} is missing: build it with `python -m texturefusion_b200.build` "
                           "(there is no CPU fallback)")
    L = C.CDLL(path)
    vp, i64 = C.c_void_p, C.c_int64
    L.tf_create.argtypes = [C.POINTER(vp), C.POINTER(Config)]
    L.tf_destroy.argtypes = [vp]
    L.tf_destroy.restype = None
    L.tf_last_error.argtypes = [vp]
    L.tf_last_error.restype = C.c_char_p
    L.tf_reset.argtypes = [vp]
    L.tf_set_truncation.argtypes = [vp, C.POINTER(Truncation)]
    L.tf_host_alloc.argtypes = [C.c_size_t]
    L.tf_host_alloc.restype = vp
    L.tf_host_free.argtypes = [vp]
    L.tf_host_free.restype = None
    L.tf_upload_frame.argtypes = [vp, C.c_int32, vp, vp, vp]
    L.tf_upload_keyframe_rgb.argtypes = [vp, C.c_int32, vp, vp]
    L.tf_release_frame.argtypes = [vp, C.c_int32]
    L.tf_frame_device_ptrs.argtypes = [vp, C.c_int32, C.c_int, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
    L.tf_comm_unique_id.argtypes = [vp]
    L.tf_comm_init.argtypes = [vp, vp]
    L.tf_broadcast_frame.argtypes = [vp, C.c_int32, C.c_int, C.c_int]
    L.tf_prepare.argtypes = [vp, C.c_int32, C.POINTER(Pose), C.POINTER(Camera), vp, vp, i64, C.POINTER(i64)]
    L.tf_integrate.argtypes = [vp, C.c_int32, C.c_int, C.POINTER(Pose), C.POINTER(Camera), vp, i64, C.c_int, vp, vp]
    L.tf_integrate_group.argtypes = [vp, C.POINTER(GroupFrame), C.c_int32, C.POINTER(Camera), vp, i64, vp, vp]
    L.tf_remove_chunks.argtypes = [vp, vp, i64]
    L.tf_integrate_frame.argtypes = [vp, C.c_int32, C.c_int, C.POINTER(Pose), C.POINTER(Camera),
                                     C.POINTER(FrameStats), vp, vp, vp, vp, i64]
    L.tf_integrate_frame_begin.argtypes = [vp, C.c_int32, C.c_int, C.POINTER(Pose), C.POINTER(Camera), vp, vp, vp, vp, i64]
    L.tf_integrate_frame_end.argtypes = [vp, C.POINTER(FrameStats)]
    L.tf_stream_step.argtypes = [vp, C.POINTER(Camera), C.POINTER(StreamStepArgs), C.POINTER(FrameStats)]
    L.tf_integrate_batch.argtypes = [vp, C.POINTER(BatchItem), i64, C.POINTER(Camera)]
    L.tf_has_chunk.argtypes = [vp, ChunkId]
    L.tf_chunk_count.argtypes = [vp]
    L.tf_chunk_count.restype = i64
    L.tf_list_chunks.argtypes = [vp, vp, i64, C.POINTER(i64)]
    L.tf_download_chunks.argtypes = [vp, vp, i64, vp, vp, vp]
    L.tf_mesh_chunks.argtypes = [vp, vp, i64, vp, vp, vp, vp, vp, vp, i64, i64]
    L.tf_atlas_alloc_slot.argtypes = [vp, ChunkId, C.POINTER(C.c_uint64)]
    L.tf_atlas_update.argtypes = [vp, C.POINTER(PatchDesc), i64]
    L.tf_atlas_download.argtypes = [vp, C.c_uint64, C.c_uint64, vp]
    L.tf_atlas_copy_to_device.argtypes = [vp, C.c_uint64, C.c_uint64, vp]
    L.tf_atlas_patch_size.argtypes = [vp, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    L.tf_patch_texcoords.argtypes = [vp, C.c_int32, C.POINTER(Pose), C.POINTER(Camera), i64, vp, vp, vp, vp, vp, vp]
    camp = C.POINTER(Camera)
    L.tf_pre_normal_map.argtypes = [vp, C.c_int32, camp]
    L.tf_pre_upload_depth_u16.argtypes = [vp, C.c_int32, vp, C.c_float, C.c_float]
    L.tf_pre_bilateral.argtypes = [vp, C.c_int32, C.c_int32, C.c_float, C.c_float]
    L.tf_pre_refine_keyframe.argtypes = [vp, C.c_int32, C.c_int32, vp, camp]
    L.tf_pre_refine_newframe.argtypes = [vp, C.c_int32, C.c_int32, vp, camp]
    L.tf_pre_refine_depth_by_normal.argtypes = [vp, C.c_int32, camp]
    L.tf_pre_color_quality.argtypes = [vp, C.c_int32, vp, camp]
    L.tf_pre_download.argtypes = [vp, C.c_int32, vp, vp, vp, vp, vp]
    L.tf_sync.argtypes = [vp]
    L.tf_wait_upload.argtypes = [vp, C.c_int32]
    L.tf_get_counters.argtypes = [vp, C.POINTER(Counters)]
    L.tf_stream.argtypes = [vp]
    L.tf_stream.restype = vp
    L.tf_copy_stream.argtypes = [vp]
    L.tf_copy_stream.restype = vp
    L.tf_set_profiling.argtypes = [vp, C.c_int]
    L.tf_get_kernel_time.argtypes = [vp, C.c_int, C.POINTER(C.c_double), C.POINTER(i64), C.POINTER(C.c_double)]
    L.tf_get_stage_times.argtypes = [vp, C.c_int, C.POINTER(C.c_double)]
    L.tf_debug_project.argtypes = [vp, vp, vp, i64, C.c_float, C.c_float, vp, vp, vp]
    L.tf_debug_divide.argtypes = [vp, vp, vp, i64, vp, vp, vp]
    _LIB = L
    return L


def comm_unique_id() -> bytes:
    """ncclGetUniqueId through the library (call on one rank, share the bytes with the others)."""
    L = load()
    buf = (C.c_uint8 * 128)()
    rc = L.tf_comm_unique_id(buf)
    if rc != TF_OK:
        raise TexFusionError(rc, L.tf_last_error(None).decode())
    return bytes(buf)


def share_unique_id(dist, device) -> bytes:
    """Rank 0's NCCL id handed to every rank through an initialised torch.distributed group."""
    import torch
    t = torch.zeros(128, dtype=torch.uint8, device=device)
    if dist.get_rank() == 0:
        t = torch.frombuffer(bytearray(comm_unique_id()), dtype=torch.uint8).to(device)
    dist.broadcast(t, src=0)
    return bytes(t.cpu().numpy().tobytes())


class TexFusionError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"[{code}] {msg}")
        self.code = code


def make_pose(pose) -> Pose:
    """4x4 camera->world (row-major NumPy) -> column-major tf_pose (Eigen::Affine3f layout)."""
    p = Pose()
    flat = np.ascontiguousarray(np.asarray(pose, np.float32).T).reshape(16)
    for i in range(16):
        p.m[i] = float(flat[i])
    return p


def make_camera(cam) -> Camera:
    return Camera(cam.fx, cam.fy, cam.cx, cam.cy, cam.width, cam.height, cam.near, cam.far)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class PinnedBuffer:
    """A NumPy view over tf_host_alloc'ed (page-locked) memory."""

    def __init__(self, shape, dtype):
        self.L = load()
        self.nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        self.ptr = self.L.tf_host_alloc(self.nbytes)
        if not self.ptr:
            raise TexFusionError(TF_ERR_CUDA, "tf_host_alloc failed")
        buf = (C.c_uint8 * self.nbytes).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=dtype).reshape(shape)

    def free(self):
        if self.ptr:
            self.array = None
            self.L.tf_host_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Map:
    """Thin object wrapper over a tf_map handle.  Methods mirror the C ABI one to one."""

    def __init__(self, voxel_res: float, *, use_color=True, trunc=DEFAULT_TRUNC, device=0, n_ranks=1, rank=0,
                 max_chunks=0, max_frames=0, width=640, height=480, dot3_order=0):
        self.L = load()
        self.h = C.c_void_p()
        cfg = Config(8, voxel_res, int(use_color), Truncation(*trunc), device, n_ranks, rank, max_chunks,
                     max_frames, width, height, dot3_order)
        rc = self.L.tf_create(C.byref(self.h), C.byref(cfg))
        if rc != TF_OK:
            msg = self.L.tf_last_error(None).decode()
            self.h = None
            raise TexFusionError(rc, msg)
        self.width, self.height = width, height
        self.voxel_res = float(np.float32(voxel_res))
        self.list_cap = 1 << 19

    def close(self):
        if getattr(self, "h", None):
            self.L.tf_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc < 0:
            raise TexFusionError(rc, self.L.tf_last_error(self.h).decode())
        return rc

    def set_truncation(self, trunc):
        self._check(self.L.tf_set_truncation(self.h, C.byref(Truncation(*trunc))))

    # frame store -----------------------------------------------------------------------
    def upload_frame(self, frame_index, depth, rgba=None, quality=None):
        d = np.ascontiguousarray(depth, np.float32)
        c = None if rgba is None else np.ascontiguousarray(rgba, np.uint8)
        q = None if quality is None else np.ascontiguousarray(quality, np.float32)
        assert d.size == self.width * self.height
        self._check(self.L.tf_upload_frame(self.h, frame_index, _p(d), _p(c), _p(q)))
        # pageable sources are staged by the runtime before the call returns; pinned ones
        # must stay untouched until the next synchronising call, so keep references.
        self._keep = (d, c, q)

    def upload_keyframe_rgb(self, frame_index, rgb, color_valid=None):
        r = np.ascontiguousarray(rgb, np.uint8)
        v = None if color_valid is None else np.ascontiguousarray(color_valid, np.uint8)
        self._check(self.L.tf_upload_keyframe_rgb(self.h, frame_index, _p(r), _p(v)))
        self._keep2 = (r, v)

    def release_frame(self, frame_index):
        self._check(self.L.tf_release_frame(self.h, frame_index))

    def frame_device_ptrs(self, frame_index, has_color):
        d, c, q = C.c_void_p(), C.c_void_p(), C.c_void_p()
        self._check(self.L.tf_frame_device_ptrs(self.h, frame_index, int(has_color), C.byref(d), C.byref(c), C.byref(q)))
        return d.value, c.value, q.value

    # multi-GPU ---------------------------------------------------------------------------
    def comm_init(self, unique_id: bytes):
        """ncclCommInitRank with the id from comm_unique_id() of one rank (collective)."""
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        self._check(self.L.tf_comm_init(self.h, buf))

    def broadcast_frame(self, frame_index, has_color, root=0):
        self._check(self.L.tf_broadcast_frame(self.h, frame_index, int(has_color), root))

    def wait_upload(self, frame_index):
        self._check(self.L.tf_wait_upload(self.h, frame_index))

    # hot path ----------------------------------------------------------------------------
    def prepare(self, frame_index, pose, cam, cap=None):
        cap = cap or self.list_cap
        ids = np.empty((cap, 3), np.int32)
        new = np.empty(cap, np.uint8)
        n = C.c_int64(0)
        self._check(self.L.tf_prepare(self.h, frame_index, C.byref(make_pose(pose)), C.byref(make_camera(cam)),
                                      _p(ids), _p(new), cap, C.byref(n)))
        return ids[:n.value].copy(), new[:n.value].copy()

    def integrate(self, frame_index, use_color, pose, cam, ids, flag, needs_update=None):
        ids = np.ascontiguousarray(ids, np.int32).reshape(-1, 3)
        n = len(ids)
        nu = np.zeros(n, np.uint8) if needs_update is None else np.ascontiguousarray(needs_update, np.uint8)
        q = np.zeros(n, np.float32)
        self._check(self.L.tf_integrate(self.h, frame_index, int(use_color), C.byref(make_pose(pose)),
                                        C.byref(make_camera(cam)), _p(ids), n, int(flag), _p(nu), _p(q)))
        return nu, q

    def integrate_group(self, frames, cam, ids, needs_update=None):
        """frames: list of (frame_index, use_color, flag, pose)."""
        arr = (GroupFrame * len(frames))()
        for i, (fi, uc, fl, pose) in enumerate(frames):
            arr[i] = GroupFrame(fi, int(uc), int(fl), 0, make_pose(pose))
        ids = np.ascontiguousarray(ids, np.int32).reshape(-1, 3)
        n = len(ids)
        nu = np.zeros(n, np.uint8) if needs_update is None else np.ascontiguousarray(needs_update, np.uint8)
        q = np.zeros(n, np.float32)
        self._check(self.L.tf_integrate_group(self.h, arr, len(frames), C.byref(make_camera(cam)), _p(ids), n,
                                              _p(nu), _p(q)))
        return nu, q

    def remove_chunks(self, ids):
        ids = np.ascontiguousarray(ids, np.int32).reshape(-1, 3)
        self._check(self.L.tf_remove_chunks(self.h, _p(ids), len(ids)))

    def integrate_frame(self, frame_index, use_color, pose, cam, want_lists=True, cap=None):
        cap = cap or self.list_cap
        st = FrameStats()
        if want_lists:
            ids = np.empty((cap, 3), np.int32)
            new = np.empty(cap, np.uint8)
            upd = np.empty(cap, np.uint8)
            q = np.empty(cap, np.float32)
        else:
            ids = new = upd = q = None
        self._check(self.L.tf_integrate_frame(self.h, frame_index, int(use_color), C.byref(make_pose(pose)),
                                              C.byref(make_camera(cam)), C.byref(st), _p(ids), _p(new), _p(upd),
                                              _p(q), cap if want_lists else 0))
        if want_lists:
            n = st.n_chunks
            return st, ids[:n].copy(), new[:n].copy(), upd[:n].copy(), q[:n].copy()
        return st, None, None, None, None

    def marshal_batch(self, items, cap=None, want_lists=True):
        """Builds the tf_batch_item array of integrate_batch (kept separate so that a benchmark can
        time the C call alone).  Returns an opaque tuple for run_batch.  want_lists=False: the
        re-integration items ask for nothing back (streaming fusion of a frame sequence)."""
        cap = cap or self.list_cap
        n = len(items)
        arr = (BatchItem * n)()
        keep, outs = [], []
        for k, it in enumerate(items):
            fr = (GroupFrame * len(it["frames"]))()
            for i, (fi, uc, pose) in enumerate(it["frames"]):
                fr[i] = GroupFrame(fi, int(uc), int(it["flag"]), 0, make_pose(pose))
            keep.append(fr)
            arr[k].flag = int(it["flag"])
            arr[k].n_frames = len(it["frames"])
            arr[k].frames = fr
            if it["flag"] == 0:
                ids = np.ascontiguousarray(it["ids"], np.int32).reshape(-1, 3)
                keep.append(ids)
                arr[k].ids = ids.ctypes.data
                arr[k].n_ids = len(ids)
                outs.append(None)
            elif not want_lists:
                outs.append(None)
            else:
                valid = np.empty((cap, 3), np.int32)
                q = np.empty(cap, np.float32)
                nv = C.c_int64(0)
                keep += [valid, q, nv]
                arr[k].valid_out = valid.ctypes.data
                arr[k].quality_out = q.ctypes.data
                arr[k].cap = cap
                arr[k].n_valid_out = C.pointer(nv)
                outs.append((valid, q, nv))
        return arr, n, keep, outs

    def run_batch(self, marshalled, cam):
        arr, n, _keep, outs = marshalled
        self._check(self.L.tf_integrate_batch(self.h, arr, n, C.byref(make_camera(cam))))
        return [None if o is None else (o[0][:o[2].value].copy(), o[1][:o[2].value].copy()) for o in outs]

    def integrate_batch(self, items, cam):
        """items: list of dicts {flag, frames:[(frame_index,use_color,pose)], ids (flag 0)}.
        Returns, per item, None (flag 0) or (valid_ids, quality) (flag 1)."""
        return self.run_batch(self.marshal_batch(items), cam)

    # queries -------------------------------------------------------------------------------
    def has_chunk(self, id3) -> bool:
        return bool(self._check(self.L.tf_has_chunk(self.h, ChunkId(int(id3[0]), int(id3[1]), int(id3[2])))))

    def chunk_count(self) -> int:
        return int(self.L.tf_chunk_count(self.h))

    def list_chunks(self) -> np.ndarray:
        n = self.chunk_count()
        out = np.empty((max(n, 1), 3), np.int32)
        got = C.c_int64(0)
        self._check(self.L.tf_list_chunks(self.h, _p(out), n, C.byref(got)))
        return out[:got.value]

    def download_chunks(self, ids):
        ids = np.ascontiguousarray(ids, np.int32).reshape(-1, 3)
        n = len(ids)
        sdf = np.empty((n, 512), np.float32)
        w = np.empty((n, 512), np.float32)
        col = np.empty((n, 2048), np.uint16)
        self._check(self.L.tf_download_chunks(self.h, _p(ids), n, _p(sdf), _p(w), _p(col)))
        return sdf, w, col

    def mesh_chunks(self, ids):
        """ChunkManager::GenerateMeshEfficient per chunk on the device: (vert_off, idx_off, vertices,
        normals, colors, indices)."""
        ids = np.ascontiguousarray(ids, np.int32).reshape(-1, 3)
        n = len(ids)
        voff = np.zeros(n + 1, np.int64)
        ioff = np.zeros(n + 1, np.int64)
        self._check(self.L.tf_mesh_chunks(self.h, _p(ids), n, _p(voff), _p(ioff), None, None, None, None, 0, 0))
        nv, ni = int(voff[-1]), int(ioff[-1])
        vert = np.empty((max(nv, 1), 3), np.float32)
        norm = np.empty((max(nv, 1), 3), np.float32)
        col = np.empty((max(nv, 1), 3), np.float32)
        idx = np.empty(max(ni, 1), np.int32)
        self._check(self.L.tf_mesh_chunks(self.h, _p(ids), n, _p(voff), _p(ioff), _p(vert), _p(norm), _p(col), _p(idx), nv, ni))
        return voff, ioff, vert[:nv], norm[:nv], col[:nv], idx[:ni]

    # atlas ---------------------------------------------------------------------------------
    def atlas_patch_size(self):
        w, h = C.c_int32(), C.c_int32()
        self._check(self.L.tf_atlas_patch_size(self.h, C.byref(w), C.byref(h)))
        return w.value, h.value

    def atlas_alloc_slot(self, id3) -> int:
        loc = C.c_uint64()
        self._check(self.L.tf_atlas_alloc_slot(self.h, ChunkId(int(id3[0]), int(id3[1]), int(id3[2])), C.byref(loc)))
        return int(loc.value)

    def atlas_update(self, patches):
        """patches: iterable of (texloc, frame_index, x, y, w, h)."""
        patches = list(patches)
        arr = (PatchDesc * max(len(patches), 1))()
        for i, p in enumerate(patches):
            arr[i] = PatchDesc(*[int(v) for v in p])
        self._check(self.L.tf_atlas_update(self.h, arr, len(patches)))

    def atlas_download(self, hot_start, hot_end) -> np.ndarray:
        out = np.empty((hot_end - hot_start) * 3, np.uint8)
        self._check(self.L.tf_atlas_download(self.h, C.c_uint64(hot_start), C.c_uint64(hot_end), _p(out)))
        return out

    def atlas_copy_to_device(self, hot_start, hot_end, device_ptr: int):
        """Hot rows device to device (the CUDA side of a CUDA-GL pixel-unpack-buffer upload)."""
        self._check(self.L.tf_atlas_copy_to_device(self.h, int(hot_start), int(hot_end), C.c_void_p(device_ptr)))

    def patch_texcoords(self, frame_index, world_to_camera, cam, offsets, vertices, colors):
        """Patch::CalculateTexCoords for a batch of meshes; returns (texcoord, texcolor, results[n,6])."""
        off = np.ascontiguousarray(offsets, np.int64)
        v = np.ascontiguousarray(vertices, np.float32).reshape(-1, 3)
        c = np.ascontiguousarray(colors, np.float32).reshape(-1, 3)
        n = len(off) - 1
        tc = np.empty((len(v), 2), np.float32)
        col = np.empty((len(v), 3), np.float32)
        res = np.empty((max(n, 1), 6), np.int32)
        self._check(self.L.tf_patch_texcoords(self.h, frame_index, C.byref(make_pose(world_to_camera)),
                                              C.byref(make_camera(cam)), n, _p(off), _p(v), _p(c), _p(tc), _p(col), _p(res)))
        return tc, col, res[:n]

    # frame pre-processing (BasicAPI.cpp loops on the frame store) --------------------------------
    def pre_upload_depth_u16(self, frame_index, depth_u16, depth_scale, max_depth):
        d = np.ascontiguousarray(depth_u16, np.uint16)
        self._check(self.L.tf_pre_upload_depth_u16(self.h, frame_index, _p(d), depth_scale, max_depth))
        self._keep = d  # (the copy is asynchronous)

    def pre_bilateral(self, frame_index, d=9, sigma_color=0.03, sigma_space=10.0):
        self._check(self.L.tf_pre_bilateral(self.h, frame_index, d, sigma_color, sigma_space))

    def pre_normal_map(self, frame_index, cam):
        self._check(self.L.tf_pre_normal_map(self.h, frame_index, C.byref(make_camera(cam))))

    def pre_refine_keyframe(self, keyframe_index, new_index, ref_to_new, cam):
        T = np.ascontiguousarray(np.asarray(ref_to_new)[:3, :4], np.float32)
        self._check(self.L.tf_pre_refine_keyframe(self.h, keyframe_index, new_index, _p(T), C.byref(make_camera(cam))))

    def pre_refine_newframe(self, keyframe_index, new_index, new_to_ref, cam):
        T = np.ascontiguousarray(np.asarray(new_to_ref)[:3, :4], np.float32)
        self._check(self.L.tf_pre_refine_newframe(self.h, keyframe_index, new_index, _p(T), C.byref(make_camera(cam))))

    def pre_refine_depth_by_normal(self, frame_index, cam):
        self._check(self.L.tf_pre_refine_depth_by_normal(self.h, frame_index, C.byref(make_camera(cam))))

    def pre_color_quality(self, frame_index, rgb, cam):
        c = np.ascontiguousarray(rgb, np.uint8)
        self._check(self.L.tf_pre_color_quality(self.h, frame_index, _p(c), C.byref(make_camera(cam))))
        self._keep = c  # (the copy is asynchronous)

    def pre_download(self, frame_index, *, depth=False, normal=False, weight=False, color_valid=False, quality=False) -> dict:
        H, W = self.height, self.width
        out = {}
        if depth:
            out["depth"] = np.empty((H, W), np.float32)
        if normal:
            out["normal"] = np.empty((3, H, W), np.float32)
        if weight:
            out["weight"] = np.empty((H, W), np.float32)
        if color_valid:
            out["color_valid"] = np.empty((H, W), np.uint8)
        if quality:
            out["quality"] = np.empty((H, W), np.float32)
        g = lambda k: _p(out[k]) if k in out else None  # noqa: E731
        self._check(self.L.tf_pre_download(self.h, frame_index, g("depth"), g("normal"), g("weight"), g("color_valid"), g("quality")))
        return out

    # misc -----------------------------------------------------------------------------------
    def sync(self):
        self._check(self.L.tf_sync(self.h))

    def reset(self):
        self._check(self.L.tf_reset(self.h))

    def counters(self) -> dict:
        c = Counters()
        self._check(self.L.tf_get_counters(self.h, C.byref(c)))
        return {k: getattr(c, k) for k, _ in Counters._fields_}

    def stream(self) -> int:
        return int(self.L.tf_stream(self.h) or 0)

    def copy_stream(self) -> int:
        return int(self.L.tf_copy_stream(self.h) or 0)

    def set_profiling(self, level):
        self._check(self.L.tf_set_profiling(self.h, int(level)))

    def debug_project(self, c, cz, f, ch):
        c = np.ascontiguousarray(c, np.float32)
        cz = np.ascontiguousarray(cz, np.float32)
        n = c.size
        uf, ue, acc = np.empty(n, np.int32), np.empty(n, np.int32), np.empty(n, np.uint8)
        self._check(self.L.tf_debug_project(self.h, _p(c), _p(cz), n, C.c_float(f), C.c_float(ch), _p(uf), _p(ue), _p(acc)))
        return uf, ue, acc

    def debug_divide(self, num, den):
        num = np.ascontiguousarray(num, np.float32)
        den = np.ascontiguousarray(den, np.float32)
        n = num.size
        qk, qi, acc = np.empty(n, np.float32), np.empty(n, np.float32), np.empty(n, np.uint8)
        self._check(self.L.tf_debug_divide(self.h, _p(num), _p(den), n, _p(qk), _p(qi), _p(acc)))
        return qk, qi, acc

    def stage_times(self, reset=True) -> dict:
        arr = (C.c_double * 6)()
        self._check(self.L.tf_get_stage_times(self.h, int(reset), arr))
        t = dict(zip(("bbox", "cull", "_2", "alloc", "integrate+finalize", "_5"), arr))
        return {k: v for k, v in t.items() if not k.startswith("_")}

    def kernel_time(self, reset=True):
        ms, n, b = C.c_double(), C.c_int64(), C.c_double()
        self._check(self.L.tf_get_kernel_time(self.h, int(reset), C.byref(ms), C.byref(n), C.byref(b)))
        return ms.value, n.value, b.value
