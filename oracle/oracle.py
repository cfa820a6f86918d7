"""ctypes binding of the CPU checkers (test infrastructure only):

  impl="port"      oracle/libtexfusion_oracle.so — the AVX2 restatement (tf_oracle.cpp)
  impl="ref"       oracle/_ref/libtexfusion_ref.so — the reference's OWN sources compiled against
                   oracle/eigen_standin (ref_driver.cpp); built here, where /root/reference exists,
                   and shipped prebuilt to the GPU box
  impl="ref_l2r"   the same with left-to-right 3-term products (Eigen 3.2 association)
  impl="ref_chisel" the same driver around the reference's own chisel::Chisel object (Structure/Chisel.h inline
                   methods instead of the restated glue; ref_driver.cpp -DTF_REF_REAL_CHISEL) — for lists below the
                   reference's threading threshold of 1000 chunks
"""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess
import tempfile

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

# QuadraticTruncator(0.0019, 0.00152, 0.001504, 6) + ConstantWeighter(1): GCFusion/MobileFusion.h:215-228
DEFAULT_TRUNC = (0.0019, 0.00152, 0.001504, 6.0, 1.0)


def oracle_lib_path() -> str:
    return os.path.join(_HERE, "libtexfusion_oracle.so")


def build_oracle(force: bool = False) -> str:
    path = oracle_lib_path()
    src = os.path.join(_HERE, "tf_oracle.cpp")
    if force or not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libtexfusion_oracle.so"],
                              stdout=subprocess.DEVNULL)
    return path


class _Cam(C.Structure):
    _fields_ = [("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float),
                ("width", C.c_int32), ("height", C.c_int32), ("near_plane", C.c_float),
                ("far_plane", C.c_float)]


def ref_lib_path(impl: str = "ref") -> str:
    """TF_REF_LIB overrides the library (a build against the real Eigen, tools/ref_golden/);
    it serves as impl "ref", or as "ref_l2r" when TF_REF_ORDER=l2r."""
    ov = os.environ.get("TF_REF_LIB")
    if ov and (impl == "ref_l2r") == (os.environ.get("TF_REF_ORDER") == "l2r"):
        return ov
    return os.path.join(_HERE, "_ref", {"ref": "libtexfusion_ref.so", "ref_l2r": "libtexfusion_ref_l2r.so",
                                        "ref_chisel": "libtexfusion_ref_chisel.so"}[impl])


def have_ref(impl: str = "ref") -> bool:
    return os.path.exists(ref_lib_path(impl))


def build_ref() -> bool:
    """(Re)builds oracle/_ref from the reference tree when it is present; True if the library exists."""
    if os.path.isdir("/root/reference/Structure"):
        subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL)
    return have_ref()


_REF_LIBS = {}
_REF_TMP = None


def _ref_lib(impl: str, res: float):
    """One loaded copy of the reference library per (impl, resolution): ChunkManager::GetIDAt and the
    mesher keep resolution-dependent function-local statics (Structure/ChunkManager.h:197-203)."""
    global _REF_TMP
    key = (impl, float(np.float32(res)))
    if key not in _REF_LIBS:
        src = ref_lib_path(impl)
        if not os.path.exists(src):
            raise FileNotFoundError(f"{src} not built (make -C oracle ref needs /root/reference)")
        if _REF_TMP is None:
            _REF_TMP = tempfile.mkdtemp(prefix="tf_ref_")
        dst = os.path.join(_REF_TMP, f"libtexfusion_{impl}_{len(_REF_LIBS)}.so")
        shutil.copyfile(src, dst)
        L = C.CDLL(dst)
        _declare(L)
        L.tfo_impl.restype = C.c_char_p
        L.tfo_mesh_chunks.argtypes = [C.c_void_p] * 9 + [C.c_int64, C.c_int64]
        _REF_LIBS[key] = L
    return _REF_LIBS[key]


def _lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build_oracle())
        _declare(L, full=True)
        L.tfo_impl.restype = C.c_char_p
        L.tfo_set_dot3_order.argtypes = [C.c_int]
        _LIB = L
    return _LIB


def set_dot3_order(left_to_right: bool):
    """Association of the restatement's 3-term products (process-wide; impl "port" only — the
    reference build has one library per order: impl "ref" / "ref_l2r")."""
    _lib().tfo_set_dot3_order(1 if left_to_right else 0)


def _declare(L, full=False):
    if True:
        vp, f32p, u8p, i32p = C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_uint8), C.POINTER(C.c_int32)
        L.tfo_create.restype = vp
        L.tfo_create.argtypes = [C.c_float, f32p, C.c_int]
        L.tfo_destroy.argtypes = [vp]
        L.tfo_reset.argtypes = [vp]
        L.tfo_threads_used.argtypes = [vp]
        L.tfo_threads_used.restype = C.c_int
        L.tfo_boundary_ids.argtypes = [vp, vp, vp, C.POINTER(_Cam), i32p, i32p]
        L.tfo_observed_ids.restype = C.c_int64
        L.tfo_observed_ids.argtypes = [vp, vp, vp, C.POINTER(_Cam), vp, C.c_int64]
        L.tfo_prepare.restype = C.c_int64
        L.tfo_prepare.argtypes = [vp, vp, vp, C.POINTER(_Cam), vp, vp, C.c_int64]
        L.tfo_integrate.restype = C.c_int
        L.tfo_integrate.argtypes = [vp, vp, vp, vp, vp, C.POINTER(_Cam), vp, C.c_int64, C.c_int, C.c_int, vp, vp]
        L.tfo_finalize.restype = C.c_int64
        L.tfo_finalize.argtypes = [vp, vp, C.c_int64, vp, vp, vp]
        L.tfo_integrate_frame.restype = C.c_int64
        L.tfo_integrate_frame.argtypes = [vp, vp, vp, vp, vp, C.POINTER(_Cam), C.c_int, C.POINTER(C.c_int64)]
        L.tfo_has_chunk.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32]
        L.tfo_chunk_count.restype = C.c_int64
        L.tfo_chunk_count.argtypes = [vp]
        L.tfo_list_chunks.restype = C.c_int64
        L.tfo_list_chunks.argtypes = [vp, vp, C.c_int64]
        L.tfo_download_chunks.argtypes = [vp, vp, C.c_int64, vp, vp, vp]
        L.tfo_get_observation.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32, C.c_int, f32p]
        L.tfo_meshes_to_update.restype = C.c_int64
        L.tfo_meshes_to_update.argtypes = [vp, vp, C.c_int64]
        L.tfo_clear_meshes_to_update.argtypes = [vp]
        L.tfo_retract_observations.argtypes = [vp, vp, C.c_int64, C.c_int]
        L.tfo_truncation_distance.restype = C.c_float
        L.tfo_truncation_distance.argtypes = [f32p, C.c_float]
        L.tfo_centroids.argtypes = [vp, vp, vp]
        if not full:
            return
        L.tfo_atlas_patch_size.argtypes = [vp, i32p, i32p]
        L.tfo_atlas_alloc_slot.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_uint64)]
        L.tfo_atlas_update.argtypes = [vp, C.c_uint64, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        L.tfo_atlas_download.argtypes = [vp, C.c_uint64, C.c_uint64, vp]
        L.tfo_patch_texcoords.argtypes = [vp, vp, vp, C.POINTER(_Cam), C.c_int64, vp, vp, vp, vp, vp, vp]


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _pose(pose) -> np.ndarray:
    """4x4 camera->world (row-major NumPy) -> 16 floats column-major (Eigen layout)."""
    return np.ascontiguousarray(np.asarray(pose, np.float32).T).reshape(16)


def _cam(cam) -> _Cam:
    return _Cam(cam.fx, cam.fy, cam.cx, cam.cy, cam.width, cam.height, cam.near, cam.far)


def truncation_distance(z: float, trunc=DEFAULT_TRUNC, impl: str = "port") -> float:
    t = (C.c_float * 5)(*trunc)
    L = _lib() if impl == "port" else _ref_lib(impl, 0.0)
    return float(L.tfo_truncation_distance(t, C.c_float(z)))


_PATCH_REF = None


def have_patch_ref() -> bool:
    return os.path.exists(os.path.join(_HERE, "_ref", "libtexfusion_ref_patch.so"))


def _patch_lib(impl: str):
    """impl "port": the restatement in tf_oracle.cpp; "ref": the reference's own Structure/Patch.cpp functions
    (oracle/ref_patch_driver.cpp)."""
    global _PATCH_REF
    if impl == "port":
        return _lib()
    if _PATCH_REF is None:
        L = C.CDLL(os.path.join(_HERE, "_ref", "libtexfusion_ref_patch.so"))
        vp = C.c_void_p
        L.tfo_patch_texcoords.argtypes = [vp, vp, vp, C.POINTER(_Cam), C.c_int64, vp, vp, vp, vp, vp, vp]
        _PATCH_REF = L
    return _PATCH_REF


def patch_texcoords(rgb, depth, world_to_camera, cam, offsets, vertices, colors, impl: str = "port"):
    """Patch::CalculateTexCoords (Structure/Patch.cpp:40-108) for a batch of meshes on the CPU."""
    rgb = np.ascontiguousarray(rgb, np.uint8)
    d = np.ascontiguousarray(depth, np.float32)
    off = np.ascontiguousarray(offsets, np.int64)
    v = np.ascontiguousarray(vertices, np.float32).reshape(-1, 3)
    c = np.ascontiguousarray(colors, np.float32).reshape(-1, 3)
    n = len(off) - 1
    tc = np.empty((len(v), 2), np.float32)
    col = np.empty((len(v), 3), np.float32)
    res = np.empty((max(n, 1), 6), np.int32)
    _patch_lib(impl).tfo_patch_texcoords(_p(rgb), _p(d), _p(_pose(world_to_camera)), C.byref(_cam(cam)), n, _p(off), _p(v), _p(c),
                                         _p(tc), _p(col), _p(res))
    return tc, col, res[:n]


class OracleMap:
    """CPU restatement of chisel::Chisel's fusion path (prepare / integrate / finalize / atlas)."""

    def __init__(self, res: float, trunc=DEFAULT_TRUNC, threads: int = 1, impl: str = "port"):
        self.impl = impl
        self.L = _lib() if impl == "port" else _ref_lib(impl, res)
        self.res = float(np.float32(res))
        self.h = self.L.tfo_create(C.c_float(res), (C.c_float * 5)(*trunc), threads)

    def close(self):
        if self.h:
            self.L.tfo_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def threads_used(self) -> int:
        return self.L.tfo_threads_used(self.h)

    def reset(self):
        self.L.tfo_reset(self.h)

    def boundary_ids(self, depth, pose, cam):
        mn, mx = (C.c_int32 * 3)(), (C.c_int32 * 3)()
        d = np.ascontiguousarray(depth, np.float32)
        self.L.tfo_boundary_ids(self.h, _p(d), _p(_pose(pose)), C.byref(_cam(cam)), mn, mx)
        return np.array(mn[:], np.int32), np.array(mx[:], np.int32)

    def observed_ids(self, depth, pose, cam, cap=1 << 20):
        d = np.ascontiguousarray(depth, np.float32)
        ids = np.empty((cap, 3), np.int32)
        n = self.L.tfo_observed_ids(self.h, _p(d), _p(_pose(pose)), C.byref(_cam(cam)), _p(ids), cap)
        assert n <= cap
        return ids[:n].copy()

    def prepare(self, depth, pose, cam, cap=1 << 20):
        d = np.ascontiguousarray(depth, np.float32)
        ids = np.empty((cap, 3), np.int32)
        new = np.empty(cap, np.uint8)
        n = self.L.tfo_prepare(self.h, _p(d), _p(_pose(pose)), C.byref(_cam(cam)), _p(ids), _p(new), cap)
        if n < 0:
            raise RuntimeError("oracle prepare: capacity")
        return ids[:n].copy(), new[:n].copy()

    def integrate(self, depth, rgba, quality, pose, cam, ids, flag, keyframe_id=-1, needs_update=None):
        d = np.ascontiguousarray(depth, np.float32)
        c = None if rgba is None else np.ascontiguousarray(rgba, np.uint8)
        q = None if quality is None else np.ascontiguousarray(quality, np.float32)
        ids = np.ascontiguousarray(ids, np.int32).reshape(-1, 3)
        n = len(ids)
        nu = np.zeros(n, np.uint8) if needs_update is None else np.ascontiguousarray(needs_update, np.uint8)
        qo = np.zeros(n, np.float32)
        rc = self.L.tfo_integrate(self.h, _p(d), _p(c), _p(q), _p(_pose(pose)), C.byref(_cam(cam)),
                                  _p(ids), n, int(flag), int(keyframe_id), _p(nu), _p(qo))
        if rc != 0:
            raise KeyError("oracle integrate: chunk not found")
        return nu, qo

    def finalize(self, ids, needs_update, is_new):
        ids = np.ascontiguousarray(ids, np.int32).reshape(-1, 3)
        n = len(ids)
        valid = np.empty((max(n, 1), 3), np.int32)
        nv = self.L.tfo_finalize(self.h, _p(ids), n, _p(np.ascontiguousarray(needs_update, np.uint8)),
                                 _p(np.ascontiguousarray(is_new, np.uint8)), _p(valid))
        return valid[:nv].copy()

    def integrate_frame(self, depth, rgba, quality, pose, cam, keyframe_id=-1):
        d = np.ascontiguousarray(depth, np.float32)
        c = None if rgba is None else np.ascontiguousarray(rgba, np.uint8)
        q = None if quality is None else np.ascontiguousarray(quality, np.float32)
        nupd = C.c_int64(0)
        n = self.L.tfo_integrate_frame(self.h, _p(d), _p(c), _p(q), _p(_pose(pose)), C.byref(_cam(cam)),
                                       int(keyframe_id), C.byref(nupd))
        return int(n), int(nupd.value)

    def has_chunk(self, id3) -> bool:
        return bool(self.L.tfo_has_chunk(self.h, int(id3[0]), int(id3[1]), int(id3[2])))

    def chunk_count(self) -> int:
        return int(self.L.tfo_chunk_count(self.h))

    def list_chunks(self) -> np.ndarray:
        n = self.chunk_count()
        out = np.empty((max(n, 1), 3), np.int32)
        self.L.tfo_list_chunks(self.h, _p(out), n)
        return out[:n]

    def download_chunks(self, ids):
        ids = np.ascontiguousarray(ids, np.int32).reshape(-1, 3)
        n = len(ids)
        sdf = np.empty((n, 512), np.float32)
        w = np.empty((n, 512), np.float32)
        col = np.empty((n, 2048), np.uint16)
        if self.L.tfo_download_chunks(self.h, _p(ids), n, _p(sdf), _p(w), _p(col)) != 0:
            raise KeyError("oracle download: chunk not found")
        return sdf, w, col

    def observation(self, id3, keyframe):
        v = C.c_float(0)
        rc = self.L.tfo_get_observation(self.h, int(id3[0]), int(id3[1]), int(id3[2]), int(keyframe), C.byref(v))
        return float(v.value) if rc == 1 else None

    def meshes_to_update(self) -> np.ndarray:
        n = self.L.tfo_meshes_to_update(self.h, None, 0)
        out = np.empty((max(n, 1), 3), np.int32)
        self.L.tfo_meshes_to_update(self.h, _p(out), n)
        return out[:n]

    def clear_meshes_to_update(self):
        self.L.tfo_clear_meshes_to_update(self.h)

    def retract_observations(self, ids, keyframe):
        ids = np.ascontiguousarray(ids, np.int32).reshape(-1, 3)
        self.L.tfo_retract_observations(self.h, _p(ids), len(ids), int(keyframe))

    def centroids(self, pose) -> np.ndarray:
        out = np.empty((3, 512), np.float32)
        self.L.tfo_centroids(self.h, _p(_pose(pose)), _p(out))
        return out

    def mesh_chunks(self, ids):
        """ChunkManager::GenerateMeshEfficient per chunk (impl "ref" only): offsets + concatenated
        vertices / normals / colours / indices."""
        ids = np.ascontiguousarray(ids, np.int32).reshape(-1, 3)
        n = len(ids)
        voff = np.zeros(n + 1, np.int64)
        ioff = np.zeros(n + 1, np.int64)
        self.L.tfo_mesh_chunks(self.h, _p(ids), n, _p(voff), _p(ioff), None, None, None, None, 0, 0)
        nv, ni = int(voff[-1]), int(ioff[-1])
        vert = np.empty((max(nv, 1), 3), np.float32)
        norm = np.empty((max(nv, 1), 3), np.float32)
        col = np.empty((max(nv, 1), 3), np.float32)
        idx = np.empty(max(ni, 1), np.int32)
        rc = self.L.tfo_mesh_chunks(self.h, _p(ids), n, _p(voff), _p(ioff), _p(vert), _p(norm), _p(col), _p(idx), nv, ni)
        assert rc == 0
        return voff, ioff, vert[:nv], norm[:nv], col[:nv], idx[:ni]

    # atlas
    def atlas_patch_size(self):
        w, h = C.c_int32(), C.c_int32()
        self.L.tfo_atlas_patch_size(self.h, C.byref(w), C.byref(h))
        return w.value, h.value

    def atlas_alloc_slot(self, id3) -> int:
        loc = C.c_uint64()
        rc = self.L.tfo_atlas_alloc_slot(self.h, int(id3[0]), int(id3[1]), int(id3[2]), C.byref(loc))
        if rc != 0:
            raise OverflowError("No enough space for texture storage.")
        return int(loc.value)

    def atlas_update(self, texloc, rgb, box):
        rgb = np.ascontiguousarray(rgb, np.uint8)
        h, w, _ = rgb.shape
        rc = self.L.tfo_atlas_update(self.h, C.c_uint64(texloc), _p(rgb), w, h, *[int(v) for v in box])
        if rc != 0:
            raise ValueError("oracle atlas_update: bad box")

    def atlas_download(self, hot_start, hot_end) -> np.ndarray:
        out = np.empty((hot_end - hot_start) * 3, np.uint8)
        self.L.tfo_atlas_download(self.h, C.c_uint64(hot_start), C.c_uint64(hot_end), _p(out))
        return out
