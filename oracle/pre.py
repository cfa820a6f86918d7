"""ctypes binding of the CPU checkers of the frame pre-processing step (TEST INFRASTRUCTURE ONLY):

  impl="port"      oracle/libtexfusion_pre_oracle.so — scalar restatement (tf_pre_oracle.cpp)
  impl="ref"       oracle/_ref/libtexfusion_ref_pre.so — the reference's own BasicAPI.cpp loops compiled
                   between stand-ins (ref_pre_driver.cpp); built where /root/reference exists, shipped
                   prebuilt to the GPU box
  impl="ref_l2r"   the same with Eigen 3.2's association in the two dot products

Every function takes and returns numpy arrays; images are (H, W) float32 / uint8, normals (3, H, W).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {}


def lib_path(impl: str = "port") -> str:
    if impl == "port":
        return os.path.join(_HERE, "libtexfusion_pre_oracle.so")
    return os.path.join(_HERE, "_ref", "libtexfusion_ref_pre.so" if impl == "ref" else "libtexfusion_ref_pre_l2r.so")


def have(impl: str) -> bool:
    return os.path.exists(lib_path(impl))


def build(force: bool = False) -> str:
    path, src = lib_path("port"), os.path.join(_HERE, "tf_pre_oracle.cpp")
    if force or not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libtexfusion_pre_oracle.so"], stdout=subprocess.DEVNULL)
    return path


def _lib(impl: str):
    if impl not in _LIBS:
        if impl == "port":
            build()
        L = C.CDLL(lib_path(impl))
        vp, f, i = C.c_void_p, C.c_float, C.c_int
        cam = [f, f, f, f]
        L.tfp_impl.restype = C.c_char_p
        L.tfp_normal_map.argtypes = [vp, vp, i, i] + cam
        L.tfp_refine_depth_by_normal.argtypes = [vp, vp, i, i] + cam
        L.tfp_refine_keyframe.argtypes = [vp, vp, vp, vp, i, i] + cam
        L.tfp_refine_newframe.argtypes = [vp, vp, vp, i, i] + cam
        L.tfp_color_valid.argtypes = [vp, vp, i, i] + cam
        L.tfp_color_quality.argtypes = [vp, vp, vp, vp, i, i] + cam
        L.tfp_gray.argtypes = [vp, vp, i, i]
        L.tfp_sobel11.argtypes = [vp, vp, i, i]
        L.tfp_rsqrt_bits.argtypes = [C.c_uint32]
        L.tfp_rsqrt_bits.restype = C.c_uint32
        L.tfp_host_rsqrt_matches.restype = C.c_int
        if impl == "port":
            L.tfp_set_dot3_order.argtypes = [C.c_int]
        _LIBS[impl] = L
    return _LIBS[impl]


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class Pre:
    """The six loops of main.cpp:117-147 on numpy images.  cam = (fx, fy, cx, cy) as floats (NOT truncated)."""

    def __init__(self, impl: str = "port", left_to_right: bool = False):
        if impl == "ref" and left_to_right:
            impl = "ref_l2r"
        self.impl = impl
        self.L = _lib(impl)
        if impl == "port":
            self.L.tfp_set_dot3_order(1 if left_to_right else 0)

    def describe(self) -> str:
        return self.L.tfp_impl().decode()

    def host_rsqrt_matches(self) -> bool:
        return bool(self.L.tfp_host_rsqrt_matches())

    def rsqrt_bits(self, b: int) -> int:
        return int(self.L.tfp_rsqrt_bits(b))

    def normal_map(self, depth, cam):
        d = _f32(depth)
        H, W = d.shape
        n = np.zeros((3, H, W), np.float32)
        self.L.tfp_normal_map(_p(d), _p(n), W, H, *cam)
        return n

    def refine_depth_by_normal(self, normal, depth, cam):
        n, d = _f32(normal).copy(), _f32(depth).copy()
        H, W = d.shape
        self.L.tfp_refine_depth_by_normal(_p(n), _p(d), W, H, *cam)
        return n, d

    def refine_keyframe(self, kf_depth, kf_weight, new_depth, T_ref_to_new, cam):
        d, w, nd = _f32(kf_depth).copy(), _f32(kf_weight).copy(), _f32(new_depth)
        T = _f32(np.asarray(T_ref_to_new)[:3, :4])
        H, W = d.shape
        self.L.tfp_refine_keyframe(_p(d), _p(w), _p(nd), _p(T), W, H, *cam)
        return d, w

    def refine_newframe(self, kf_depth, new_depth, T_new_to_ref, cam):
        d, nd = _f32(kf_depth), _f32(new_depth).copy()
        T = _f32(np.asarray(T_new_to_ref)[:3, :4])
        H, W = d.shape
        self.L.tfp_refine_newframe(_p(d), _p(nd), _p(T), W, H, *cam)
        return nd

    def color_valid(self, normal, cam):
        n = _f32(normal)
        _, H, W = n.shape
        f = np.zeros((H, W), np.uint8)
        self.L.tfp_color_valid(_p(n), _p(f), W, H, *cam)
        return f

    def color_quality(self, depth, normal, rgb, cam):
        d, n = _f32(depth), _f32(normal)
        c = np.ascontiguousarray(rgb, dtype=np.uint8)
        H, W = d.shape
        q = np.zeros((H, W), np.float32)
        self.L.tfp_color_quality(_p(d), _p(n), _p(c), _p(q), W, H, *cam)
        return q

    def gray(self, rgb):
        c = np.ascontiguousarray(rgb, dtype=np.uint8)
        H, W, _ = c.shape
        g = np.zeros((H, W), np.uint8)
        self.L.tfp_gray(_p(c), _p(g), W, H)
        return g

    def sobel11(self, gray):
        g = np.ascontiguousarray(gray, dtype=np.uint8)
        H, W = g.shape
        o = np.zeros((H, W), np.float32)
        self.L.tfp_sobel11(_p(g), _p(o), W, H)
        return o


def relative_transform(pose_a, pose_b):
    """inverse(pose_a) * pose_b as float64 4x4 (poses are camera -> world): maps b's camera frame into a's."""
    return np.linalg.inv(np.asarray(pose_a, np.float64)) @ np.asarray(pose_b, np.float64)
