// tf_host_math.h — per-frame constants computed on the host.
//
// These follow the reference's scalar set-up code op for op (un-fused float math; this file
// is compiled with -ffp-contract=off and without -mfma, like the reference,
// CMakeLists.txt:57-58), so that the kernels start from bit-identical inputs.
#pragma once
#include <cmath>
#include "../../include/texfusion.h"
#include "tf_device.cuh"

namespace tfb {

// (see dot3 in tf_device.cuh)
inline float h_dot3(int l2r, float a0, float b0, float a1, float b1, float a2, float b2) {
  return l2r ? (a0 * b0 + a1 * b1) + a2 * b2 : a0 * b0 + (a1 * b1 + a2 * b2);
}

// tf_pose is column-major camera->world: R(i,j) = m[j*4+i]; Rt(i,j) = m[i*4+j].
inline void pose_split(const tf_pose& p, float R[9], float Rt[9], float t[3]) {
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      R[i * 3 + j] = p.m[j * 4 + i];
      Rt[i * 3 + j] = p.m[i * 4 + j];
    }
  t[0] = p.m[12], t[1] = p.m[13], t[2] = p.m[14];
}

// GetChunkIDsObservedByCamera set-up (Structure/ChunkManager.h:398-470) + bbox constants.
inline void make_cull_params(float res, const tf_truncation& tr, int l2r, const tf_pose& pose, const tf_camera& cam,
                             CullParams& cp) {
  pose_split(pose, cp.R, cp.Rt, cp.t);
  cp.l2r = l2r;
  for (int k = 0; k < 3; k++)
    cp.tau[k] = h_dot3(l2r, cp.Rt[k * 3 + 0], cp.t[0], cp.Rt[k * 3 + 1], cp.t[1], cp.Rt[k * 3 + 2], cp.t[2]);
  for (int k = 0; k < 3; k++)
    for (int i = 0; i < 3; i++) cp.r[k][i] = cp.Rt[i * 3 + k] * 8.0f * res;
  float diag = 8 * res / 2;
  int step = 4;
  float neg_trunc = (float)0.03;
  if (res > 0.01) {  // float vs double literal: 0.01f does not pass
    diag = (float)((double)(8 * res) * std::sqrt(3.0));  // chunkSize*res*sqrt(3): float * double
    step = 1;
    neg_trunc = (float)(0.05 * (double)res / 0.005);
  }
  const float half = res * 0.5f;
  for (int x = 0; x < 2; x++)
    for (int y = 0; y < 2; y++)
      for (int z = 0; z < 2; z++) {
        const float c0 = (float)(x * 8), c1 = (float)(y * 8), c2 = (float)(z * 8);
        const int idx = x + y * 2 + z * 4;
        for (int k = 0; k < 3; k++) {
          const float rc = h_dot3(l2r, cp.Rt[k * 3 + 0], c0, cp.Rt[k * 3 + 1], c1, cp.Rt[k * 3 + 2], c2);
          cp.off_c[idx][k] = rc * res * (float)step + half;
          cp.off_f[idx][k] = rc * res * 1.0f + half;
        }
      }
  // PinholeCamera::GetFx/GetFy/GetCx/GetCy return int (PinholeCamera.h:46-49)
  cp.fx = (float)(int)cam.fx, cp.fy = (float)(int)cam.fy, cp.cx = (float)(int)cam.cx, cp.cy = (float)(int)cam.cy;
  cp.W = cam.width, cp.H = cam.height;
  cp.near_p = cam.near_plane, cp.far_p = cam.far_plane;
  cp.inv_chunk = 1.0f / (8 * res);
  cp.res = res;
  cp.diag = diag;
  cp.diag_step = diag * step;
  cp.dtn_c = neg_trunc + diag * step;
  cp.dtn_f = neg_trunc + diag;
  cp.step = step;
  cp.step_log2 = step == 4 ? 2 : 0;
  cp.trunc = TruncDev{tr.quad, tr.lin, tr.cst, tr.scale, tr.weight};
}

// voxelUpdateSIMD set-up (ProjectionIntegrator.cpp:74-130).
inline void make_frame_dev(const tf_pose& pose, const tf_camera& cam, float res, int flag, const float* depth,
                           const uchar4* rgba, const float* quality, FrameDev& f) {
  float R[9];
  pose_split(pose, R, f.Rt, f.t);
  const float cx = (float)(int)cam.cx, cy = (float)(int)cam.cy;
  f.fx = (float)(int)cam.fx, f.fy = (float)(int)cam.fy;
  f.cxh = (float)((double)cx + 0.5);  // double add, rounded to float by _mm256_set1_ps
  f.cyh = (float)((double)cy + 0.5);
  // Range test of project_safe (tf_device.cuh): a voxel centre is origin + (Rt (x,y,z)) res + res/2 with
  // x,y,z in 0..7, so |centre_k - origin_k| <= (7 (|Rt_k0| + |Rt_k1| + |Rt_k2|) + 0.5) res =: b_k (taken 1 % up
  // for the roundings).  Chunks whose origin depth exceeds z_safe = b_2 + 2^-16 have every cz > 2^-17; the
  // test is switched off (z_safe = inf) when a b_k is not small against 2^20 or the intrinsics are out of range.
  double b[3];
  bool sane = std::fabs((double)f.fx) < 1048576.0 && std::fabs((double)f.fy) < 1048576.0;
  for (int k = 0; k < 3; k++) {
    b[k] = 1.01 * (7.0 * (std::fabs((double)f.Rt[3 * k]) + std::fabs((double)f.Rt[3 * k + 1]) + std::fabs((double)f.Rt[3 * k + 2])) + 0.5) * (double)res;
    sane = sane && b[k] < 1048576.0;  // (false for NaN)
  }
  f.z_safe = sane ? (float)(b[2] + 1.0 / 65536.0) * 1.0001f : INFINITY;
  f.pad0 = 0.0f;
  f.W = cam.width, f.H = cam.height;
  f.near_p = cam.near_plane, f.far_p = cam.far_plane;
  f.flag = flag;
  f.depth = depth, f.rgba = rgba, f.quality = quality;
}

inline void make_group_consts(float res, const tf_truncation& tr, int l2r, GroupParams& gp) {
  gp.res = res;
  gp.l2r = l2r;
  gp.half = res * 0.5f;
  // `sqrt(3.0f) * resolution` binds to ::sqrt(double): double product rounded to float.
  // (spelled out: in a .cu file sqrt(float) would resolve to CUDA's float overload)
  const float diag = (float)(std::sqrt((double)3.0f) * (double)res);
  gp.diag = diag;
  gp.thr_c = (float)((double)(diag / 2) + 0.01);
  gp.trunc = TruncDev{tr.quad, tr.lin, tr.cst, tr.scale, tr.weight};
}

}  // namespace tfb
