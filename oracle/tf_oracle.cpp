// tf_oracle.cpp — CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE)
//
// A from-scratch CPU restatement of the fusion hot path of THU-luvision/TextureFusion,
// written against the behaviour of the reference sources cited at each function
// (paths relative to the reference tree).  It exists only to check the CUDA path
// (tests/, __graft_entry__.smoke) and to serve as the CPU baseline of bench.py;
// nothing under texturefusion_b200/ may link, import or call it.
//
// HOW IT IS PINNED: the reference ships no tests, golden vectors or fixtures for this path.  Its
// fusion sources do compile here once Eigen (absent from this image) is replaced by the stand-in
// in oracle/eigen_standin: oracle/_ref/libtexfusion_ref.so (oracle/Makefile, ref_driver.cpp) is the
// reference's own ProjectionIntegrator.cpp / ChunkManager.{h,cpp} / Chunk / truncator / weighter /
// camera code, and tests/test_ref_cpu.py checks this restatement against it bit for bit (chunk
// lists in order, flags, quality sums, every voxel) at 40/20/10/5 mm; tests/golden/*.npz are
// generated from that library.  What stays unverifiable here is the one thing the stand-in has to
// choose: the association order of Eigen's 3-term products (runtime switch below;
// tools/ref_golden/ is the recipe for a machine that has Eigen).  Also checked against an
// independent scalar restatement (oracle/scalar_ref.py) and hand-computed known-answer cases.
// tfo_patch_texcoords is pinned by oracle/_ref/libtexfusion_ref_patch.so (the reference's own
// Patch::CalculateTexCoords / bilinear / bilinear_depth, oracle/ref_patch_driver.cpp; tests/test_patch.py).
//
// Arithmetic rules it follows (see DESIGN.md "Arithmetic contract"):
//   * the reference is built with -mavx2 and WITHOUT -mfma (CMakeLists.txt:57-58): every
//     float op is a separate IEEE binary32 op; this file is compiled with
//     -mno-fma -ffp-contract=off and uses the same AVX2 intrinsics for the vector parts.
//   * Eigen fixed-size 3-vector inner products (coefficient-based lazy product ->
//     redux_novec_unroller<0,3>) associate as a0*b0 + (a1*b1 + a2*b2) from Eigen 3.3 on;
//     Eigen 3.2 accumulates (a0*b0 + a1*b1) + a2*b2.  tfo_set_dot3_order switches (process-wide).
//   * PinholeCamera::GetFx/GetFy/GetCx/GetCy return int
//     (3rd_party/open_chisel/camera/PinholeCamera.h:46-49).
#include <immintrin.h>
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <thread>
#include <unordered_map>
#include <vector>

namespace {

struct Id3 {
  int32_t x, y, z;
  bool operator==(const Id3& o) const { return x == o.x && y == o.y && z == o.z; }
};

// chisel::ChunkHasher — Structure/ChunkManager.h:44-53 (ints sign-extend to size_t).
struct IdHash {
  size_t operator()(const Id3& k) const {
    return ((size_t)(int64_t)k.x * (size_t)73856093) ^ ((size_t)(int64_t)k.y * (size_t)19349663) ^
           ((size_t)(int64_t)k.z * (size_t)83492791);
  }
};

struct Cam {
  float fx, fy, cx, cy;
  int32_t width, height;
  float near_plane, far_plane;
};

struct Trunc {
  float quad, lin, cst, scale, weight;
};

// chisel::Chunk + DistVoxel + ColorVoxel — 3rd_party/open_chisel/geometry/Chunk.cpp:38-68,
// ColorVoxel.cpp:26-31.  512 voxels, voxel index (z*8+y)*8+x (Chunk.h:91-93).
struct Chunk {
  Id3 id;
  float origin[3];
  alignas(32) float sdf[512];
  alignas(32) float weight[512];
  alignas(32) uint16_t color[2048];
  std::map<int, float> observations;
};

struct Patch {
  uint64_t texloc;
};

constexpr int kAtlasDim = 96 * 72 * 2;  // MAX_PATCH_WIDTH == MAX_PATCH_HEIGHT, Structure/Atlas.h:29-30

struct Map {
  float res;
  Trunc trunc;
  int threads;  // 0 = reference policy (hardware_concurrency()-2), 1 = serial
  std::unordered_map<Id3, std::unique_ptr<Chunk>, IdHash> chunks;
  std::unordered_map<Id3, bool, IdHash> meshes_to_update;
  // atlas
  std::vector<uint8_t> atlas;  // lazily allocated
  uint64_t loc_next = 0;
  int patch_w = 0, patch_h = 0;
  std::unordered_map<Id3, Patch, IdHash> patches;
  // scratch: per-frame centroid tables (ProjectionIntegrator::centroids_simd{0,1,2})
  alignas(32) float cen[3][512];
};

int g_dot3_l2r = 0;  // tfo_set_dot3_order
inline float dot3(float a0, float b0, float a1, float b1, float a2, float b2) {
  return g_dot3_l2r ? (a0 * b0 + a1 * b1) + a2 * b2 : a0 * b0 + (a1 * b1 + a2 * b2);
}

// pose: column-major 4x4, camera->world.  R(i,j) = m[j*4+i]; Rt(i,j) = R(j,i).
inline float R(const float* m, int i, int j) { return m[j * 4 + i]; }
inline float Rt(const float* m, int i, int j) { return m[i * 4 + j]; }

// Rt * v for a fixed-size 3-vector.
inline void rt_mul(const float* m, const float v[3], float out[3]) {
  for (int k = 0; k < 3; k++) out[k] = dot3(Rt(m, k, 0), v[0], Rt(m, k, 1), v[1], Rt(m, k, 2), v[2]);
}

// QuadraticTruncator::GetTruncationDistance — 3rd_party/open_chisel/truncation/QuadraticTruncator.h:45-48.
// float*pow(float,int) promotes to double; lin*z stays float; std::abs(double); *scale in double.
inline float truncation_distance(const Trunc& t, float z) {
  double v = (double)t.quad * std::pow((double)z, 2) + (double)(t.lin * z) + (double)t.cst;
  return (float)(std::fabs(v) * (double)t.scale);
}

// ConstantWeighter::GetWeight — 3rd_party/open_chisel/weighting/ConstantWeighter.h:43-46.
inline float weight_of(const Trunc& t, float trunc) { return t.weight / (2 * trunc); }

struct Intr {
  float fx, fy, cx, cy;  // float(int(.)) — PinholeCamera.h:46-49
};
inline Intr truncated(const Cam& c) {
  return Intr{(float)(int)c.fx, (float)(int)c.fy, (float)(int)c.cx, (float)(int)c.cy};
}

// Chisel::bufferIntegratorSIMDCentroids — Structure/Chisel.cpp:52-110.
void buffer_centroids(Map& m, const float* pose) {
  const float half = m.res * 0.5f;
  int i = 0;
  for (int z = 0; z < 8; z++)
    for (int y = 0; y < 8; y++)
      for (int x = 0; x < 8; x++, i++) {
        float v[3] = {(float)x, (float)y, (float)z}, r[3];
        rt_mul(pose, v, r);
        for (int k = 0; k < 3; k++) m.cen[k][i] = r[k] * m.res + half;
      }
}

// ChunkManager::GetIDAt — Structure/ChunkManager.h:197-207.
inline Id3 id_at(const Map& m, const float p[3]) {
  const float f = 1.0f / (8 * m.res);
  return Id3{(int)std::floor(p[0] * f), (int)std::floor(p[1] * f), (int)std::floor(p[2] * f)};
}

// ChunkManager::findCubeCornerByMat / GetBoundaryChunkID — Structure/ChunkManager.h:303-378.
void boundary_ids(const Map& m, const float* depth, const Cam& cam, const float* pose, Id3& max_id,
                  Id3& min_id) {
  const Intr in = truncated(cam);
  const int W = cam.width, H = cam.height;
  __m256 mx[3], mn[3];
  for (int k = 0; k < 3; k++) {
    mx[k] = _mm256_set1_ps(-1e8);
    mn[k] = _mm256_set1_ps(1e8);
  }
  const __m256 inc = _mm256_set_ps(7, 6, 5, 4, 3, 2, 1, 0);
  const __m256 cx = _mm256_set1_ps(in.cx), cy = _mm256_set1_ps(in.cy);
  const __m256 fx = _mm256_set1_ps(in.fx), fy = _mm256_set1_ps(in.fy);
  for (int i = 0; i < H; i++) {
    for (int j = 0; j < W; j += 8) {
      __m256 d = _mm256_add_ps(_mm256_loadu_ps(depth + i * W + j), _mm256_set1_ps(0.2));
      __m256 x = _mm256_add_ps(inc, _mm256_set1_ps((float)j));
      __m256 y = _mm256_set1_ps((float)i);
      __m256 X = _mm256_mul_ps(_mm256_div_ps(_mm256_sub_ps(x, cx), fx), d);
      __m256 Y = _mm256_mul_ps(_mm256_div_ps(_mm256_sub_ps(y, cy), fy), d);
      for (int k = 0; k < 3; k++) {
        __m256 v = _mm256_add_ps(
            _mm256_add_ps(_mm256_add_ps(_mm256_mul_ps(_mm256_set1_ps(R(pose, k, 0)), X),
                                        _mm256_mul_ps(_mm256_set1_ps(R(pose, k, 1)), Y)),
                          _mm256_mul_ps(_mm256_set1_ps(R(pose, k, 2)), d)),
            _mm256_set1_ps(pose[12 + k]));
        mx[k] = _mm256_max_ps(v, mx[k]);
        mn[k] = _mm256_min_ps(v, mn[k]);
      }
    }
  }
  float hi[3] = {-1e8f, -1e8f, -1e8f}, lo[3] = {1e8f, 1e8f, 1e8f};
  for (int k = 0; k < 3; k++) {
    alignas(32) float a[8], b[8];
    _mm256_store_ps(a, mx[k]);
    _mm256_store_ps(b, mn[k]);
    for (int l = 0; l < 8; l++) {
      hi[k] = fmaxf(hi[k], a[l]);
      lo[k] = fminf(lo[k], b[l]);
    }
  }
  max_id = id_at(m, hi);
  min_id = id_at(m, lo);
}

// ChunkManager::CheckCornerIntersectingSIMD — Structure/ChunkManager.h:561-636.
inline bool corner_test(const Cam& cam, const Intr& in, const float o[3], const float* depth,
                        float dtp, float dtn, const __m256 off[3]) {
  const __m256 o2 = _mm256_set1_ps(o[2]);
  const __m256 c0 = _mm256_add_ps(_mm256_set1_ps(o[0]), off[0]);
  const __m256 c1 = _mm256_add_ps(_mm256_set1_ps(o[1]), off[1]);
  const __m256 c2 = _mm256_add_ps(o2, off[2]);
  const __m256 pu =
      _mm256_add_ps(_mm256_mul_ps(_mm256_div_ps(c0, c2), _mm256_set1_ps(in.fx)), _mm256_set1_ps(in.cx));
  const __m256 pv =
      _mm256_add_ps(_mm256_mul_ps(_mm256_div_ps(c1, c2), _mm256_set1_ps(in.fy)), _mm256_set1_ps(in.cy));
  const __m256i u = _mm256_cvtps_epi32(pu), v = _mm256_cvtps_epi32(pv);
  const __m256 depth_ok = _mm256_and_ps(_mm256_cmp_ps(o2, _mm256_set1_ps(cam.near_plane), _CMP_GT_OS),
                                        _mm256_cmp_ps(_mm256_set1_ps(cam.far_plane), o2, _CMP_GT_OS));
  __m256i valid = _mm256_and_si256(_mm256_cmpgt_epi32(u, _mm256_set1_epi32(1)),
                                   _mm256_cmpgt_epi32(_mm256_set1_epi32(cam.width - 1), u));
  valid = _mm256_and_si256(valid, _mm256_cmpgt_epi32(v, _mm256_set1_epi32(1)));
  valid = _mm256_and_si256(valid, _mm256_cmpgt_epi32(_mm256_set1_epi32(cam.height - 1), v));
  const __m256 validf = _mm256_castsi256_ps(valid);
  if (_mm256_testz_ps(validf, validf)) return false;
  const __m256i pix = _mm256_add_epi32(_mm256_mullo_epi32(v, _mm256_set1_epi32(cam.width)), u);
  const __m256 d = _mm256_mask_i32gather_ps(_mm256_set1_ps(0.0f), depth, pix, validf, 4);
  const __m256 sd = _mm256_sub_ps(d, c2);
  __m256 hit = _mm256_and_ps(_mm256_cmp_ps(sd, _mm256_set1_ps(-dtn), _CMP_GT_OS),
                             _mm256_cmp_ps(_mm256_set1_ps(dtp), sd, _CMP_GT_OS));
  hit = _mm256_and_ps(_mm256_and_ps(validf, hit), depth_ok);
  return !_mm256_testz_ps(hit, hit);
}

// ChunkManager::GetChunkIDsObservedByCamera — Structure/ChunkManager.h:380-559.
void observed_ids(const Map& m, const float* depth, const Cam& cam, const float* pose,
                  std::vector<Id3>& out) {
  Id3 min_id, max_id;
  boundary_ids(m, depth, cam, pose, max_id, min_id);
  const float res = m.res;
  float diag = 8 * res / 2;
  int step = 4;
  float neg_trunc = 0.03;
  if (res > 0.01) {  // float vs double literal, as the reference
    diag = 8 * res * sqrt(3);
    step = 1;
    neg_trunc = 0.05 * res / 0.005;
  }
  const Intr in = truncated(cam);
  const float t[3] = {pose[12], pose[13], pose[14]};
  float tau[3];
  rt_mul(pose, t, tau);  // rotation * cameraPose.translation()
  float r[3][3];         // r[k] = Rt.col(k) * 8 * res
  for (int k = 0; k < 3; k++)
    for (int i = 0; i < 3; i++) r[k][i] = Rt(pose, i, k) * 8.0f * res;
  const float half = res * 0.5f;
  alignas(32) float coarse[3][8], fine[3][8];
  for (int x = 0; x < 2; x++)
    for (int y = 0; y < 2; y++)
      for (int z = 0; z < 2; z++) {
        float cur[3] = {(float)(x * 8), (float)(y * 8), (float)(z * 8)}, rc[3];
        rt_mul(pose, cur, rc);
        const int idx = x + y * 2 + z * 4;
        for (int k = 0; k < 3; k++) {
          coarse[k][7 - idx] = rc[k] * res * (float)step + half;  // _mm256_set_ps reverses lanes
          fine[k][7 - idx] = rc[k] * res * 1.0f + half;
        }
      }
  __m256 off_c[3], off_f[3];
  for (int k = 0; k < 3; k++) {
    off_c[k] = _mm256_load_ps(coarse[k]);
    off_f[k] = _mm256_load_ps(fine[k]);
  }
  float ox[3], oy[3], o[3];
  for (int x = min_id.x - 1; x <= max_id.x + 1; x += step) {
    for (int k = 0; k < 3; k++) ox[k] = r[0][k] * (float)x - tau[k];
    for (int y = min_id.y - 1; y <= max_id.y + 1; y += step) {
      for (int k = 0; k < 3; k++) oy[k] = ox[k] + r[1][k] * (float)y;
      for (int z = min_id.z - 1; z <= max_id.z + 1; z += step) {
        for (int k = 0; k < 3; k++) o[k] = oy[k] + (float)z * r[2][k];
        float trunc = truncation_distance(m.trunc, o[2]);
        float dtp = trunc + diag * step;
        float dtn = neg_trunc + diag * step;
        if (!corner_test(cam, in, o, depth, dtp, dtn, off_c)) continue;
        for (int i = x; i < x + step; i++)
          for (int j = y; j < y + step; j++)
            for (int k2 = z; k2 < z + step; k2++) {
              float org[3] = {(float)(i * 8) * res, (float)(j * 8) * res, (float)(k2 * 8) * res}, oc[3];
              rt_mul(pose, org, oc);
              for (int k = 0; k < 3; k++) oc[k] = oc[k] - tau[k];
              float tr = truncation_distance(m.trunc, oc[2]);
              if (corner_test(cam, in, oc, depth, tr + diag, neg_trunc + diag, off_f))
                out.push_back(Id3{i, j, k2});
            }
      }
    }
  }
}

// ChunkManager::CreateChunk — Structure/ChunkManager.cpp:266-270; Chunk ctor Chunk.cpp:38-53.
Chunk* create_chunk(Map& m, const Id3& id) {
  std::unique_ptr<Chunk> c(new Chunk);
  c->id = id;
  c->origin[0] = (float)(8 * id.x) * m.res;
  c->origin[1] = (float)(8 * id.y) * m.res;
  c->origin[2] = (float)(8 * id.z) * m.res;
  for (int i = 0; i < 512; i++) {
    c->sdf[i] = 999.0f;
    c->weight[i] = 0.0f;
  }
  std::memset(c->color, 0, sizeof(c->color));
  Chunk* p = c.get();
  m.chunks.emplace(id, std::move(c));
  return p;
}

// ProjectionIntegrator::voxelUpdateSIMD — 3rd_party/open_chisel/utils/ProjectionIntegrator.cpp:67-426.
bool voxel_update(const Map& m, const float* depth, const uint8_t* rgba, const Cam& cam,
                  const float* pose, int integrate, Chunk* ch, const float* quality, float& q_out) {
  float qsum = 0;
  bool updated = false;
  const float res = m.res;
  const float diag = sqrt(3.0f) * res;  // ::sqrt(double) * float -> double -> float (:77)
  const Intr in = truncated(cam);
  float d0[3] = {ch->origin[0] - pose[12], ch->origin[1] - pose[13], ch->origin[2] - pose[14]}, o[3];
  rt_mul(pose, d0, o);
  const float trunc = truncation_distance(m.trunc, o[2]);
  float w_d = weight_of(m.trunc, trunc);
  if (!integrate) w_d *= -1.0f;
  const float thr_color = diag / 2 + 0.01;  // double add, then float (:101)

  const __m256 o0 = _mm256_set1_ps(o[0]), o1 = _mm256_set1_ps(o[1]), o2 = _mm256_set1_ps(o[2]);
  const __m256 fx = _mm256_set1_ps(in.fx), fy = _mm256_set1_ps(in.fy);
  const __m256 cxh = _mm256_set1_ps((in.cx + 0.5)), cyh = _mm256_set1_ps((in.cy + 0.5));
  const __m256i zero = _mm256_set1_epi32(0);
  const __m256i wm1 = _mm256_set1_epi32(cam.width - 1), hm1 = _mm256_set1_epi32(cam.height - 1);
  const __m256i wv = _mm256_set1_epi32(cam.width);
  const __m256 fzero = _mm256_set1_ps(0.0);
  const __m256 sigma = _mm256_set1_ps(1e-4);
  const __m256 dmin = _mm256_set1_ps(cam.near_plane), dmax = _mm256_set1_ps(cam.far_plane);
  const __m256 wd8 = _mm256_set1_ps(w_d);

  int pos = 0;
  for (int it = 0; it < 64; it++) {  // 8x8 (z,y) iterations, one 8-voxel x-row each
    const __m256 c0 = _mm256_add_ps(o0, _mm256_load_ps(&m.cen[0][pos * 8]));
    const __m256 c1 = _mm256_add_ps(o1, _mm256_load_ps(&m.cen[1][pos * 8]));
    const __m256 c2 = _mm256_add_ps(o2, _mm256_load_ps(&m.cen[2][pos * 8]));
    const __m256 pu = _mm256_add_ps(_mm256_mul_ps(_mm256_div_ps(c0, c2), fx), cxh);
    const __m256 pv = _mm256_add_ps(_mm256_mul_ps(_mm256_div_ps(c1, c2), fy), cyh);
    const __m256i u = _mm256_cvtps_epi32(pu), v = _mm256_cvtps_epi32(pv);
    __m256i valid = _mm256_and_si256(_mm256_cmpgt_epi32(u, zero), _mm256_cmpgt_epi32(wm1, u));
    valid = _mm256_and_si256(valid, _mm256_cmpgt_epi32(v, zero));
    valid = _mm256_and_si256(valid, _mm256_cmpgt_epi32(hm1, v));
    // reference: `continue` before `pos++` (:176-178 vs :420) — the row index never advances
    // again, so every remaining iteration re-tests this row and skips.
    if (_mm256_testz_si256(valid, valid)) continue;
    const __m256 validf = _mm256_castsi256_ps(valid);
    const __m256i pix = _mm256_add_epi32(_mm256_mullo_epi32(v, wv), u);
    const __m256 d = _mm256_mask_i32gather_ps(fzero, depth, pix, validf, 4);
    const __m256 sd = _mm256_sub_ps(d, c2);

    if (rgba != nullptr) {
      // valid & cvtps_epi32(all-ones mask) == valid & 0x80000000 per lane (:202-208)
      const __m256i upd = _mm256_and_si256(
          valid, _mm256_cvtps_epi32(_mm256_and_ps(
                     _mm256_cmp_ps(sd, _mm256_set1_ps(-thr_color), _CMP_GT_OS),
                     _mm256_cmp_ps(_mm256_set1_ps(thr_color), sd, _CMP_GT_OS))));
      __m256i oob = _mm256_or_si256(_mm256_cmpgt_epi32(zero, u), _mm256_cmpgt_epi32(u, wm1));
      oob = _mm256_or_si256(oob, _mm256_cmpgt_epi32(zero, v));
      oob = _mm256_or_si256(oob, _mm256_cmpgt_epi32(v, hm1));
      if (!_mm256_testz_si256(oob, oob)) qsum = -99999999999;
      if (!_mm256_testz_si256(upd, upd)) {
        if (quality != nullptr) {
          const __m256 q = _mm256_mask_i32gather_ps(fzero, quality, pix, _mm256_cvtepi32_ps(upd), 4);
          alignas(32) float ql[8];
          _mm256_store_ps(ql, q);
          float s = 0;
          for (int l = 0; l < 8; l++) s += ql[l];
          qsum += s;
        }
        const __m256i px =
            _mm256_mask_i32gather_epi32(_mm256_set1_epi16(0), (const int*)rgba, pix, upd, 4);
        for (int h = 0; h < 2; h++) {
          __m256i* dst = (__m256i*)&ch->color[pos * 32 + 16 * h];
          const __m256i cur = _mm256_load_si256(dst);
          const __m256i add = _mm256_cvtepu8_epi16(h ? _mm256_extracti128_si256(px, 1)
                                                     : _mm256_extracti128_si256(px, 0));
          __m256i nv;
          if (integrate) {
            nv = _mm256_add_epi16(cur, add);
            __m256i sat = _mm256_cmpgt_epi16(nv, _mm256_set1_epi16(120));
            sat = _mm256_shufflehi_epi16(sat, 255);  // broadcast the count channel's flag
            sat = _mm256_shufflelo_epi16(sat, 255);
            nv = _mm256_blendv_epi8(nv, _mm256_srli_epi16(nv, 2), sat);
          } else {
            nv = _mm256_sub_epi16(cur, add);
          }
          _mm256_store_si256(dst, nv);
        }
      }
    }

    const __m256 in_range =
        _mm256_and_ps(_mm256_cmp_ps(d, dmin, _CMP_GT_OS), _mm256_cmp_ps(dmax, d, _CMP_GT_OS));
    const __m256 in_band = _mm256_and_ps(_mm256_cmp_ps(sd, _mm256_set1_ps(-0.03), _CMP_GT_OS),
                                         _mm256_cmp_ps(_mm256_set1_ps(trunc + diag), sd, _CMP_GT_OS));
    const __m256 flag = _mm256_and_ps(in_range, in_band);
    if (!_mm256_testz_ps(flag, flag)) {
      updated = true;
      const __m256 w = _mm256_loadu_ps(&ch->weight[pos * 8]);
      const __m256 s = _mm256_loadu_ps(&ch->sdf[pos * 8]);
      const __m256 nw = _mm256_blendv_ps(fzero, wd8, flag);
      __m256 ns = _mm256_div_ps(_mm256_add_ps(_mm256_mul_ps(s, w), _mm256_mul_ps(sd, nw)),
                                _mm256_add_ps(_mm256_add_ps(w, nw), sigma));
      __m256 nwt = _mm256_add_ps(w, nw);
      const __m256 keep = _mm256_cmp_ps(nwt, _mm256_set1_ps(0.5), _CMP_GT_OS);
      ns = _mm256_blendv_ps(_mm256_set1_ps(999), ns, keep);
      nwt = _mm256_blendv_ps(_mm256_set1_ps(0.0f), nwt, keep);
      _mm256_storeu_ps(&ch->weight[pos * 8], nwt);
      _mm256_storeu_ps(&ch->sdf[pos * 8], ns);
    }
    pos++;
  }
  q_out = qsum;
  return updated;
}

// chisel::parallel_for — 3rd_party/open_chisel/threading/Threading.h:35-53.
template <class F>
void parallel_for(int64_t n, int threads_cfg, F&& f) {
  int nthreads = threads_cfg > 0 ? threads_cfg : (int)std::thread::hardware_concurrency() - 2;
  if (nthreads < 1) nthreads = 1;
  if (threads_cfg == 1) {
    for (int64_t i = 0; i < n; i++) f(i);
    return;
  }
  const int64_t group = std::max<int64_t>(std::max<int64_t>(1, 1000), n / nthreads);
  std::vector<std::thread> th;
  int64_t it = 0;
  for (; it < n - group; it = std::min(it + group, n)) {
    const int64_t a = it, b = std::min(it + group, n);
    th.emplace_back([a, b, &f]() {
      for (int64_t i = a; i < b; i++) f(i);
    });
  }
  for (int64_t i = it; i < n; i++) f(i);
  for (auto& t : th) t.join();
}

}  // namespace

extern "C" {

struct tfo_map;  // opaque alias of Map

const char* tfo_impl() { return "restatement (tf_oracle.cpp)"; }
// 0: x0 + (x1 + x2) (Eigen >= 3.3), 1: (x0 + x1) + x2 (Eigen 3.2).  Process-wide.
void tfo_set_dot3_order(int left_to_right) { g_dot3_l2r = left_to_right ? 1 : 0; }

tfo_map* tfo_create(float res, const float* trunc5, int threads) {
  Map* m = new Map;
  m->res = res;
  m->trunc = Trunc{trunc5[0], trunc5[1], trunc5[2], trunc5[3], trunc5[4]};
  m->threads = threads;
  // Atlas::SetResolution — Structure/Atlas.h:62-65
  m->patch_w = (int)std::floor(4800 * res);
  m->patch_h = (int)std::floor(3600 * res);
  return (tfo_map*)m;
}

void tfo_destroy(tfo_map* h) { delete (Map*)h; }

int tfo_threads_used(tfo_map* h) {
  Map* m = (Map*)h;
  if (m->threads == 1) return 1;
  int n = m->threads > 0 ? m->threads : (int)std::thread::hardware_concurrency() - 2;
  return n < 1 ? 1 : n;
}

void tfo_reset(tfo_map* h) {
  Map* m = (Map*)h;
  m->chunks.clear();
  m->meshes_to_update.clear();
}

void tfo_boundary_ids(tfo_map* h, const float* depth, const float* pose, const Cam* cam, int32_t* min3,
                      int32_t* max3) {
  Id3 mn, mx;
  boundary_ids(*(Map*)h, depth, *cam, pose, mx, mn);
  min3[0] = mn.x, min3[1] = mn.y, min3[2] = mn.z;
  max3[0] = mx.x, max3[1] = mx.y, max3[2] = mx.z;
}

// Culling only (no allocation): GetChunkIDsObservedByCamera.  Returns the hit count.
int64_t tfo_observed_ids(tfo_map* h, const float* depth, const float* pose, const Cam* cam,
                         int32_t* ids_out, int64_t cap) {
  std::vector<Id3> ids;
  observed_ids(*(Map*)h, depth, *cam, pose, ids);
  const int64_t n = (int64_t)ids.size();
  for (int64_t i = 0; i < std::min(n, cap); i++) {
    ids_out[3 * i] = ids[i].x, ids_out[3 * i + 1] = ids[i].y, ids_out[3 * i + 2] = ids[i].z;
  }
  return n;
}

// Chisel::PrepareIntersectChunks — Structure/Chisel.h:103-140.
int64_t tfo_prepare(tfo_map* h, const float* depth, const float* pose, const Cam* cam, int32_t* ids_out,
                    uint8_t* is_new_out, int64_t cap) {
  Map& m = *(Map*)h;
  buffer_centroids(m, pose);
  std::vector<Id3> ids;
  observed_ids(m, depth, *cam, pose, ids);
  const int64_t n = (int64_t)ids.size();
  if (n > cap) return -n;
  for (int64_t i = 0; i < n; i++) {
    bool is_new = false;
    if (m.chunks.find(ids[i]) == m.chunks.end()) {
      is_new = true;
      create_chunk(m, ids[i]);
    }
    ids_out[3 * i] = ids[i].x, ids_out[3 * i + 1] = ids[i].y, ids_out[3 * i + 2] = ids[i].z;
    is_new_out[i] = is_new;
  }
  return n;
}

// Chisel::IntegrateDepthScanColor (list form) — Structure/Chisel.h:218-249.
// quality_out (optional) receives the raw chunkObservationQuality per chunk.
int tfo_integrate(tfo_map* h, const float* depth, const uint8_t* rgba, const float* quality,
                  const float* pose, const Cam* cam, const int32_t* ids, int64_t n, int flag,
                  int keyframe_id, uint8_t* needs_update, float* quality_out) {
  Map& m = *(Map*)h;
  buffer_centroids(m, pose);
  if (n < 1) return 0;
  for (int64_t i = 0; i < n; i++)
    if (m.chunks.find(Id3{ids[3 * i], ids[3 * i + 1], ids[3 * i + 2]}) == m.chunks.end()) return -4;
  parallel_for(n, m.threads, [&](int64_t i) {
    Chunk* ch = m.chunks.find(Id3{ids[3 * i], ids[3 * i + 1], ids[3 * i + 2]})->second.get();
    float q = 0;
    bool upd = voxel_update(m, depth, rgba, *cam, pose, flag, ch, quality, q);
    needs_update[i] = (needs_update[i] || upd);
    if (quality_out) quality_out[i] = q;
    if (keyframe_id >= 0 && q > 0 && needs_update[i]) ch->observations[keyframe_id] = q;
  });
  return 0;
}

// Chisel::FinalizeIntegrateChunks + GarbageCollect — Structure/Chisel.h:184-216,472-477.
int64_t tfo_finalize(tfo_map* h, const int32_t* ids, int64_t n, const uint8_t* needs_update,
                     const uint8_t* is_new, int32_t* valid_out) {
  Map& m = *(Map*)h;
  int64_t nv = 0;
  std::vector<Id3> garbage;
  static const int nb[7][3] = {{0, 0, 0}, {-1, 0, 0}, {1, 0, 0}, {0, -1, 0}, {0, 1, 0}, {0, 0, -1}, {0, 0, 1}};
  for (int64_t i = 0; i < n; i++) {
    const Id3 id{ids[3 * i], ids[3 * i + 1], ids[3 * i + 2]};
    if (needs_update[i]) {
      for (auto& d : nb) m.meshes_to_update[Id3{id.x + d[0], id.y + d[1], id.z + d[2]}] = true;
      if (valid_out) valid_out[3 * nv] = id.x, valid_out[3 * nv + 1] = id.y, valid_out[3 * nv + 2] = id.z;
      nv++;
    } else if (is_new[i]) {
      garbage.push_back(id);
    }
  }
  for (auto& id : garbage) {
    m.chunks.erase(id);
    m.meshes_to_update.erase(id);
  }
  return nv;
}

// Chisel::IntegrateDepthScanColor (convenience form) — Structure/Chisel.h:453-468:
// Prepare + Integrate(flag 1, keyframeID -1, no quality plane) + Finalize.
// keyframe_id/quality are an extension used by the synthetic key-frame protocol (same calls
// ReIntegrateKeyframe makes for a key-frame without local frames).  Returns |chunksIntersecting|.
int64_t tfo_integrate_frame(tfo_map* h, const float* depth, const uint8_t* rgba, const float* quality,
                            const float* pose, const Cam* cam, int keyframe_id, int64_t* n_updated) {
  Map& m = *(Map*)h;
  buffer_centroids(m, pose);
  std::vector<Id3> ids;
  observed_ids(m, depth, *cam, pose, ids);
  const int64_t n = (int64_t)ids.size();
  std::vector<uint8_t> is_new(n), upd(n, 0);
  std::vector<int32_t> flat(3 * n);
  for (int64_t i = 0; i < n; i++) {
    is_new[i] = 0;
    if (m.chunks.find(ids[i]) == m.chunks.end()) {
      is_new[i] = 1;
      create_chunk(m, ids[i]);
    }
    flat[3 * i] = ids[i].x, flat[3 * i + 1] = ids[i].y, flat[3 * i + 2] = ids[i].z;
  }
  tfo_integrate(h, depth, rgba, quality, pose, cam, flat.data(), n, 1, keyframe_id, upd.data(), nullptr);
  int64_t nv = tfo_finalize(h, flat.data(), n, upd.data(), is_new.data(), nullptr);
  if (n_updated) *n_updated = nv;
  return n;
}

int tfo_has_chunk(tfo_map* h, int32_t x, int32_t y, int32_t z) {
  Map& m = *(Map*)h;
  return m.chunks.find(Id3{x, y, z}) != m.chunks.end();
}

int64_t tfo_chunk_count(tfo_map* h) { return (int64_t)((Map*)h)->chunks.size(); }

int64_t tfo_list_chunks(tfo_map* h, int32_t* out, int64_t cap) {
  Map& m = *(Map*)h;
  int64_t i = 0;
  for (auto& kv : m.chunks) {
    if (i < cap) out[3 * i] = kv.first.x, out[3 * i + 1] = kv.first.y, out[3 * i + 2] = kv.first.z;
    i++;
  }
  return i;
}

int tfo_download_chunks(tfo_map* h, const int32_t* ids, int64_t n, float* sdf, float* weight,
                        uint16_t* color) {
  Map& m = *(Map*)h;
  for (int64_t i = 0; i < n; i++) {
    auto it = m.chunks.find(Id3{ids[3 * i], ids[3 * i + 1], ids[3 * i + 2]});
    if (it == m.chunks.end()) return -4;
    if (sdf) std::memcpy(sdf + 512 * i, it->second->sdf, 2048);
    if (weight) std::memcpy(weight + 512 * i, it->second->weight, 2048);
    if (color) std::memcpy(color + 2048 * i, it->second->color, 4096);
  }
  return 0;
}

// chunk->observations[keyframe] lookup (consumed by TexMap::update_datacost, Structure/TexMap.cpp:63-105)
int tfo_get_observation(tfo_map* h, int32_t x, int32_t y, int32_t z, int keyframe, float* out) {
  Map& m = *(Map*)h;
  auto it = m.chunks.find(Id3{x, y, z});
  if (it == m.chunks.end()) return -4;
  auto o = it->second->observations.find(keyframe);
  if (o == it->second->observations.end()) return 0;
  *out = o->second;
  return 1;
}

// Chisel::CompressMeshes ends with chunksToUpdate.clear() on meshesToUpdate (Structure/Chisel.cpp:146)
void tfo_clear_meshes_to_update(tfo_map* h) { ((Map*)h)->meshes_to_update.clear(); }
// MobileFusion::RetractObservations (GCFusion/MobileFusion.cpp:252-272): chunk->observations.erase(frame_id)
// for the listed chunks that are in the map
void tfo_retract_observations(tfo_map* h, const int32_t* ids, int64_t n, int keyframe) {
  Map& m = *(Map*)h;
  for (int64_t i = 0; i < n; i++) {
    auto it = m.chunks.find(Id3{ids[3 * i], ids[3 * i + 1], ids[3 * i + 2]});
    if (it != m.chunks.end()) it->second->observations.erase(keyframe);
  }
}

int64_t tfo_meshes_to_update(tfo_map* h, int32_t* out, int64_t cap) {
  Map& m = *(Map*)h;
  int64_t i = 0;
  for (auto& kv : m.meshes_to_update) {
    if (!kv.second) continue;
    if (i < cap) out[3 * i] = kv.first.x, out[3 * i + 1] = kv.first.y, out[3 * i + 2] = kv.first.z;
    i++;
  }
  return i;
}

// ---- atlas: Structure/Atlas.cpp:32-91 -------------------------------------------------

void tfo_atlas_patch_size(tfo_map* h, int32_t* w, int32_t* ph) {
  *w = ((Map*)h)->patch_w;
  *ph = ((Map*)h)->patch_h;
}

// Atlas::AddPatch placement — Structure/Atlas.cpp:43-64.
int tfo_atlas_alloc_slot(tfo_map* h, int32_t x, int32_t y, int32_t z, uint64_t* texloc) {
  Map& m = *(Map*)h;
  auto it = m.patches.find(Id3{x, y, z});
  if (it != m.patches.end()) {
    *texloc = it->second.texloc;
    return 0;
  }
  const uint64_t loc = m.loc_next;
  uint64_t px = loc % kAtlasDim, py = loc / kAtlasDim;
  if (px >= (uint64_t)kAtlasDim || py >= (uint64_t)kAtlasDim) return -5;
  if (px + m.patch_w >= (uint64_t)kAtlasDim) {
    px = 0;
    py += m.patch_h;
  } else {
    px += m.patch_w;
  }
  m.loc_next = px + py * kAtlasDim;
  m.patches.emplace(Id3{x, y, z}, Patch{loc});
  *texloc = loc;
  return 0;
}

// Atlas::UpdateBuffer — Structure/Atlas.cpp:71-91 — the copyTo branch and the geometry of
// the resize branch.  The resize itself (cv::resize, INTER_LINEAR, 8UC3) is OpenCV code
// that is not in the reference tree (README pins commit 8f1356c); its published algorithm
// (fixed-point 11-bit coefficients, resize.cpp HResizeLinear/VResizeLinear) is restated in
// resize_linear_8uc3 below and pinned against Python cv2.resize in tests/.
static void resize_linear_8uc3(const uint8_t* src, int sstride, int sw, int sh, uint8_t* dst,
                               int dstride, int dw, int dh) {
  // cv::resize: inv_scale = dsize/ssize, scale = 1/inv_scale (imgproc resize.cpp)
  const double sx = 1.0 / ((double)dw / (double)sw), sy = 1.0 / ((double)dh / (double)sh);
  std::vector<int> xo(dw), yo(dh);
  std::vector<short> xa(2 * dw), ya(2 * dh);
  auto coef = [](int dn, int sn, double scale, std::vector<int>& ofs, std::vector<short>& ab) {
    for (int d = 0; d < dn; d++) {
      float f = (float)((d + 0.5) * scale - 0.5);
      int s = (int)std::floor(f);
      f -= s;
      if (s < 0) { f = 0; s = 0; }
      if (s >= sn - 1) { f = 0; s = sn - 1; }
      ofs[d] = s;
      // saturate_cast<short>(v*2048) with round-half-even (cvRound)
      ab[2 * d] = (short)lrintf((1.f - f) * 2048);
      ab[2 * d + 1] = (short)lrintf(f * 2048);
    }
  };
  coef(dw, sw, sx, xo, xa);
  coef(dh, sh, sy, yo, ya);
  std::vector<int> r0(dw * 3), r1(dw * 3);
  for (int y = 0; y < dh; y++) {
    const int s0 = yo[y], s1 = std::min(s0 + 1, sh - 1);
    for (int x = 0; x < dw; x++) {
      const int a = xo[x], b = std::min(a + 1, sw - 1);
      for (int c = 0; c < 3; c++) {
        r0[x * 3 + c] = src[s0 * sstride + a * 3 + c] * xa[2 * x] + src[s0 * sstride + b * 3 + c] * xa[2 * x + 1];
        r1[x * 3 + c] = src[s1 * sstride + a * 3 + c] * xa[2 * x] + src[s1 * sstride + b * 3 + c] * xa[2 * x + 1];
      }
    }
    const int b0 = ya[2 * y], b1 = ya[2 * y + 1];
    for (int i = 0; i < dw * 3; i++) {
      int v = (((b0 * (r0[i] >> 4)) >> 16) + ((b1 * (r1[i] >> 4)) >> 16) + 2) >> 2;
      dst[y * dstride + i] = (uint8_t)std::min(255, std::max(0, v));
    }
  }
}

int tfo_atlas_update(tfo_map* h, uint64_t texloc, const uint8_t* rgb, int img_w, int img_h, int bx,
                     int by, int bw, int bh) {
  Map& m = *(Map*)h;
  if (bw <= 0 || bh <= 0 || bx < 0 || by < 0 || bx + bw > img_w || by + bh > img_h) return -1;
  if (m.atlas.empty()) m.atlas.assign((size_t)kAtlasDim * kAtlasDim * 3, 0);
  const int ox = (int)(texloc % kAtlasDim), oy = (int)(texloc / kAtlasDim);
  const uint8_t* crop = rgb + ((size_t)by * img_w + bx) * 3;
  const int sstride = img_w * 3;
  float rx = 1, ry = 1;
  if (bw > m.patch_w) rx = float(m.patch_w) / bw;
  if (bh > m.patch_h) ry = float(m.patch_h) / bh;
  if (rx < 1 || ry < 1) {
    if (ox + m.patch_w > kAtlasDim || oy + m.patch_h > kAtlasDim) return -1;
    resize_linear_8uc3(crop, sstride, bw, bh, &m.atlas[((size_t)oy * kAtlasDim + ox) * 3], kAtlasDim * 3,
                       m.patch_w, m.patch_h);
  } else {
    if (ox + bw > kAtlasDim || oy + bh > kAtlasDim) return -1;
    for (int r = 0; r < bh; r++)
      std::memcpy(&m.atlas[((size_t)(oy + r) * kAtlasDim + ox) * 3], crop + (size_t)r * sstride, (size_t)bw * 3);
  }
  return 0;
}

int tfo_atlas_download(tfo_map* h, uint64_t hot_start, uint64_t hot_end, uint8_t* out) {
  Map& m = *(Map*)h;
  if (hot_end < hot_start || hot_end > (uint64_t)kAtlasDim * kAtlasDim) return -1;
  if (m.atlas.empty()) {
    std::memset(out, 0, (hot_end - hot_start) * 3);
    return 0;
  }
  std::memcpy(out, &m.atlas[hot_start * 3], (hot_end - hot_start) * 3);
  return 0;
}

// ---- Patch::CalculateTexCoords — Structure/Patch.cpp:40-108 (+ bilinear :110-146, bilinear_depth :148-170)
struct PatchResult { int32_t x, y, w, h, wrong_mapping, flag; };

// cv::Mat::at beyond the last column reads the next row (contiguous image); beyond the buffer -> 0
static inline float px(const uint8_t* rgb, int W, int H, int y, int x, int c) {
  const long idx = (long)y * W + x;
  return (idx >= 0 && idx < (long)W * H) ? (float)rgb[idx * 3 + c] : 0.0f;
}
static inline float dp(const float* d, int W, int H, int y, int x) {
  const long idx = (long)y * W + x;
  return (idx >= 0 && idx < (long)W * H) ? d[idx] : 0.0f;
}
static void bilinear_rgb(const uint8_t* rgb, int W, int H, float lx, float ly, float out[3]) {
  const int x = (int)std::floor(lx), y = (int)std::floor(ly);
  for (int c = 0; c < 3; c++) {
    if (x < W - 1 && y < H - 1) {
      const float c1 = px(rgb, W, H, y, x, c), c2 = px(rgb, W, H, y, x + 1, c), c3 = px(rgb, W, H, y + 1, x, c);
      // (the reference uses c2 where c4 is meant, Patch.cpp:125-128)
      out[c] = c1 * (x + 1 - lx) * (y + 1 - ly) + c2 * (lx - x) * (y + 1 - ly) + c3 * (x + 1 - lx) * (ly - y) +
               c2 * (lx - x) * (ly - y);
    } else if (x < W - 1 && y == H - 1) {
      out[c] = px(rgb, W, H, y, x, c) * (x + 1 - lx) + px(rgb, W, H, y, x + 1, c) * (lx - x);
    } else if (x == W - 1 && y < H - 1) {
      out[c] = px(rgb, W, H, y, x, c) * (y + 1 - ly) + px(rgb, W, H, y + 1, x, c) * (ly - y);
    } else {
      out[c] = px(rgb, W, H, y, x, c);
    }
  }
}
static float bilinear_d(const float* d, int W, int H, float lx, float ly) {
  const int x = (int)std::floor(lx), y = (int)std::floor(ly);
  if (x < W - 1 && y < H - 1) {
    const float c1 = dp(d, W, H, y, x), c2 = dp(d, W, H, y, x + 1), c3 = dp(d, W, H, y + 1, x);
    return c1 * (x + 1 - lx) * (y + 1 - ly) + c2 * (lx - x) * (y + 1 - ly) + c3 * (x + 1 - lx) * (ly - y) +
           c2 * (lx - x) * (ly - y);
  } else if (x < W - 1 && y == H - 1) {
    return dp(d, W, H, y, x) * (x + 1 - lx) + dp(d, W, H, y, x + 1) * (lx - x);
  } else if (x == W - 1 && y < H - 1) {
    return dp(d, W, H, y, x) * (y + 1 - ly) + dp(d, W, H, y + 1, x) * (ly - y);
  }
  return dp(d, W, H, y, x);
}

// T: world->camera 4x4 (column-major), i.e. pose_sophus[0].inverse().matrix().cast<float>()
int tfo_patch_texcoords(const uint8_t* rgb, const float* depth, const float* T, const Cam* cam, int64_t n_patches,
                        const int64_t* offsets, const float* verts, const float* colors, float* texcoord,
                        float* texcolor, PatchResult* results) {
  const int W = cam->width, H = cam->height;
  const Intr in = truncated(*cam);
  for (int64_t p = 0; p < n_patches; p++) {
    const int64_t a = offsets[p], b = offsets[p + 1], n = b - a;
    float minX = (float)W, maxX = 0.0f, minY = (float)H, maxY = 0.0f;
    int flag = 0, depth_compare = 0, color_compare = 0;
    for (int64_t i = a; i < b; i++) {
      const float vx = verts[3 * i], vy = verts[3 * i + 1], vz = verts[3 * i + 2];
      float vl[3];
      // Matrix4f * Vector4f (packet evaluation): ((m0*x + m1*y) + m2*z) + m3*1
      for (int k = 0; k < 3; k++) vl[k] = ((T[k] * vx + T[4 + k] * vy) + T[8 + k] * vz) + T[12 + k] * 1.0f;
      const float dist = vl[2];
      const float x = vl[0] / vl[2], y = vl[1] / vl[2];
      float cameraX = x * in.fx + in.cx + 0.5;  // float*int + int, then + 0.5 in double
      float cameraY = y * in.fy + in.cy + 0.5;
      if (cameraX < 0 || cameraX >= W || cameraY < 0 || cameraY >= H) flag = -1;
      if (cameraX < 0) cameraX = 0;
      if (cameraX >= W) cameraX = W;
      if (cameraY < 0) cameraY = 0;
      if (cameraY >= H) cameraY = H;
      texcoord[2 * i] = cameraX;
      texcoord[2 * i + 1] = cameraY;
      minX = minX < cameraX ? minX : cameraX;
      maxX = maxX > cameraX ? maxX : cameraX;
      minY = minY < cameraY ? minY : cameraY;
      maxY = maxY > cameraY ? maxY : cameraY;
      float tc[3];
      bilinear_rgb(rgb, W, H, cameraX, cameraY, tc);
      for (int c = 0; c < 3; c++) tc[c] = tc[c] / 255.0f;
      for (int c = 0; c < 3; c++) texcolor[3 * i + c] = tc[c];
      const float dd = bilinear_d(depth, W, H, cameraX, cameraY);
      const float e0 = tc[0] - colors[3 * i], e1 = tc[1] - colors[3 * i + 1], e2 = tc[2] - colors[3 * i + 2];
      const float nrm = std::sqrt(e0 * e0 + (e1 * e1 + e2 * e2));
      if (nrm > 0.6) color_compare++;
      if (std::fabs(dist - dd) > 0.7) depth_compare++;
    }
    PatchResult& r = results[p];
    r.wrong_mapping = (depth_compare > 0.3 * n || color_compare > 0.3 * n) ? 1 : 0;
    r.flag = flag;
    if (maxX >= minX && maxY >= minY) {
      // cv::Rect(float...) truncates towards zero; & = intersection (empty -> all zero)
      int bx = (int)(minX - 2), by = (int)(minY - 2), bw = (int)(maxX - minX + 5), bh = (int)(maxY - minY + 5);
      const int x1 = std::max(bx, 0), y1 = std::max(by, 0);
      const int x2 = std::min(bx + bw, 0 + W - 1), y2 = std::min(by + bh, 0 + H - 1);
      if (x2 - x1 <= 0 || y2 - y1 <= 0) { r.x = r.y = r.w = r.h = 0; }
      else { r.x = x1; r.y = y1; r.w = x2 - x1; r.h = y2 - y1; }
      for (int64_t i = a; i < b; i++) {
        texcoord[2 * i] -= (float)r.x;
        texcoord[2 * i + 1] -= (float)r.y;
      }
    } else {
      r.x = r.y = r.w = r.h = 0;
    }
  }
  return 0;
}

// expose the scalar helpers for unit tests
float tfo_truncation_distance(const float* trunc5, float z) {
  return truncation_distance(Trunc{trunc5[0], trunc5[1], trunc5[2], trunc5[3], trunc5[4]}, z);
}
void tfo_centroids(tfo_map* h, const float* pose, float* out3x512) {
  Map& m = *(Map*)h;
  buffer_centroids(m, pose);
  std::memcpy(out3x512, m.cen, sizeof(m.cen));
}

}  // extern "C"
