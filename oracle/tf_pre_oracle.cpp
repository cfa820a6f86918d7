// TEST INFRASTRUCTURE — not part of the product.  Only tests/, __graft_entry__.smoke() and tools that
// time the CPU baseline may load the library built from this file; nothing under texturefusion_b200/ does.
//
// libtexfusion_pre_oracle.so: a scalar CPU restatement of the reference's frame pre-processing loops
// (SURVEY.md §8 f3), one pixel at a time, each function citing the lines it follows.  It exports the same
// tfp_* entry points as oracle/_ref/libtexfusion_ref_pre.so (the reference's own text compiled between
// stand-ins, oracle/ref_pre_driver.cpp) and is PINNED by it: tests/test_pre_cpu.py requires bit-identical
// outputs of every function on seeded inputs, in this container where the reference tree exists; the
// committed fixtures tests/golden/pre_*.npz are written by the reference build.
//
// What cannot be pinned: cv::cvtColor / cv::Sobel (OpenCV is absent; checked against cv2 4.13 instead)
// and Eigen's association order in the two dot products of checkColorQuality / estimateColorQuality
// (runtime switch tfp_set_dot3_order, as for the fusion path).
//
// _mm256_rsqrt_ps is emulated with the Intel value table (tools/gen_rsqrt_table.py), so that this port
// gives the same answer on any host.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../texturefusion_b200/csrc/tf_rsqrt_table.h"

namespace {

int g_l2r = 0;  // 0: Eigen >= 3.3 (x0 + (x1 + x2));  1: Eigen 3.2 ((x0 + x1) + x2)

inline float sum3(float a, float b, float c) { return g_l2r ? (a + b) + c : a + (b + c); }

inline float rsqrt_x86(float x) {  // RSQRTPS for the operands these loops produce (positive normal floats)
  uint32_t b;
  memcpy(&b, &x, 4);
  uint32_t r;
  const uint32_t e = (b >> 23) & 0xffu, m = b & 0x7fffffu;
  if (b & 0x80000000u) r = (e == 0) ? 0xff800000u : 0xffc00000u;  // -0 / -denormal -> -inf; negative -> NaN
  else if (e == 0) r = 0x7f800000u;                               // +0 and denormals (treated as zero) -> +inf
  else if (e == 255) r = m ? (b | 0x00400000u) : 0u;              // NaN -> quiet NaN; +inf -> +0
  else {
    const uint32_t p = e & 1u;
    const int k = ((int)e - (p ? 127 : 128)) / 2;
    r = 0x3f000000u + ((uint32_t)kRsqrtTabHost[p * 1024 + (m >> 13)] << 11) - ((uint32_t)k << 23);
  }
  float o;
  memcpy(&o, &r, 4);
  return o;
}

// viewAngle = Vector3f((j-cx)/fx, (i-cy)/fy, 1); normalize(); dot with the normal (BasicAPI.cpp:794-800, 834-840)
// Eigen >= 3.3 normalize(): z = squaredNorm(); if (z > 0) v /= sqrt(z).  3.2: v /= norm().
inline float view_dot_normal(int i, int j, float fx, float fy, float cx, float cy, float nx, float ny, float nz) {
  float vx = ((float)j - cx) / fx, vy = ((float)i - cy) / fy, vz = 1.0f;
  const float z = sum3(vx * vx, vy * vy, vz * vz);
  if (g_l2r || z > 0.0f) {
    const float n = std::sqrt(z);
    vx = vx / n, vy = vy / n, vz = vz / n;
  }
  return sum3(vx * nx, vy * ny, vz * nz);
}

struct Xf {  // 3x4 row-major rigid transform
  float r[3][3], t[3];
  explicit Xf(const float* T) {
    for (int a = 0; a < 3; a++) {
      for (int b = 0; b < 3; b++) r[a][b] = T[a * 4 + b];
      t[a] = T[a * 4 + 3];
    }
  }
};

}  // namespace

extern "C" {

const char* tfp_impl() { return "oracle port (scalar restatement)"; }
void tfp_set_dot3_order(int l2r) { g_l2r = l2r ? 1 : 0; }

uint32_t tfp_rsqrt_bits(uint32_t b) {
  float f;
  memcpy(&f, &b, 4);
  const float o = rsqrt_x86(f);
  uint32_t r;
  memcpy(&r, &o, 4);
  return r;
}
int tfp_host_rsqrt_matches() { return 1; }  // (the port never executes the instruction)

// BasicAPI::extractNormalMapSIMD (BasicAPI.cpp:849-905).  Rows 1..H-2; columns 1 .. 8*floor((W-12)/8)+8
// (the vector loop `for (j = 1; j < W-1-9; j += 8)`); everything else stays 0.
void tfp_normal_map(const float* depth, float* normal, int W, int H, float fx, float fy, float cx, float cy) {
  const size_t np = (size_t)W * H;
  memset(normal, 0, np * 12);
  const float thr = 0.3f;
  const unsigned width_dst = W - 1, height_dst = H - 1;
  for (unsigned i = 1; i < height_dst; i++)
    for (unsigned j0 = 1; j0 < width_dst - 9; j0 += 8)
      for (unsigned l = 0; l < 8; l++) {
        const unsigned j = j0 + l;
        const size_t p = (size_t)i * W + j;
        const float dr = depth[p + 1], db = depth[p + W], dl = depth[p - 1], dt = depth[p - W];
        const float xj = ((float)l + (float)j0) - cx;  // inc + vec8(j) - vec8(cx)
        const float yi = (float)i - cy;                // vec8(i - cy)
        const float u1 = ((xj * (dr - dl) + dr) + dl) / fx;
        const float u2 = (yi * (dr - dl)) / fy;
        const float u3 = dr - dl;
        const float v1 = (xj * (db - dt)) / fx;
        const float v2 = ((yi * (db - dt) + db) + dt) / fy;
        const float v3 = db - dt;
        float nx = u2 * v3 - u3 * v2, ny = u3 * v1 - u1 * v3, nz = u1 * v2 - u2 * v1;
        const float nsq = (nx * nx + ny * ny) + nz * nz;
        const bool valid = u3 < thr && u3 > -thr && v3 < thr && v3 > -thr && nsq > 1e-24f;
        const float rs = rsqrt_x86(nsq);
        nx = nx * rs, ny = ny * rs, nz = nz * rs;
        normal[p] = valid ? nx : 0.0f;
        normal[p + np] = valid ? ny : 0.0f;
        normal[p + 2 * np] = valid ? nz : 0.0f;
      }
}

// BasicAPI::refineDepthUseNormalSIMD (BasicAPI.cpp:728-780): pixels seen at a grazing angle
// (|view . normal| < 0.1, which includes every pixel without a normal) lose depth and normal.
void tfp_refine_depth_by_normal(float* normal, float* depth, int W, int H, float fx, float fy, float cx, float cy) {
  const size_t np = (size_t)W * H;
  for (int i = 0; i < H; i++)
    for (int j = 0; j < W; j++) {
      const size_t p = (size_t)i * W + j;
      float vx = ((float)j - cx) / fx, vy = ((float)i - cy) / fy, vz = 1.0f;
      const float rs = rsqrt_x86((vx * vx + vy * vy) + vz * vz);
      vx = vx * rs, vy = vy * rs, vz = vz * rs;
      const float q = (vx * normal[p] + vy * normal[p + np]) + vz * normal[p + 2 * np];
      if (q > -0.1f && q < 0.1f) depth[p] = 0.0f, normal[p] = 0.0f, normal[p + np] = 0.0f, normal[p + 2 * np] = 0.0f;
    }
}

// BasicAPI::refineKeyframesSIMD (BasicAPI.cpp:506-636).  T = ref -> new.  The key-frame's depth is updated
// IN PLACE in raster order, eight pixels at a time, and the fallback sample `newDepthNearest` is read from
// that same buffer (:597-600) at the NEW frame's pixel position: a vector that reads a position stored by an
// earlier vector sees the updated value.
void tfp_refine_keyframe(float* kf_depth, float* kf_weight, const float* new_depth, const float* T, int W, int H, float fx, float fy,
                         float cx, float cy) {
  const Xf x(T);
  const float thr = 0.05f;
  for (int i = 0; i < H; i++)
    for (int j0 = 0; j0 < W; j0 += 8) {
      float out_d[8], out_w[8];
      for (int l = 0; l < 8; l++) {
        const int p = i * W + j0 + l;
        const float dc = kf_depth[p];
        const float lx = (((float)l + (float)j0) - cx) / fx * dc, ly = ((float)i - cy) / fy * dc;
        float vx = ((x.r[0][0] * lx + x.r[0][1] * ly) + x.r[0][2] * dc) + x.t[0];
        float vy = ((x.r[1][0] * lx + x.r[1][1] * ly) + x.r[1][2] * dc) + x.t[1];
        float vz = ((x.r[2][0] * lx + x.r[2][1] * ly) + x.r[2][2] * dc) + x.t[2];
        const float u = vx / vz * fx + cx, v = vy / vz * fy + cy;
        const bool valid = u > 2.0f && u < (float)(W - 2) && v > 2.0f && v < (float)(H - 2);
        const float fu = std::floor(u), fv = std::floor(v);
        float ul = 0, ur = 0, bl = 0, br = 0, nearest = 0;
        if (valid) {
          const int q = (int)std::lrintf(fu + fv * (float)W);
          ul = new_depth[q], ur = new_depth[q + 1], bl = new_depth[q + W], br = new_depth[q + W + 1];
          nearest = kf_depth[(int)std::lrintf(std::floor(u + 0.5f) + std::floor(v + 0.5f) * (float)W)];
        }
        const float dx = u - fu, dy = v - fv;
        const bool smooth = (ul - ur) < 0.1f && (ul - ur) > -0.1f && (ul - bl) < 0.1f && (ul - bl) > -0.1f && (ul - br) < 0.1f &&
                            (ul - br) > -0.1f;
        float bil = (((1.0f - dx) * (1.0f - dy) * ul + (1.0f - dx) * dy * ur) + dx * (1.0f - dy) * bl) + dx * dy * br;
        if (!smooth) bil = nearest;
        const bool ok = (bil - vz) > (-thr * vz) && (bil - vz) < (thr * vz);
        const float s = bil / vz;
        vx = vx * s - x.t[0], vy = vy * s - x.t[1], vz = vz * s - x.t[2];
        const float z = (x.r[0][2] * vx + x.r[1][2] * vy) + x.r[2][2] * vz;  // row 2 of the transposed rotation
        const float w = kf_weight[p];
        out_d[l] = ok ? (dc * w + z) / (w + 1.0f) : dc;
        out_w[l] = ok ? w + 1.0f : w;
      }
      for (int l = 0; l < 8; l++) kf_depth[i * W + j0 + l] = out_d[l], kf_weight[i * W + j0 + l] = out_w[l];
    }
}

// BasicAPI::refineNewframesSIMD (BasicAPI.cpp:378-442).  T = new -> ref.  A pixel of the new frame survives
// when the key-frame's depth at its projection agrees within 5 %.
void tfp_refine_newframe(const float* kf_depth, float* new_depth, const float* T, int W, int H, float fx, float fy, float cx, float cy) {
  const Xf x(T);
  const float thr = 0.05f;
  const float cxh = (float)((double)cx + 0.5), cyh = (float)((double)cy + 0.5);
  for (int i = 0; i < H; i++)
    for (int j = 0; j < W; j++) {
      const int p = i * W + j;
      const float dc = new_depth[p];
      const float lx = ((float)j - cx) / fx * dc, ly = ((float)i - cy) / fy * dc;
      const float vx = ((x.r[0][0] * lx + x.r[0][1] * ly) + x.r[0][2] * dc) + x.t[0];
      const float vy = ((x.r[1][0] * lx + x.r[1][1] * ly) + x.r[1][2] * dc) + x.t[1];
      const float vz = ((x.r[2][0] * lx + x.r[2][1] * ly) + x.r[2][2] * dc) + x.t[2];
      const float u = vx / vz * fx + cxh, v = vy / vz * fy + cyh;
      const bool valid = u > 1.0f && u < (float)(W - 1) && v > 1.0f && v < (float)(H - 1);
      const float nd = valid ? kf_depth[(int)std::lrintf(std::floor(u) + std::floor(v) * (float)W)] : 0.0f;
      const bool ok = (nd - vz) > (-thr * vz) && (nd - vz) < (thr * vz);
      new_depth[p] = ok ? dc : 0.0f;
    }
}

// BasicAPI::checkColorQuality (BasicAPI.cpp:783-805); flags the reference leaves unwritten are 0 here.
void tfp_color_valid(const float* normal, uint8_t* flag, int W, int H, float fx, float fy, float cx, float cy) {
  const size_t np = (size_t)W * H;
  for (int i = 0; i < H; i++)
    for (int j = 0; j < W; j++) {
      const size_t p = (size_t)i * W + j;
      const float q = view_dot_normal(i, j, fx, fy, cx, cy, normal[p], normal[p + np], normal[p + 2 * np]);
      flag[p] = std::fabs(q) >= 0.2 ? 1 : 0;  // (compared as double, :800)
    }
}

// cv::cvtColor(RGB2GRAY), 8-bit: see oracle/ref_pre_driver.cpp
void tfp_gray(const uint8_t* rgb, uint8_t* gray, int W, int H) {
  for (size_t i = 0; i < (size_t)W * H; i++)
    gray[i] = (uint8_t)((rgb[3 * i] * 9798 + rgb[3 * i + 1] * 19235 + rgb[3 * i + 2] * 3735 + (1 << 14)) >> 15);
}
// cv::Sobel(gray, CV_32F, 1, 1), 3x3, BORDER_REFLECT_101
void tfp_sobel11(const uint8_t* g, float* out, int W, int H) {
  auto at = [&](int y, int x) {
    y = y < 0 ? -y : (y >= H ? 2 * H - 2 - y : y);
    x = x < 0 ? -x : (x >= W ? 2 * W - 2 - x : x);
    return (int)g[(size_t)y * W + x];
  };
  for (int y = 0; y < H; y++)
    for (int x = 0; x < W; x++) out[(size_t)y * W + x] = (float)(at(y + 1, x + 1) - at(y + 1, x - 1) - at(y - 1, x + 1) + at(y - 1, x - 1));
}

// BasicAPI::estimateColorQuality (BasicAPI.cpp:814-847): |Sobel_xy(gray)| * |view . normal| where the pixel has
// depth; elsewhere the raw (signed) Sobel response stays.
void tfp_color_quality(const float* depth, const float* normal, const uint8_t* rgb, float* quality, int W, int H, float fx, float fy,
                       float cx, float cy) {
  const size_t np = (size_t)W * H;
  std::vector<uint8_t> gray(np);
  tfp_gray(rgb, gray.data(), W, H);
  tfp_sobel11(gray.data(), quality, W, H);
  for (int i = 0; i < H; i++)
    for (int j = 0; j < W; j++) {
      const size_t p = (size_t)i * W + j;
      const float q = std::fabs(view_dot_normal(i, j, fx, fy, cx, cy, normal[p], normal[p + np], normal[p + 2 * np]));
      if (depth[p] > 0.0f) quality[p] = std::fabs(quality[p]) * q;
    }
}

}  // extern "C"
