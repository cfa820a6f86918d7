// Frame pre-processing on the device: the per-pixel loops that run between loading a
// frame and fusing it (main.cpp:117-147).  One thread per pixel; every plane access is a coalesced
// row segment except the projective gathers.  All arithmetic is the reference's, operation by
// operation (un-fused: the library is built with -fmad=false; `/` and sqrtf are IEEE), so the planes
// are bit-identical to the AVX2 loops on an Intel host:
//   * `_mm256_rsqrt_ps` (an approximate instruction) is looked up in the instruction's value table,
//     tf_rsqrt_table.h — RSQRTPS on Intel is a function of the exponent parity and the top 10
//     mantissa bits of its operand;
//   * pixels the reference never writes (cv::Mat::create does not clear) are 0 here.
// These kernels are HBM-streaming: 4-40 B per pixel, a 640x480 frame is one wave.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "tf_rsqrt_table.h"

namespace tfb {

struct PreCam {
  float fx, fy, cx, cy;  // float intrinsics exactly as main.cpp passes them (camera.c_fx ...): NOT truncated
  int W, H;
  int l2r;  // association of Eigen's 3-term sums (tf_config.dot3_order)
};
struct PreXf {  // rigid transform, rows of [R | t]
  float r[3][3], t[3];
};

// RSQRTPS: the table for positive normal operands; zero / denormal -> inf, negative -> NaN, inf -> 0 as the instruction does
__device__ __forceinline__ float rsqrt_x86(float x) {
  const unsigned b = __float_as_uint(x), e = (b >> 23) & 0xffu, m = b & 0x7fffffu;
  unsigned r;
  if (b & 0x80000000u) r = (e == 0) ? 0xff800000u : 0xffc00000u;
  else if (e == 0) r = 0x7f800000u;
  else if (e == 255) r = m ? (b | 0x00400000u) : 0u;
  else {
    const unsigned p = e & 1u;
    const int k = ((int)e - (p ? 127 : 128)) / 2;
    r = 0x3f000000u + ((unsigned)kRsqrtTabDev[p * 1024 + (m >> 13)] << 11) - ((unsigned)k << 23);
  }
  return __uint_as_float(r);
}

__device__ __forceinline__ float pre_sum3(int l2r, float a, float b, float c) { return l2r ? (a + b) + c : a + (b + c); }

// Eigen: viewAngle = Vector3f((j-cx)/fx, (i-cy)/fy, 1).normalize(); viewAngle . normal  (BasicAPI.cpp:794-800, 834-840)
__device__ __forceinline__ float view_dot_normal(const PreCam& c, int i, int j, float nx, float ny, float nz) {
  float vx = ((float)j - c.cx) / c.fx, vy = ((float)i - c.cy) / c.fy, vz = 1.0f;
  const float z = pre_sum3(c.l2r, vx * vx, vy * vy, vz * vz);
  if (c.l2r || z > 0.0f) {
    const float n = sqrtf(z);
    vx = vx / n, vy = vy / n, vz = vz / n;
  }
  return pre_sum3(c.l2r, vx * nx, vy * ny, vz * nz);
}

// BasicAPI::extractNormalMapSIMD (BasicAPI.cpp:849-905): central differences of the back-projected
// vertex map, cross product, RSQRTPS normalisation; rows 1..H-2 and the columns the reference's
// 8-wide loop covers (j = 1, 9, ... while j < W-10), zero elsewhere.
__global__ void __launch_bounds__(256) pre_normal_kernel(const PreCam c, const float* __restrict__ depth, float* __restrict__ normal) {
  const int np = c.W * c.H;
  const int last = c.W > 11 ? 8 * ((c.W - 12) / 8) + 8 : 0;  // last column the vector loop writes
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < np; p += gridDim.x * blockDim.x) {
    const int i = p / c.W, j = p - i * c.W;
    float nx = 0.0f, ny = 0.0f, nz = 0.0f;
    if (i >= 1 && i < c.H - 1 && j >= 1 && j <= last) {
      const float dr = depth[p + 1], db = depth[p + c.W], dl = depth[p - 1], dt = depth[p - c.W];
      const int j0 = ((j - 1) & ~7) + 1;
      const float xj = ((float)(j - j0) + (float)j0) - c.cx;  // inc + vec8(j) - vec8(cx)
      const float yi = (float)i - c.cy;
      const float u1 = ((xj * (dr - dl) + dr) + dl) / c.fx;
      const float u2 = (yi * (dr - dl)) / c.fy;
      const float u3 = dr - dl;
      const float v1 = (xj * (db - dt)) / c.fx;
      const float v2 = ((yi * (db - dt) + db) + dt) / c.fy;
      const float v3 = db - dt;
      const float ax = u2 * v3 - u3 * v2, ay = u3 * v1 - u1 * v3, az = u1 * v2 - u2 * v1;
      const float nsq = (ax * ax + ay * ay) + az * az;
      const bool valid = u3 < 0.3f && u3 > -0.3f && v3 < 0.3f && v3 > -0.3f && nsq > 1e-24f;
      if (valid) {
        const float rs = rsqrt_x86(nsq);
        nx = ax * rs, ny = ay * rs, nz = az * rs;
      }
    }
    normal[p] = nx, normal[p + np] = ny, normal[p + 2 * np] = nz;
  }
}

// BasicAPI::refineDepthUseNormalSIMD (BasicAPI.cpp:728-780): grazing-angle rejection
__global__ void __launch_bounds__(256) pre_grazing_kernel(const PreCam c, float* __restrict__ normal, float* __restrict__ depth) {
  const int np = c.W * c.H;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < np; p += gridDim.x * blockDim.x) {
    const int i = p / c.W, j = p - i * c.W;
    float vx = ((float)j - c.cx) / c.fx, vy = ((float)i - c.cy) / c.fy, vz = 1.0f;
    const float rs = rsqrt_x86((vx * vx + vy * vy) + vz * vz);
    vx = vx * rs, vy = vy * rs, vz = vz * rs;
    const float q = (vx * normal[p] + vy * normal[p + np]) + vz * normal[p + 2 * np];
    if (q > -0.1f && q < 0.1f) depth[p] = 0.0f, normal[p] = 0.0f, normal[p + np] = 0.0f, normal[p + 2 * np] = 0.0f;
  }
}

// ---- BasicAPI::refineKeyframesSIMD (BasicAPI.cpp:506-636) --------------------------------------------
//
// Every key-frame pixel is projected into the new frame, the new frame's depth is sampled there
// (bilinear, or — across a depth edge — a nearest sample), and where the two agree within 5 % the
// key-frame's depth becomes the running mean of its observations.
//
// The reference updates the key-frame IN PLACE, eight pixels per step in raster order, and its
// nearest sample reads the KEY-FRAME's own depth (:597-600) at the new frame's pixel position: a step
// that reads a position written by an earlier step sees the UPDATED value.  Here every pixel is
// computed from a snapshot of the old plane straight into the key-frame's plane; the few pixels whose
// nearest sample lies in an earlier 8-pixel step are queued and resolved in dependency order by the
// last block to finish (a pixel's source is always earlier in raster order, so the queue drains).
struct RefineKfArgs {
  PreCam c;
  PreXf x;              // key-frame camera -> new camera
  const float* kf_d;    // snapshot of the key-frame's depth before this call
  const float* kf_w;    // weights (each pixel reads and writes only its own: updated in place, kf_w == out_w)
  const float* new_d;   // new frame's depth
  float* out_d;         // the key-frame's depth plane
  float* out_w;
  int* queue;           // pixels waiting for an updated nearest sample
  int* queue_n;
  unsigned* ticket;     // last-block election (resets itself)
  unsigned char* waiting;  // per pixel: 1 while queued
};

// One pixel.  use_new: take the nearest sample from out_d (the updated plane) instead of kf_d.
// Returns the pixel index of the nearest sample when it has to come from the updated plane and
// use_new is false (the caller queues the pixel), else -1.
__device__ __forceinline__ int refine_kf_pixel(const RefineKfArgs& a, int p, bool use_new) {
  const PreCam& c = a.c;
  const int i = p / c.W, j = p - i * c.W;
  const int j0 = j & ~7;
  const float dc = a.kf_d[p];
  const float lx = (((float)(j - j0) + (float)j0) - c.cx) / c.fx * dc, ly = ((float)i - c.cy) / c.fy * dc;
  float vx = ((a.x.r[0][0] * lx + a.x.r[0][1] * ly) + a.x.r[0][2] * dc) + a.x.t[0];
  float vy = ((a.x.r[1][0] * lx + a.x.r[1][1] * ly) + a.x.r[1][2] * dc) + a.x.t[1];
  float vz = ((a.x.r[2][0] * lx + a.x.r[2][1] * ly) + a.x.r[2][2] * dc) + a.x.t[2];
  const float u = vx / vz * c.fx + c.cx, v = vy / vz * c.fy + c.cy;
  const bool valid = u > 2.0f && u < (float)(c.W - 2) && v > 2.0f && v < (float)(c.H - 2);
  const float fu = floorf(u), fv = floorf(v);
  float ul = 0.0f, ur = 0.0f, bl = 0.0f, br = 0.0f;
  if (valid) {
    const int q = __float2int_rn(fu + fv * (float)c.W);
    ul = a.new_d[q], ur = a.new_d[q + 1], bl = a.new_d[q + c.W], br = a.new_d[q + c.W + 1];
  }
  const float dx = u - fu, dy = v - fv;
  const bool smooth = (ul - ur) < 0.1f && (ul - ur) > -0.1f && (ul - bl) < 0.1f && (ul - bl) > -0.1f && (ul - br) < 0.1f &&
                      (ul - br) > -0.1f;
  float bil = (((1.0f - dx) * (1.0f - dy) * ul + (1.0f - dx) * dy * ur) + dx * (1.0f - dy) * bl) + dx * dy * br;
  if (!smooth) {
    float nearest = 0.0f;
    if (valid) {
      const int qn = __float2int_rn(floorf(u + 0.5f) + floorf(v + 0.5f) * (float)c.W);
      const bool earlier = (qn >> 3) < (p >> 3);  // written by an earlier 8-pixel step (W is a multiple of 8)
      if (earlier && !use_new) return qn;
      nearest = earlier ? __ldcg(a.out_d + qn) : a.kf_d[qn];  // (updated plane: written by other blocks, read through L2)
    }
    bil = nearest;
  }
  const bool ok = (bil - vz) > (-0.05f * vz) && (bil - vz) < (0.05f * vz);
  const float s = bil / vz;
  vx = vx * s - a.x.t[0], vy = vy * s - a.x.t[1], vz = vz * s - a.x.t[2];
  const float z = (a.x.r[0][2] * vx + a.x.r[1][2] * vy) + a.x.r[2][2] * vz;
  const float w = a.kf_w[p];
  a.out_d[p] = ok ? (dc * w + z) / (w + 1.0f) : dc;
  a.out_w[p] = ok ? w + 1.0f : w;
  return -1;
}

__global__ void __launch_bounds__(256) pre_refine_kf_kernel(const RefineKfArgs a) {
  const int np = a.c.W * a.c.H;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < np; p += gridDim.x * blockDim.x) {
    const int src = refine_kf_pixel(a, p, false);
    a.waiting[p] = src >= 0;
    if (src >= 0) a.queue[atomicAdd(a.queue_n, 1)] = p;
  }
  if (!last_block_done(a.ticket)) return;
  // ---- last block: the queued pixels, round by round; a pixel whose nearest sample is no longer waiting is computed
  const int n = *(volatile int*)a.queue_n;
  if (n == 0) return;
  for (;;) {
    int left = 0;
    // (1) decide on the flags as they stand at the start of the round
    for (int k = threadIdx.x; k < n; k += blockDim.x) {
      const int p = a.queue[k];
      if (p < 0) continue;
      const int src = refine_kf_pixel(a, p, false);  // (re-derives the sample position: cheaper than storing it)
      if (__ldcg(a.waiting + src)) {
        left++;
        continue;
      }
      a.queue[k] = -(p + 2);  // ready: computed in step (2)
    }
    __syncthreads();
    // (2) compute the ready pixels from the updated plane
    for (int k = threadIdx.x; k < n; k += blockDim.x) {
      const int e = a.queue[k];
      if (e >= -1) continue;
      refine_kf_pixel(a, -e - 2, true);
      a.waiting[-e - 2] = 0;
      a.queue[k] = -1;
    }
    if (__syncthreads_count(left) == 0) break;
  }
}

// BasicAPI::refineNewframesSIMD (BasicAPI.cpp:378-442): a pixel of the new frame survives when the
// key-frame's depth at its projection agrees within 5 %.
__global__ void __launch_bounds__(256) pre_refine_new_kernel(const PreCam c, const PreXf x, const float* __restrict__ kf_d,
                                                             float* __restrict__ new_d) {
  const int np = c.W * c.H;
  const float cxh = (float)((double)c.cx + 0.5), cyh = (float)((double)c.cy + 0.5);
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < np; p += gridDim.x * blockDim.x) {
    const int i = p / c.W, j = p - i * c.W;
    const float dc = new_d[p];
    const float lx = ((float)j - c.cx) / c.fx * dc, ly = ((float)i - c.cy) / c.fy * dc;
    const float vx = ((x.r[0][0] * lx + x.r[0][1] * ly) + x.r[0][2] * dc) + x.t[0];
    const float vy = ((x.r[1][0] * lx + x.r[1][1] * ly) + x.r[1][2] * dc) + x.t[1];
    const float vz = ((x.r[2][0] * lx + x.r[2][1] * ly) + x.r[2][2] * dc) + x.t[2];
    const float u = vx / vz * c.fx + cxh, v = vy / vz * c.fy + cyh;
    const bool valid = u > 1.0f && u < (float)(c.W - 1) && v > 1.0f && v < (float)(c.H - 1);
    const float nd = valid ? kf_d[__float2int_rn(floorf(u) + floorf(v) * (float)c.W)] : 0.0f;
    const bool ok = (nd - vz) > (-0.05f * vz) && (nd - vz) < (0.05f * vz);
    if (!ok) new_d[p] = 0.0f;
  }
}

// ---- key-frame colour planes ----------------------------------------------------------------------
// BasicAPI::checkColorQuality (BasicAPI.cpp:783-805) + estimateColorQuality (:814-847, with
// cv::cvtColor(RGB2GRAY) and cv::Sobel(CV_32F, 1, 1) inlined) + the RGBA pack of
// GCFusion/MobileFusion.cpp:151-162, in one pass over the key-frame: what reaches the fusion kernels
// as `rgba` and `quality` is produced where it is consumed, from the 3-byte RGB upload.
__device__ __forceinline__ int pre_gray(const unsigned char* __restrict__ rgb, int p) {  // OpenCV 4.x: 15-bit fixed point
  return (rgb[3 * p] * 9798 + rgb[3 * p + 1] * 19235 + rgb[3 * p + 2] * 3735 + (1 << 14)) >> 15;
}
__global__ void __launch_bounds__(256) pre_color_kernel(const PreCam c, const float* __restrict__ depth, const float* __restrict__ normal,
                                                        const unsigned char* __restrict__ rgb, unsigned char* __restrict__ valid,
                                                        float* __restrict__ quality, uchar4* __restrict__ rgba) {
  const int np = c.W * c.H;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < np; p += gridDim.x * blockDim.x) {
    const int i = p / c.W, j = p - i * c.W;
    const float q = view_dot_normal(c, i, j, normal[p], normal[p + np], normal[p + 2 * np]);
    const bool ok = (double)fabsf(q) >= 0.2;  // (compared in double, :800)
    valid[p] = ok ? 1 : 0;
    rgba[p] = ok ? make_uchar4(rgb[3 * p], rgb[3 * p + 1], rgb[3 * p + 2], 1) : make_uchar4(0, 0, 0, 0);
    // Sobel dx=1, dy=1: [-1 0 1]^T x [-1 0 1], BORDER_REFLECT_101
    const int yu = i == 0 ? 1 : i - 1, yd = i == c.H - 1 ? c.H - 2 : i + 1;
    const int xl = j == 0 ? 1 : j - 1, xr = j == c.W - 1 ? c.W - 2 : j + 1;
    const float sob = (float)(pre_gray(rgb, yd * c.W + xr) - pre_gray(rgb, yd * c.W + xl) - pre_gray(rgb, yu * c.W + xr) +
                              pre_gray(rgb, yu * c.W + xl));
    quality[p] = depth[p] > 0.0f ? fabsf(sob) * fabsf(q) : sob;
  }
}


// ---- framePreprocess (Tools/DatasetWrapper.hpp:187-263; BasicAPI.cpp:942-1004) ------------------------
// Raw sensor depth -> metres: values beyond the camera's range are dropped (:213-221).
__global__ void __launch_bounds__(256) pre_depth_u16_kernel(const unsigned short* __restrict__ in, float* __restrict__ out, int np,
                                                            float depth_scale, float max_depth) {
  const float limit = max_depth * depth_scale;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < np; p += gridDim.x * blockDim.x) {
    unsigned short v = in[p];
    if ((float)v > limit) v = 0;
    out[p] = (float)v / depth_scale;
  }
}

// cv::bilateralFilter on a CV_32FC1 image (:228-230: d = 9, sigmaColor = 0.03, sigmaSpace = 10), restated from
// OpenCV's float path: range weights from a 4096-bin table of exp() over [0, max - min] with linear
// interpolation, space weights exp(-r^2 / 2 sigma^2) on the disc r <= d/2, BORDER_REFLECT_101, the centre
// tap added with weight 1.  OpenCV accumulates in float in its own (SIMD-width dependent) order, so
// this agrees with cv2 to float rounding (tests: <= 2e-6 relative), not bit for bit.
constexpr int kBilateralBins = 1 << 12;
struct BilateralState {  // device scratch: [0] encoded min, [1] encoded max (ordered-int floats), then the table
  int enc_min, enc_max;
  float scale_index;
  int flat;  // |max - min| < FLT_EPSILON: the filter is a copy
  float lut[kBilateralBins + 2];
};
__global__ void __launch_bounds__(256) pre_minmax_kernel(const float* __restrict__ src, int np, BilateralState* st) {
  float lo = 3.0e38f, hi = -3.0e38f;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < np; p += gridDim.x * blockDim.x) {
    const float v = src[p];
    lo = fminf(lo, v), hi = fmaxf(hi, v);
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) lo = fminf(lo, __shfl_xor_sync(kFull, lo, d)), hi = fmaxf(hi, __shfl_xor_sync(kFull, hi, d));
  if ((threadIdx.x & 31) == 0) {
    atomicMin(&st->enc_min, enc_f(lo));
    atomicMax(&st->enc_max, enc_f(hi));
  }
}
__global__ void __launch_bounds__(1024) pre_bilateral_lut_kernel(BilateralState* st, double gauss_color_coeff) {
  const float mn = dec_f(st->enc_min), mx = dec_f(st->enc_max);
  const float len = (float)((double)mx - (double)mn);
  const float scale_index = (float)kBilateralBins / len;
  if (threadIdx.x == 0) {
    st->scale_index = scale_index;
    st->flat = fabs((double)mn - (double)mx) < 1.1920928955078125e-7 ? 1 : 0;
  }
  for (int i = threadIdx.x; i < kBilateralBins + 2; i += blockDim.x) {
    const double val = (double)((float)i / scale_index);
    st->lut[i] = (float)exp(val * val * gauss_color_coeff);
  }
}
constexpr int kBilMaxRadius = 8;
__global__ void __launch_bounds__(256) pre_bilateral_kernel(const float* __restrict__ src, float* __restrict__ dst, int W, int H, int radius,
                                                            double gauss_space_coeff, const BilateralState* __restrict__ st) {
  // block = 32 x 8 pixels; the tile with its halo is staged in shared memory
  __shared__ float tile[8 + 2 * kBilMaxRadius][32 + 2 * kBilMaxRadius + 1];
  __shared__ float sw[(2 * kBilMaxRadius + 1) * (2 * kBilMaxRadius + 1)];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int x0 = blockIdx.x * 32, y0 = blockIdx.y * 8;
  const int tw = 32 + 2 * radius, th = 8 + 2 * radius, side = 2 * radius + 1;
  for (int k = threadIdx.x; k < tw * th; k += 256) {
    const int lx = k % tw, ly = k / tw;
    int gx = x0 + lx - radius, gy = y0 + ly - radius;
    gx = gx < 0 ? -gx : (gx >= W ? 2 * W - 2 - gx : gx);  // BORDER_REFLECT_101
    gy = gy < 0 ? -gy : (gy >= H ? 2 * H - 2 - gy : gy);
    gx = min(max(gx, 0), W - 1), gy = min(max(gy, 0), H - 1);  // (tiles beyond the image: clamped, never used)
    tile[ly][lx] = src[gy * W + gx];
  }
  for (int k = threadIdx.x; k < side * side; k += 256) {
    const int i = k / side - radius, j = k % side - radius;
    const double r = sqrt((double)i * i + (double)j * j);
    sw[k] = (r > radius || (i == 0 && j == 0)) ? -1.0f : (float)exp(r * r * gauss_space_coeff);
  }
  __syncthreads();
  const int x = x0 + tx, y = y0 + ty;
  if (x >= W || y >= H) return;
  const float val0 = tile[ty + radius][tx + radius];
  if (st->flat) { dst[y * W + x] = val0; return; }
  const float scale_index = st->scale_index;
  float sum = 0.0f, wsum = 0.0f;
  for (int i = 0; i < side; i++)
    for (int j = 0; j < side; j++) {
      const float kw = sw[i * side + j];
      if (kw < 0.0f) continue;
      const float rval = tile[ty + i][tx + j];
      float alpha = fabsf(rval - val0) * scale_index;
      const int idx = (int)floorf(alpha);
      alpha -= (float)idx;
      const float l0 = __ldg(st->lut + idx), l1 = __ldg(st->lut + idx + 1);
      const float w = kw * (l0 + alpha * (l1 - l0));
      wsum += w;
      sum += rval * w;
    }
  dst[y * W + x] = (sum + val0) / (wsum + 1.0f);
}

}  // namespace tfb
