// TEST INFRASTRUCTURE — not part of the product.  Only tests/, __graft_entry__.smoke() and tools that
// time the CPU baseline may load the library built from this file; nothing under texturefusion_b200/ does.
//
// oracle/_ref/libtexfusion_ref_pre.so: the reference's OWN pre-processing loops
//   BasicAPI::extractNormalMapSIMD      BasicAPI.cpp:849-905
//   BasicAPI::refineDepthUseNormalSIMD  BasicAPI.cpp:728-780
//   BasicAPI::refineKeyframesSIMD       BasicAPI.cpp:506-636
//   BasicAPI::refineNewframesSIMD       BasicAPI.cpp:378-442
//   BasicAPI::checkColorQuality         BasicAPI.cpp:783-805
//   BasicAPI::estimateColorQuality      BasicAPI.cpp:814-847
// compiled from the text of /root/reference/BasicAPI.cpp (cut out by oracle/ref_pre_slices.py into a
// scratch file outside the repository, see that script) between the stand-in declarations below:
// `cv::Mat` reduced to rows / cols / data / create / release, `Sophus::SE3d` reduced to a 4x4 matrix
// and its given inverse, Eigen = oracle/eigen_standin.  `cv::cvtColor(RGB2GRAY)` and `cv::Sobel(1,1)`
// are restated here (OpenCV is a third-party dependency absent from the tree; the restatement is
// checked against cv2 in tests/test_pre_cpu.py).
//
// Two things the reference leaves to chance are DEFINED here, and the product follows the definition:
//   * cv::Mat::create does not clear memory; extractNormalMapSIMD never writes row 0, row H-1, column 0
//     and columns >= 633 (BasicAPI.cpp:866-867), checkColorQuality never writes the flags it does not set
//     (:800-802).  The stand-in's create() zero-fills.
//   * _mm256_rsqrt_ps is an approximate instruction; this library returns whatever the HOST returns
//     (tfp_host_rsqrt_matches() tells whether that is the Intel table the product reproduces).
#include <immintrin.h>

#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>
#include <vector>

#include <Eigen/Core>

#include "../texturefusion_b200/csrc/tf_rsqrt_table.h"

#define CV_8U 0
#define CV_32F 5
#define CV_32FC1 5
#define CV_32FC3 21
#define CV_RGB2GRAY 7

namespace cv {
struct Mat {
  int rows = 0, cols = 0, type_ = 0;
  unsigned char* data = nullptr;
  bool owner = false;
  Mat() {}
  Mat(int r, int c, int t, void* p) : rows(r), cols(c), type_(t), data((unsigned char*)p), owner(false) {}
  Mat(const Mat&) = delete;
  Mat& operator=(const Mat&) = delete;
  ~Mat() { release(); }
  static size_t elem(int t) { return t == CV_8U ? 1 : t == CV_32F ? 4 : t == CV_32FC3 ? 12 : 0; }
  void release() {
    if (owner) free(data);
    data = nullptr, rows = cols = 0, owner = false;
  }
  void create(int r, int c, int t) {  // (zero-filled: see the header)
    release();
    rows = r, cols = c, type_ = t, owner = true;
    data = (unsigned char*)calloc((size_t)r * c, elem(t));
  }
};

// cv::cvtColor(src, dst, CV_RGB2GRAY) for 8UC3: 15-bit fixed point (OpenCV >= 4.x color_rgb: R2Y 9798,
// G2Y 19235, B2Y 3735, rounding 1 << 14).  OpenCV 3.x used the 14-bit triple 4899 / 9617 / 1868.
inline void cvtColor(const Mat& src, Mat& dst, int) {
  dst.create(src.rows, src.cols, CV_8U);
  const size_t n = (size_t)src.rows * src.cols;
  for (size_t i = 0; i < n; i++)
    dst.data[i] = (unsigned char)((src.data[3 * i] * 9798 + src.data[3 * i + 1] * 19235 + src.data[3 * i + 2] * 3735 + (1 << 14)) >> 15);
}
// cv::Sobel(src 8U, dst, CV_32FC1, 1, 1): 3x3, kernel [-1 0 1]^T x [-1 0 1], BORDER_REFLECT_101
inline void Sobel(const Mat& src, Mat& dst, int, int, int) {
  const int W = src.cols, H = src.rows;
  dst.create(H, W, CV_32F);
  float* o = (float*)dst.data;
  auto at = [&](int y, int x) {
    y = y < 0 ? -y : (y >= H ? 2 * H - 2 - y : y);
    x = x < 0 ? -x : (x >= W ? 2 * W - 2 - x : x);
    return (int)src.data[(size_t)y * W + x];
  };
  for (int y = 0; y < H; y++)
    for (int x = 0; x < W; x++) o[(size_t)y * W + x] = (float)(at(y + 1, x + 1) - at(y + 1, x - 1) - at(y - 1, x + 1) + at(y - 1, x - 1));
}
}  // namespace cv

namespace Sophus {
// a rigid transform given as its matrix AND the matrix of its inverse (the caller provides both)
struct SE3d {
  Eigen::Matrix4d m = Eigen::Matrix4d::Identity(), minv = Eigen::Matrix4d::Identity();
  bool identity = true;
  SE3d inverse() const {
    SE3d r;
    r.m = minv, r.minv = m, r.identity = identity;
    return r;
  }
  SE3d operator*(const SE3d& o) const {
    if (identity) return o;
    if (o.identity) return *this;
    SE3d r;
    r.m = m * o.m, r.minv = o.minv * minv, r.identity = false;
    return r;
  }
  const Eigen::Matrix4d& matrix() const { return m; }
};
}  // namespace Sophus

struct Frame {
  cv::Mat rgb, refined_depth, weight;
  Sophus::SE3d pose_sophus[2];
};

namespace MultiViewGeometry {
class CameraPara {
 public:
  float c_fx, c_fy, c_cx, c_cy;
  float GetFx() { return c_fx; }
  float GetFy() { return c_fy; }
  float GetCx() { return c_cx; }
  float GetCy() { return c_cy; }
};
}  // namespace MultiViewGeometry

using namespace std;
using namespace cv;
using namespace MultiViewGeometry;

namespace BasicAPI {
#include TF_REF_PRE_SLICES
}  // namespace BasicAPI

namespace {
Sophus::SE3d from_rows(const float* T) {  // 3x4 row-major; used through .matrix() or .inverse().matrix() only
  Sophus::SE3d s;
  s.identity = false;
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 4; c++) s.m(r, c) = s.minv(r, c) = (double)T[r * 4 + c];
  return s;
}
}  // namespace

extern "C" {

const char* tfp_impl() { return "reference sources (BasicAPI.cpp slices) + cv/Sophus stand-ins"; }

void tfp_normal_map(const float* depth, float* normal, int W, int H, float fx, float fy, float cx, float cy) {
  cv::Mat d(H, W, CV_32F, (void*)depth), n;
  BasicAPI::extractNormalMapSIMD(d, n, fx, fy, cx, cy);
  memcpy(normal, n.data, (size_t)W * H * 12);
}

void tfp_refine_depth_by_normal(float* normal, float* depth, int W, int H, float fx, float fy, float cx, float cy) {
  BasicAPI::refineDepthUseNormalSIMD(normal, depth, fx, fy, cx, cy, (float)W, (float)H);
}

// T: ref -> new, 3x4 row-major
void tfp_refine_keyframe(float* kf_depth, float* kf_weight, const float* new_depth, const float* T, int W, int H, float fx, float fy,
                         float cx, float cy) {
  Frame ref, nw;
  ref.rgb.rows = H, ref.rgb.cols = W;
  ref.refined_depth.data = (unsigned char*)kf_depth;
  ref.weight.data = (unsigned char*)kf_weight;
  nw.refined_depth.data = (unsigned char*)new_depth;
  nw.pose_sophus[0] = from_rows(T).inverse();  // (new.inverse() * ref).matrix() == T
  CameraPara cam{fx, fy, cx, cy};
  BasicAPI::refineKeyframesSIMD(ref, nw, cam);
}

// T: new -> ref, 3x4 row-major
void tfp_refine_newframe(const float* kf_depth, float* new_depth, const float* T, int W, int H, float fx, float fy, float cx, float cy) {
  Frame ref, nw;
  ref.rgb.rows = H, ref.rgb.cols = W;
  ref.refined_depth.data = (unsigned char*)kf_depth;
  nw.refined_depth.data = (unsigned char*)new_depth;
  nw.pose_sophus[0] = from_rows(T);  // (ref.inverse() * new).matrix() == T
  CameraPara cam{fx, fy, cx, cy};
  BasicAPI::refineNewframesSIMD(ref, nw, cam);
}

void tfp_color_valid(const float* normal, uint8_t* flag, int W, int H, float fx, float fy, float cx, float cy) {
  cv::Mat n(H, W, CV_32FC3, (void*)normal), f;
  BasicAPI::checkColorQuality(n, f, fx, fy, cx, cy);
  memcpy(flag, f.data, (size_t)W * H);
}

void tfp_color_quality(const float* depth, const float* normal, const uint8_t* rgb, float* quality, int W, int H, float fx, float fy,
                       float cx, float cy) {
  cv::Mat d(H, W, CV_32F, (void*)depth), n(H, W, CV_32FC3, (void*)normal), c(H, W, CV_8U, (void*)rgb), q;
  BasicAPI::estimateColorQuality(d, n, q, c, fx, fy, cx, cy);
  memcpy(quality, q.data, (size_t)W * H * 4);
}

void tfp_gray(const uint8_t* rgb, uint8_t* gray, int W, int H) {
  cv::Mat c(H, W, CV_8U, (void*)rgb), g;
  cv::cvtColor(c, g, CV_RGB2GRAY);
  memcpy(gray, g.data, (size_t)W * H);
}

void tfp_sobel11(const uint8_t* gray, float* out, int W, int H) {
  cv::Mat g(H, W, CV_8U, (void*)gray), o;
  cv::Sobel(g, o, CV_32FC1, 1, 1);
  memcpy(out, o.data, (size_t)W * H * 4);
}

uint32_t tfp_rsqrt_bits(uint32_t b) {  // the host's instruction
  float f;
  memcpy(&f, &b, 4);
  const float o = _mm_cvtss_f32(_mm_rsqrt_ss(_mm_set_ss(f)));
  uint32_t r;
  memcpy(&r, &o, 4);
  return r;
}

// 1 when the host's RSQRTPS returns the tabulated (Intel) values; checked on a stride of the positive normals
int tfp_host_rsqrt_matches() {
  for (uint32_t b = 0x00800000u; b < 0x7f800000u; b += 4099u) {
    const uint32_t e = b >> 23, m = b & 0x7fffffu, p = e & 1u;
    const int k = ((int)e - (p ? 127 : 128)) / 2;
    const uint32_t want = 0x3f000000u + ((uint32_t)kRsqrtTabHost[p * 1024 + (m >> 13)] << 11) - ((uint32_t)k << 23);
    if (tfp_rsqrt_bits(b) != want) return 0;
  }
  return 1;
}

}  // extern "C"
