// TEST INFRASTRUCTURE (oracle/_ref): the few OpenCV types the reference's Structure/Patch.{h,cpp} touch, so that
// the reference's own text compiles where OpenCV is absent.  Semantics restated from OpenCV's documentation:
// cv::Rect_<int> (member-wise truncating construction from floats, operator& = intersection, empty -> 0,0,0,0),
// cv::Mat as rows / cols / data with at<T>(y, x) on continuous storage.
#ifndef TF_CV_STANDIN_H
#define TF_CV_STANDIN_H
#include <algorithm>
#include <cstdlib>
#include <cstring>

#define CV_8U 0
#define CV_8UC3 16
#define CV_32F 5
#define CV_32FC1 5
#define CV_32FC3 21

namespace cv {
struct Vec3b {
  unsigned char v[3];
  unsigned char& operator[](int i) { return v[i]; }
  const unsigned char& operator[](int i) const { return v[i]; }
};
template <class T>
struct Rect_ {
  T x, y, width, height;
  Rect_() : x(0), y(0), width(0), height(0) {}
  Rect_(T x_, T y_, T w_, T h_) : x(x_), y(y_), width(w_), height(h_) {}
};
typedef Rect_<int> Rect;
template <class T>
inline Rect_<T> operator&(const Rect_<T>& a0, const Rect_<T>& b) {  // modules/core/include/opencv2/core/types.hpp
  Rect_<T> a = a0;
  const T x1 = std::max(a.x, b.x), y1 = std::max(a.y, b.y);
  a.width = std::min(a.x + a.width, b.x + b.width) - x1;
  a.height = std::min(a.y + a.height, b.y + b.height) - y1;
  a.x = x1;
  a.y = y1;
  if (a.width <= 0 || a.height <= 0) a = Rect_<T>();
  return a;
}
struct Mat {
  int rows = 0, cols = 0, type_ = 0;
  unsigned char* data = nullptr;
  unsigned char* datalimit = nullptr;  // == data: create() zero-fills, the reference's own `memset(data, 0, datalimit - data)`
  bool owner = false;                  // (Atlas.cpp:35-36, 573 MB) becomes a no-op instead of touching every page
  Mat() {}
  Mat(int r, int c, int t, void* p) : rows(r), cols(c), type_(t), data((unsigned char*)p) {}
  Mat(const Mat& o) : rows(o.rows), cols(o.cols), type_(o.type_), data(o.data), owner(false) {}  // header copy, like cv::Mat
  Mat& operator=(const Mat& o) {
    if (this != &o) { release(); rows = o.rows, cols = o.cols, type_ = o.type_, data = o.data, owner = false; }
    return *this;
  }
  ~Mat() { release(); }
  static size_t elem(int t) { return t == CV_8U ? 1 : t == CV_8UC3 ? 3 : t == CV_32F ? 4 : t == CV_32FC3 ? 12 : 0; }
  void release() {
    if (owner) free(data);
    data = nullptr, rows = cols = 0, owner = false;
  }
  void create(int r, int c, int t) {  // zero-filled (cv::Mat::create leaves memory uninitialised)
    release();
    rows = r, cols = c, type_ = t, owner = true;
    data = (unsigned char*)calloc((size_t)r * c, elem(t));
    datalimit = data;
  }
  template <class T>
  T& at(int y, int x) { return reinterpret_cast<T*>(data)[(size_t)y * cols + x]; }
  template <class T>
  const T& at(int y, int x) const { return reinterpret_cast<const T*>(data)[(size_t)y * cols + x]; }
  Mat operator()(const Rect&) const { return *this; }  // (ROI views are not needed by the compiled slices)
};
}  // namespace cv
#endif
