// TEST INFRASTRUCTURE (oracle/_ref): the input record of the texture path, reduced to the fields
// Structure/Patch.cpp reads (GCSLAM/frame.h:35-66 in the reference): rgb, refined_depth, pose_sophus.
#ifndef TF_FRAME_STANDIN_H
#define TF_FRAME_STANDIN_H
#include <Eigen/Core>
#include <opencv2/opencv.hpp>
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
  int frame_index = 0;
};
#endif
