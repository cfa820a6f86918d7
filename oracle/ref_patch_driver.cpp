// TEST INFRASTRUCTURE — not part of the product.  Only tests/ may load the library built from this file.
//
// oracle/_ref/libtexfusion_ref_patch.so: the reference's OWN texture-coordinate path
//   Patch::Patch, Patch::CalculateTexCoords, Patch::bilinear, Patch::bilinear_depth   Structure/Patch.cpp:27-170
// compiled from the text of /root/reference/Structure/Patch.cpp (cut out by oracle/ref_pre_slices.py into a
// scratch file outside the repository) against the reference's own Structure/Patch.h, geometry/Mesh.h and
// camera/PinholeCamera.{h,cpp}; <opencv2/opencv.hpp> and GCSLAM/frame.h resolve to oracle/cv_standin, Eigen to
// oracle/eigen_standin.  Exports tfo_patch_texcoords with the signature of the restatement in tf_oracle.cpp,
// which tests/test_ref_cpu.py requires to be bit-identical.
//
// What the stand-ins decide: cv::Rect's truncating construction and operator& (restated from OpenCV), and
// Eigen's evaluation order of Matrix4f * Vector4f (packet-wise, left to right) and of Vector3f::norm().
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

#include "Patch.h"

namespace chisel {
#include TF_REF_PATCH_SLICES
}  // namespace chisel

namespace {
struct Cam {
  float fx, fy, cx, cy;
  int32_t width, height;
  float near_plane, far_plane;
};
struct PatchResult {
  int32_t x, y, w, h, wrong_mapping, flag;
};
}  // namespace

extern "C" {

const char* tfo_patch_impl() { return "reference sources (Structure/Patch.cpp slices) + cv/Sophus/Eigen stand-ins"; }

// T: world->camera 4x4 (column-major), i.e. pose_sophus[0].inverse().matrix().cast<float>()
int tfo_patch_texcoords(const uint8_t* rgb, const float* depth, const float* T, const Cam* cam, int64_t n_patches,
                        const int64_t* offsets, const float* verts, const float* colors, float* texcoord, float* texcolor,
                        PatchResult* results) {
  chisel::PinholeCamera camera;
  camera.SetIntrinsics(cam->fx, cam->fy, cam->cx, cam->cy);
  camera.SetWidth(cam->width), camera.SetHeight(cam->height);
  camera.SetNearPlane(cam->near_plane), camera.SetFarPlane(cam->far_plane);
  // A vertex clamped to x = width or y = height makes bilinear / bilinear_depth read pixel index y * cols + x beyond
  // the last row (Patch.cpp:135-136, :168-169: cv::Mat::at does not check bounds).  What lies there is undefined
  // in the reference; the images are copied into buffers followed by one zeroed row (+ 2 pixels) so that it is 0 —
  // the definition the restatement and the device kernel use.
  const size_t np = (size_t)cam->width * cam->height, pad = (size_t)cam->width + 2;
  std::vector<uint8_t> rgb_p((np + pad) * 3, 0);
  std::vector<float> depth_p(np + pad, 0.0f);
  memcpy(rgb_p.data(), rgb, np * 3);
  memcpy(depth_p.data(), depth, np * 4);
  Frame fr;
  fr.rgb = cv::Mat(cam->height, cam->width, CV_8UC3, (void*)rgb_p.data());
  fr.refined_depth = cv::Mat(cam->height, cam->width, CV_32F, (void*)depth_p.data());
  fr.pose_sophus[0].identity = false;
  for (int c = 0; c < 4; c++)
    for (int r = 0; r < 4; r++) fr.pose_sophus[0].minv(r, c) = (double)T[c * 4 + r];  // .inverse().matrix() returns this
  for (int64_t p = 0; p < n_patches; p++) {
    chisel::MeshPtr mesh = std::make_shared<chisel::Mesh>();
    for (int64_t i = offsets[p]; i < offsets[p + 1]; i++) {
      mesh->vertices.push_back(chisel::Vec3(verts[3 * i], verts[3 * i + 1], verts[3 * i + 2]));
      mesh->colors.push_back(chisel::Vec3(colors[3 * i], colors[3 * i + 1], colors[3 * i + 2]));
    }
    chisel::Patch patch(mesh);
    const int flag = patch.CalculateTexCoords(fr, camera);
    for (int64_t i = offsets[p]; i < offsets[p + 1]; i++) {
      const int64_t k = i - offsets[p];
      texcoord[2 * i] = patch.texcoord[k](0), texcoord[2 * i + 1] = patch.texcoord[k](1);
      for (int c = 0; c < 3; c++) texcolor[3 * i + c] = patch.texcolor[k](c);
    }
    results[p] = PatchResult{patch.boundingbox.x, patch.boundingbox.y, patch.boundingbox.width, patch.boundingbox.height,
                             patch.wrong_mapping ? 1 : 0, flag};
  }
  return 0;
}

}  // extern "C"
