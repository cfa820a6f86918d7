// Host-side check (no GPU) of the range test behind integrate_kernel's division-free projections
// (texturefusion_b200/csrc/tf_host_math.h: FrameDev::z_safe; tf_device.cuh: project_safe).
//
// For random rigid poses and voxel sizes the centroid table the kernel builds — (Rt (x,y,z)) res + res/2 in float,
// x,y,z in 0..7, either association of the 3-term product — is recomputed here, and for chunk origins just above
// z_safe every voxel centre must have 2^-17 < cz < 2^21 and |c| < 2^21: the operand range in which the inline
// quotient sequence is exact.  Prints "ok <cases>" or the first violation.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include "../../texturefusion_b200/csrc/tf_host_math.h"

int main() {
  std::mt19937 rng(7);
  std::uniform_real_distribution<float> uni(-1.0f, 1.0f);
  long cases = 0;
  const float resolutions[] = {0.04f, 0.02f, 0.01f, 0.005f, 0.0025f, 1.0f, 37.5f};
  for (int trial = 0; trial < 400; trial++) {
    // random rotation from a unit quaternion, random translation
    float q[4];
    float n2 = 0;
    for (float& v : q) { v = uni(rng); n2 += v * v; }
    const float inv = 1.0f / std::sqrt(n2);
    for (float& v : q) v *= inv;
    const float w = q[0], x = q[1], y = q[2], z = q[3];
    const float R[9] = {1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
                        2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
                        2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)};
    tf_pose pose{};
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) pose.m[j * 4 + i] = R[i * 3 + j];  // column-major camera->world
    pose.m[12] = 3 * uni(rng), pose.m[13] = 3 * uni(rng), pose.m[14] = 3 * uni(rng), pose.m[15] = 1;
    tf_camera cam{};
    cam.fx = 525, cam.fy = 525, cam.cx = 319.5f, cam.cy = 239.5f, cam.width = 640, cam.height = 480;
    cam.near_plane = 0.01f, cam.far_plane = 5.0f;
    for (float res : resolutions) {
      tfb::FrameDev f;
      tfb::make_frame_dev(pose, cam, res, 1, nullptr, nullptr, nullptr, f);
      if (!(f.z_safe > 0) || !std::isfinite(f.z_safe)) { std::printf("z_safe not finite for res %g\n", res); return 1; }
      const float half = res * 0.5f;
      for (int l2r = 0; l2r < 2; l2r++) {
        // extreme origins the per-chunk test admits: depth just above z_safe, lateral coordinates just inside 2^20
        const float o2s[] = {std::nextafter(f.z_safe, INFINITY), f.z_safe * 1.5f, 1048575.0f};
        const float o01s[] = {0.0f, 1048575.0f, -1048575.0f};
        for (float o2 : o2s)
          for (float o0 : o01s)
            for (int v = 0; v < 512; v++) {
              const float xf = (float)(v & 7), yf = (float)((v >> 3) & 7), zf = (float)(v >> 6);
              float c[3];
              const float o[3] = {o0, -o0, o2};
              for (int k = 0; k < 3; k++) {
                const float m = tfb::h_dot3(l2r, f.Rt[k * 3], xf, f.Rt[k * 3 + 1], yf, f.Rt[k * 3 + 2], zf);
                c[k] = o[k] + (m * res + half);
              }
              if (!(c[2] > 0x1p-17f && c[2] < 0x1p21f && std::fabs(c[0]) < 0x1p21f && std::fabs(c[1]) < 0x1p21f)) {
                std::printf("violation: res %g o2 %g voxel %d -> c = %g %g %g (z_safe %g)\n", res, o2, v, c[0], c[1], c[2], f.z_safe);
                return 1;
              }
              cases++;
            }
      }
    }
  }
  // a transform that is not rigid (scale 1e7) or intrinsics out of range switch the test off: z_safe = inf
  tf_pose big{};
  big.m[0] = big.m[5] = big.m[10] = 1e7f, big.m[15] = 1;
  tf_camera cam{};
  cam.fx = 525, cam.fy = 525, cam.width = 640, cam.height = 480;
  tfb::FrameDev f;
  tfb::make_frame_dev(big, cam, 0.005f, 1, nullptr, nullptr, nullptr, f);
  if (std::isfinite(f.z_safe) && f.z_safe < 1e5f) { std::printf("scaled transform: z_safe %g does not cover the table\n", f.z_safe); return 1; }
  cam.fx = 4e6f;
  tf_pose id{};
  id.m[0] = id.m[5] = id.m[10] = id.m[15] = 1;
  tfb::make_frame_dev(id, cam, 0.005f, 1, nullptr, nullptr, nullptr, f);
  if (std::isfinite(f.z_safe)) { std::printf("focal length 4e6: the test should be switched off\n"); return 1; }
  std::printf("ok %ld\n", cases);
  return 0;
}
