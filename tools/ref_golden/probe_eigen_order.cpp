// probe_eigen_order.cpp — prints the association order an installed Eigen uses for the 3-term
// products on the fusion path, i.e. the value tf_config.dot3_order must have to match a reference
// built against that Eigen.  Needs the real Eigen (not oracle/eigen_standin):
//   g++ -O3 -mavx2 -mno-fma -ffp-contract=off -I/usr/include/eigen3 probe_eigen_order.cpp -o probe && ./probe
// The three expressions are the ones the reference evaluates:
//   (a) Matrix3f^T * Vector3f        ProjectionIntegrator.cpp:88-89, Structure/Chisel.cpp:67-69
//   (b) MatrixXf * Vector3f          Structure/ChunkManager.h:429-453, 521-524
//   (c) MatrixXf * Vector3f -> VectorXf   Structure/ChunkManager.h:430
// Operands are chosen so that (x0 + x1) + x2 != x0 + (x1 + x2) in binary32.
#include <Eigen/Core>
#include <Eigen/Geometry>
#include <cstdio>
#include <cstring>

static unsigned bits(float f) { unsigned u; std::memcpy(&u, &f, 4); return u; }

int main() {
  Eigen::Affine3f T = Eigen::Affine3f::Identity();
  // first column of linear() = first row of its transpose: (1, 2^-24, 2^-24) against v = (1, 1, 1)
  //   left to right: (1 + 2^-24) + 2^-24 = 1            (each addend is half an ulp: ties to even, twice)
  //   tree:          1 + (2^-24 + 2^-24) = 1 + 2^-23
  const float e = 5.9604644775390625e-8f;  // 2^-24
  T.linear()(0, 0) = 1.0f; T.linear()(1, 0) = e; T.linear()(2, 0) = e;
  const Eigen::Vector3f v(1.0f, 1.0f, 1.0f);
  const float l2r = (1.0f + e) + e, tree = 1.0f + (e + e);
  const Eigen::Vector3f a = T.linear().transpose() * v;
  Eigen::MatrixXf R = T.linear().transpose();
  const Eigen::Vector3f b = R * v;
  const Eigen::VectorXf c = R * v;
  const char* name[3] = {"(a) Matrix3f^T * Vector3f", "(b) MatrixXf * Vector3f", "(c) MatrixXf * Vector3f -> VectorXf"};
  const float got[3] = {a(0), b(0), c(0)};
  std::printf("Eigen %d.%d.%d\n", EIGEN_WORLD_VERSION, EIGEN_MAJOR_VERSION, EIGEN_MINOR_VERSION);
  int order = -1;
  bool consistent = true;
  for (int i = 0; i < 3; i++) {
    const int o = bits(got[i]) == bits(tree) ? 0 : bits(got[i]) == bits(l2r) ? 1 : -1;
    std::printf("%-40s %s\n", name[i], o == 0 ? "x0 + (x1 + x2)  -> dot3_order 0" : o == 1 ? "(x0 + x1) + x2  -> dot3_order 1" : "neither (fused multiply-add? check the flags)");
    if (i == 0) order = o;
    consistent = consistent && o == order;
  }
  if (!consistent) std::printf("WARNING: the expressions disagree; texfusion-b200 uses one order for all of them\n");
  std::printf("tf_config.dot3_order = %d\n", order);
  return consistent && order >= 0 ? 0 : 1;
}
