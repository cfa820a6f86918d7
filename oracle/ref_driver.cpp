// ref_driver.cpp — C driver around the REFERENCE'S OWN fusion sources (test infrastructure only).
//
// oracle/_ref/libtexfusion_ref.so is built by oracle/Makefile from the files where they lie under
// /root/reference (never copied into this repository):
//   3rd_party/open_chisel/utils/ProjectionIntegrator.cpp   voxelUpdateSIMD (:67-426)
//   Structure/ChunkManager.{h,cpp}                         findCubeCornerByMat, GetChunkIDsObservedByCamera,
//                                                          CheckCornerIntersectingSIMD, CreateChunk, GenerateMeshEfficient,
//                                                          extractGradientFromCubic
//   3rd_party/open_chisel/geometry/{Chunk,ColorVoxel,DistVoxel,AABB,Plane,Frustum,Mesh}.cpp
//   3rd_party/open_chisel/camera/{PinholeCamera,Intrinsics}.cpp, marching_cubes/MarchingCubes.cpp
//   + headers: truncation/QuadraticTruncator.h, weighting/ConstantWeighter.h, threading/Threading.h
// against oracle/eigen_standin (Eigen itself is absent from this image; the stand-in's header states
// what it restates and the one freedom it has: the association order of 3-term products).
//
// Structure/Chisel.{h,cpp} need OpenCV and Sophus (via Atlas.h, Patch.h, GCSLAM/frame.h); the few lines of
// glue Chisel.h adds on this path are restated below, each citing the lines it follows.  Everything
// voxel- or pixel-sized runs in the reference's code.
//
// -DTF_REF_REAL_CHISEL (oracle/_ref/libtexfusion_ref_chisel.so) builds the same driver around the reference's OWN
// chisel::Chisel: Structure/Chisel.h compiled with the stand-in <opencv2/opencv.hpp> / GCSLAM/frame.h of
// oracle/cv_standin, its out-of-line constructor / Reset / bufferIntegratorSIMDCentroids and Atlas's constructor cut
// out of Chisel.cpp / Atlas.cpp (oracle/ref_pre_slices.py chisel).  prepare / integrate / finalize then ARE
// Chisel::PrepareIntersectChunks, IntegrateDepthScanColor (list form) and FinalizeIntegrateChunks; tests/test_ref_cpu.py
// requires that build and the restated glue to agree (on lists short enough for the reference's parallel_for to stay
// on one thread: its concurrent std::vector<bool> writes are a race a checker must not run).
//
// The exported tfo_* symbols have the signatures of oracle/tf_oracle.cpp, so oracle.py drives
// either library.  NOTE ChunkManager::GetIDAt caches 1/(8*res) in function-local statics
// (Structure/ChunkManager.h:197-203): one voxel resolution per loaded copy of this library
// (oracle.py loads one copy per resolution).
#include <cstdint>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

#ifdef TF_REF_REAL_CHISEL
#include "Chisel.h"
namespace chisel {
#include TF_REF_CHISEL_SLICES
}  // namespace chisel
#endif
#include "ChunkManager.h"
#include "camera/PinholeCamera.h"
#include "geometry/Chunk.h"
#include "geometry/Frustum.h"
#include "threading/Threading.h"
#include "truncation/QuadraticTruncator.h"
#include "utils/ProjectionIntegrator.h"
#include "weighting/ConstantWeighter.h"

using namespace chisel;

namespace {

struct Cam {  // == tf_camera
  float fx, fy, cx, cy;
  int32_t width, height;
  float near_plane, far_plane;
};

struct RefMap {
  EIGEN_MAKE_ALIGNED_OPERATOR_NEW
#ifdef TF_REF_REAL_CHISEL
  Chisel chisel;                 // the reference's own object; the two references below are its public fields
  ChunkManager& chunkManager;    // (Structure/Chisel.h:488-489)
  ChunkSet& meshesToUpdate;
#else
  ChunkManager chunkManager;
  ChunkSet meshesToUpdate;
#endif
  ProjectionIntegrator integrator;
  float res;
  int threads;  // 1: serial loop; 0: the reference's parallel_for policy
  RefMap(float r, const float* t5, int th)
#ifdef TF_REF_REAL_CHISEL
      : chisel(Eigen::Vector3i(8, 8, 8), r, true), chunkManager(chisel.chunkManager), meshesToUpdate(chisel.meshesToUpdate), res(r), threads(th) {
#else
      : chunkManager(Eigen::Vector3i(8, 8, 8), r, true), res(r), threads(th) {
#endif
    // MobileFusion::initChiselMap (GCFusion/MobileFusion.h:243-251)
    integrator.SetCentroids(chunkManager.GetCentroids());
    integrator.SetTruncator(TruncatorPtr(new QuadraticTruncator(t5[0], t5[1], t5[2], t5[3])));
    integrator.SetWeighter(WeighterPtr(new ConstantWeighter(t5[4])));
    integrator.SetCarvingDist(0.05f);
    integrator.SetCarvingEnabled(true);
  }
};

PinholeCamera make_camera(const Cam* c) {  // GCFusion/MobileFusion.h:253-257
  PinholeCamera cam;
  cam.SetIntrinsics(c->fx, c->fy, c->cx, c->cy);
  cam.SetNearPlane(c->near_plane);
  cam.SetFarPlane(c->far_plane);
  cam.SetWidth(c->width);
  cam.SetHeight(c->height);
  return cam;
}

Transform make_pose(const float* m16) {  // column-major 4x4, camera -> world
  Transform T;
  for (int i = 0; i < 3; i++) {
    for (int j = 0; j < 3; j++) T.linear()(i, j) = m16[j * 4 + i];
    T.translation()(i) = m16[12 + i];
  }
  return T;
}

// Chisel::bufferIntegratorSIMDCentroids (Structure/Chisel.cpp:52-110), restated: Chisel.cpp itself
// needs OpenCV.  The expression of :67-69 is kept verbatim so that it evaluates through the same
// (stand-in) Eigen operators.
void buffer_centroids(RefMap* m, const Transform& depthExtrinsic) {
#ifdef TF_REF_REAL_CHISEL
  m->chisel.bufferIntegratorSIMDCentroids(m->integrator, depthExtrinsic);
  return;
#endif
  ChunkManager& chunkManager = m->chunkManager;
  ProjectionIntegrator& integrator = m->integrator;
  Vec3 halfVoxel = Vec3(chunkManager.GetResolution(), chunkManager.GetResolution(), chunkManager.GetResolution()) * 0.5f;
  Vec3List diff_centroids;
  diff_centroids.resize(512);
  int i = 0;
  for (int z = 0; z < 8; z++)
    for (int y = 0; y < 8; y++)
      for (int x = 0; x < 8; x++) {
        diff_centroids[i] = depthExtrinsic.linear().transpose() * Vec3(x, y, z) * chunkManager.GetResolution() + halfVoxel;
        i++;
      }
  integrator.diff_centroids = diff_centroids;
  float data0[8], data1[8], data2[8];
  for (int pos = 0; pos < 512; pos += 8) {  // :83-107: 8 x float3 -> 3 x float8
    for (int a = 0; a < 8; a++) {
      data0[a] = integrator.diff_centroids[pos + a](0);
      data1[a] = integrator.diff_centroids[pos + a](1);
      data2[a] = integrator.diff_centroids[pos + a](2);
    }
    integrator.centroids_simd0[pos / 8] = _mm256_loadu_ps(data0);
    integrator.centroids_simd1[pos / 8] = _mm256_loadu_ps(data1);
    integrator.centroids_simd2[pos / 8] = _mm256_loadu_ps(data2);
  }
}

// Chisel::PrepareIntersectChunks (Structure/Chisel.h:103-140)
void prepare(RefMap* m, float* depthImage, const Transform& depthExtrinsic, const PinholeCamera& depthCamera,
             ChunkIDList& chunksIntersecting, std::vector<bool>& needsUpdateFlag, std::vector<bool>& newChunkFlag) {
#ifdef TF_REF_REAL_CHISEL
  m->chisel.PrepareIntersectChunks(m->integrator, depthImage, depthExtrinsic, depthCamera, chunksIntersecting, needsUpdateFlag,
                                   newChunkFlag);
  return;
#endif
  buffer_centroids(m, depthExtrinsic);
  chunksIntersecting.clear();
  needsUpdateFlag.clear();
  newChunkFlag.clear();
  ChunkID maxChunkID, minChunkID;
  m->chunkManager.GetBoundaryChunkID(depthImage, depthCamera, depthExtrinsic, maxChunkID, minChunkID);
  Frustum frustum;
  depthCamera.SetupFrustum(depthExtrinsic, &frustum);
  m->chunkManager.GetChunkIDsObservedByCamera(m->integrator, frustum, &chunksIntersecting, depthImage, depthCamera,
                                              depthExtrinsic);
  for (const ChunkID& chunkID : chunksIntersecting) {
    bool chunkNew = false;
    if (!m->chunkManager.HasChunk(chunkID)) {
      chunkNew = true;
      m->chunkManager.CreateChunk(chunkID);
    }
    newChunkFlag.emplace_back(chunkNew);
    needsUpdateFlag.emplace_back(false);
  }
}

// Chisel::IntegrateDepthScanColor, list form (Structure/Chisel.h:218-249).  needsUpdateFlag is a
// byte vector here: the reference's concurrent writes to std::vector<bool> are a latent race
//  that a test oracle must not reproduce.  q_out (optional) = raw chunkObservationQuality.
void integrate(RefMap* m, float* depthImage, unsigned char* colorImage, const Transform& depthExtrinsic,
               const PinholeCamera& depthCamera, ChunkIDList& chunksIntersecting, std::vector<uint8_t>& needsUpdateFlag,
               int integrate_flag, int keyframeID, float* observationQualityPointer, float* q_out) {
#ifdef TF_REF_REAL_CHISEL
  {  // (q_out, the raw per-chunk quality sum, does not leave the reference's function: reported as 0)
    std::vector<bool> nu(needsUpdateFlag.begin(), needsUpdateFlag.end());
    m->chisel.IntegrateDepthScanColor(m->integrator, depthImage, colorImage, depthExtrinsic, depthCamera, chunksIntersecting, nu,
                                      integrate_flag, keyframeID, observationQualityPointer);
    for (size_t i = 0; i < nu.size(); i++) needsUpdateFlag[i] = nu[i] ? 1 : 0;
    if (q_out) std::fill(q_out, q_out + chunksIntersecting.size(), 0.0f);
    return;
  }
#endif
  buffer_centroids(m, depthExtrinsic);
  if (chunksIntersecting.size() < 1) return;
  std::vector<int> threadIndex;
  for (int i = 0; i < (int)chunksIntersecting.size(); i++) threadIndex.emplace_back(i);
  auto body = [&](const int& i) {
    ChunkID chunkID = chunksIntersecting[i];
    const ChunkPtr& chunk = m->chunkManager.GetChunk(chunkID);
    float chunkObservationQuality;
    bool needsUpdate = m->integrator.IntegrateColor(depthImage, depthCamera, depthExtrinsic, colorImage, chunk.get(),
                                                    integrate_flag, chunkObservationQuality, observationQualityPointer);
    needsUpdateFlag[i] = (needsUpdateFlag[i] || needsUpdate);
    if (q_out) q_out[i] = chunkObservationQuality;
    if (keyframeID >= 0 && chunkObservationQuality > 0 && needsUpdateFlag[i]) chunk->observations[keyframeID] = chunkObservationQuality;
  };
  if (m->threads == 1) std::for_each(threadIndex.begin(), threadIndex.end(), body);
  else parallel_for(threadIndex.begin(), threadIndex.end(), body);  // 3rd_party/open_chisel/threading/Threading.h:35-53
}

// Chisel::FinalizeIntegrateChunks + GarbageCollect (Structure/Chisel.h:184-216, 472-477)
void finalize(RefMap* m, ChunkIDList& chunksIntersecting, std::vector<uint8_t>& needsUpdateFlag,
              std::vector<bool>& newChunkFlag, ChunkIDList& validChunks) {
#ifdef TF_REF_REAL_CHISEL
  {
    std::vector<bool> nu(needsUpdateFlag.begin(), needsUpdateFlag.end());
    m->chisel.FinalizeIntegrateChunks(chunksIntersecting, nu, newChunkFlag, validChunks);
    return;
  }
#endif
  validChunks.clear();
  ChunkIDList garbageChunks;
  for (int i = 0; i < (int)chunksIntersecting.size(); i++) {
    const ChunkID& chunkID = chunksIntersecting[i];
    if (needsUpdateFlag[i]) {
      m->meshesToUpdate[chunkID] = true;
      m->meshesToUpdate[chunkID + ChunkID(-1, 0, 0)] = true;
      m->meshesToUpdate[chunkID + ChunkID(1, 0, 0)] = true;
      m->meshesToUpdate[chunkID + ChunkID(0, -1, 0)] = true;
      m->meshesToUpdate[chunkID + ChunkID(0, 1, 0)] = true;
      m->meshesToUpdate[chunkID + ChunkID(0, 0, -1)] = true;
      m->meshesToUpdate[chunkID + ChunkID(0, 0, 1)] = true;
      validChunks.emplace_back(chunkID);
    } else if (newChunkFlag[i]) {
      garbageChunks.emplace_back(chunkID);
    }
  }
  for (const ChunkID& chunkID : garbageChunks) {
    m->chunkManager.RemoveChunk(chunkID);
    m->meshesToUpdate.erase(chunkID);
  }
}

ChunkIDList to_list(const int32_t* ids, int64_t n) {
  ChunkIDList l;
  l.reserve((size_t)n);
  for (int64_t i = 0; i < n; i++) l.emplace_back(ChunkID(ids[3 * i], ids[3 * i + 1], ids[3 * i + 2]));
  return l;
}

}  // namespace

extern "C" {

struct tfo_map;

const char* tfo_impl() {
#ifdef TF_REF_REAL_CHISEL
  return "reference sources incl. chisel::Chisel (Chisel.h + Chisel.cpp / Atlas.cpp slices) + Eigen / cv stand-ins";
#elif defined(EIGEN_STANDIN_LEFT_TO_RIGHT)
  return "reference sources + Eigen stand-in (left-to-right products)";
#else
  return "reference sources + Eigen stand-in (tree-reduced products)";
#endif
}

tfo_map* tfo_create(float res, const float* trunc5, int threads) { return (tfo_map*)new RefMap(res, trunc5, threads); }
void tfo_destroy(tfo_map* h) { delete (RefMap*)h; }
int tfo_threads_used(tfo_map* h) {
  RefMap* m = (RefMap*)h;
  if (m->threads == 1) return 1;
  return std::max(1, (int)std::thread::hardware_concurrency() - 2);  // Threading.h:39
}
void tfo_reset(tfo_map* h) {  // Chisel::Reset (Structure/Chisel.cpp:47-50)
  RefMap* m = (RefMap*)h;
#ifdef TF_REF_REAL_CHISEL
  m->chisel.Reset();
  return;
#endif
  m->chunkManager.Reset();
  m->meshesToUpdate.clear();
}

void tfo_boundary_ids(tfo_map* h, const float* depth, const float* pose, const Cam* cam, int32_t* min3, int32_t* max3) {
  RefMap* m = (RefMap*)h;
  ChunkID maxID, minID;
  m->chunkManager.GetBoundaryChunkID(depth, make_camera(cam), make_pose(pose), maxID, minID);
  for (int k = 0; k < 3; k++) { min3[k] = minID(k); max3[k] = maxID(k); }
}

int64_t tfo_observed_ids(tfo_map* h, const float* depth, const float* pose, const Cam* cam, int32_t* ids_out, int64_t cap) {
  RefMap* m = (RefMap*)h;
  ChunkIDList l;
  Frustum frustum;
  PinholeCamera c = make_camera(cam);
  Transform T = make_pose(pose);
  c.SetupFrustum(T, &frustum);
  m->chunkManager.GetChunkIDsObservedByCamera(m->integrator, frustum, &l, depth, c, T);
  for (int64_t i = 0; i < (int64_t)l.size() && i < cap; i++)
    for (int k = 0; k < 3; k++) ids_out[3 * i + k] = l[i](k);
  return (int64_t)l.size();
}

int64_t tfo_prepare(tfo_map* h, const float* depth, const float* pose, const Cam* cam, int32_t* ids_out, uint8_t* new_out,
                    int64_t cap) {
  RefMap* m = (RefMap*)h;
  ChunkIDList l;
  std::vector<bool> nu, nw;
  prepare(m, const_cast<float*>(depth), make_pose(pose), make_camera(cam), l, nu, nw);
  if ((int64_t)l.size() > cap) return -1;
  for (size_t i = 0; i < l.size(); i++) {
    for (int k = 0; k < 3; k++) ids_out[3 * i + k] = l[i](k);
    new_out[i] = nw[i] ? 1 : 0;
  }
  return (int64_t)l.size();
}

int tfo_integrate(tfo_map* h, const float* depth, const uint8_t* rgba, const float* quality, const float* pose,
                  const Cam* cam, const int32_t* ids, int64_t n, int flag, int keyframe_id, uint8_t* needs_update,
                  float* q_out) {
  RefMap* m = (RefMap*)h;
  ChunkIDList l = to_list(ids, n);
  for (const ChunkID& id : l)
    if (!m->chunkManager.HasChunk(id)) return -4;  // ChunkMap::at would throw (Structure/ChunkManager.h:137-139)
  std::vector<uint8_t> nu(needs_update, needs_update + n);
  integrate(m, const_cast<float*>(depth), const_cast<uint8_t*>(rgba), make_pose(pose), make_camera(cam), l, nu, flag,
            keyframe_id, const_cast<float*>(quality), q_out);
  for (int64_t i = 0; i < n; i++) needs_update[i] = nu[i];
  return 0;
}

int64_t tfo_finalize(tfo_map* h, const int32_t* ids, int64_t n, const uint8_t* needs_update, const uint8_t* is_new,
                     int32_t* valid_out) {
  RefMap* m = (RefMap*)h;
  ChunkIDList l = to_list(ids, n), valid;
  std::vector<uint8_t> nu(needs_update, needs_update + n);
  std::vector<bool> nw((size_t)n);
  for (int64_t i = 0; i < n; i++) nw[i] = is_new[i] != 0;
  finalize(m, l, nu, nw, valid);
  if (valid_out)
    for (size_t i = 0; i < valid.size(); i++)
      for (int k = 0; k < 3; k++) valid_out[3 * i + k] = valid[i](k);
  return (int64_t)valid.size();
}

// Chisel::IntegrateDepthScanColor, convenience form (Structure/Chisel.h:453-468), with the key-frame
// arguments of the list form so that one call covers MobileFusion::IntegrateFrame and the benchmark.
int64_t tfo_integrate_frame(tfo_map* h, const float* depth, const uint8_t* rgba, const float* quality, const float* pose,
                            const Cam* cam, int keyframe_id, int64_t* n_updated) {
  RefMap* m = (RefMap*)h;
  ChunkIDList l, valid;
  std::vector<bool> nu0, nw;
  PinholeCamera c = make_camera(cam);
  Transform T = make_pose(pose);
  prepare(m, const_cast<float*>(depth), T, c, l, nu0, nw);
  std::vector<uint8_t> nu(l.size(), 0);
  integrate(m, const_cast<float*>(depth), const_cast<uint8_t*>(rgba), T, c, l, nu, 1, keyframe_id,
            const_cast<float*>(quality), nullptr);
  finalize(m, l, nu, nw, valid);
  if (n_updated) *n_updated = (int64_t)valid.size();
  return (int64_t)l.size();
}

int tfo_has_chunk(tfo_map* h, int32_t x, int32_t y, int32_t z) { return ((RefMap*)h)->chunkManager.HasChunk(ChunkID(x, y, z)) ? 1 : 0; }
int64_t tfo_chunk_count(tfo_map* h) { return (int64_t)((RefMap*)h)->chunkManager.GetChunks().size(); }
int64_t tfo_list_chunks(tfo_map* h, int32_t* out, int64_t cap) {
  int64_t i = 0;
  for (const auto& kv : ((RefMap*)h)->chunkManager.GetChunks()) {
    if (i < cap)
      for (int k = 0; k < 3; k++) out[3 * i + k] = kv.first(k);
    i++;
  }
  return i;
}

int tfo_download_chunks(tfo_map* h, const int32_t* ids, int64_t n, float* sdf, float* weight, uint16_t* color) {
  RefMap* m = (RefMap*)h;
  for (int64_t i = 0; i < n; i++) {
    const ChunkID id(ids[3 * i], ids[3 * i + 1], ids[3 * i + 2]);
    if (!m->chunkManager.HasChunk(id)) return -4;
    const ChunkPtr c = m->chunkManager.GetChunk(id);
    if (sdf) memcpy(sdf + i * 512, c->voxels.sdf.data(), 2048);
    if (weight) memcpy(weight + i * 512, c->voxels.weight.data(), 2048);
    if (color) memcpy(color + i * 2048, c->colors.colorData, 4096);
  }
  return 0;
}

int tfo_get_observation(tfo_map* h, int32_t x, int32_t y, int32_t z, int keyframe, float* out) {
  RefMap* m = (RefMap*)h;
  const ChunkID id(x, y, z);
  if (!m->chunkManager.HasChunk(id)) return -4;
  const ChunkPtr c = m->chunkManager.GetChunk(id);
  auto it = c->observations.find(keyframe);
  if (it == c->observations.end()) return 0;
  *out = it->second;
  return 1;
}

int64_t tfo_meshes_to_update(tfo_map* h, int32_t* out, int64_t cap) {
  int64_t i = 0;
  for (const auto& kv : ((RefMap*)h)->meshesToUpdate) {
    if (!kv.second) continue;
    if (out && i < cap)
      for (int k = 0; k < 3; k++) out[3 * i + k] = kv.first(k);
    i++;
  }
  return i;
}

// Chisel::CompressMeshes ends with chunksToUpdate.clear() on meshesToUpdate (Structure/Chisel.cpp:146)
void tfo_clear_meshes_to_update(tfo_map* h) { ((RefMap*)h)->meshesToUpdate.clear(); }
// MobileFusion::RetractObservations (GCFusion/MobileFusion.cpp:252-272), the chunk half
void tfo_retract_observations(tfo_map* h, const int32_t* ids, int64_t n, int keyframe) {
  RefMap* m = (RefMap*)h;
  for (int64_t i = 0; i < n; i++) {
    const ChunkID id(ids[3 * i], ids[3 * i + 1], ids[3 * i + 2]);
    if (!m->chunkManager.HasChunk(id)) continue;
    m->chunkManager.GetChunk(id)->observations.erase(keyframe);
  }
}

float tfo_truncation_distance(const float* t5, float z) { return QuadraticTruncator(t5[0], t5[1], t5[2], t5[3]).GetTruncationDistance(z); }

void tfo_centroids(tfo_map* h, const float* pose, float* out3x512) {
  RefMap* m = (RefMap*)h;
  buffer_centroids(m, make_pose(pose));
  for (int v = 0; v < 512; v++)
    for (int k = 0; k < 3; k++) out3x512[k * 512 + v] = m->integrator.diff_centroids[v](k);
}

// ChunkManager::GenerateMeshEfficient (Structure/ChunkManager.cpp:595-1002) for a list of chunks,
// the per-chunk body of RecomputeMeshes (:232-264).  Chunks that are not in the map yield empty
// meshes.  Two-call pattern: with vert == nullptr only the offsets are filled.
//   vert_off / idx_off [n+1]: prefix sums of vertices / indices per chunk
//   vert, norm, col: xyz float triples; idx: int32 (local to the chunk's vertex block)
int tfo_mesh_chunks(tfo_map* h, const int32_t* ids, int64_t n, int64_t* vert_off, int64_t* idx_off, float* vert,
                    float* norm, float* col, int32_t* idx, int64_t vert_cap, int64_t idx_cap) {
  RefMap* m = (RefMap*)h;
  int64_t nv = 0, ni = 0;
  vert_off[0] = idx_off[0] = 0;
  for (int64_t i = 0; i < n; i++) {
    const ChunkID id(ids[3 * i], ids[3 * i + 1], ids[3 * i + 2]);
    Mesh mesh;
    if (m->chunkManager.HasChunk(id)) {
      mesh.chunkID = id;
      m->chunkManager.GenerateMeshEfficient(m->chunkManager.GetChunk(id), &mesh);
    }
    if (vert) {
      if (nv + (int64_t)mesh.vertices.size() > vert_cap || ni + (int64_t)mesh.indices.size() > idx_cap) return -3;
      for (size_t v = 0; v < mesh.vertices.size(); v++)
        for (int k = 0; k < 3; k++) {
          vert[3 * (nv + v) + k] = mesh.vertices[v](k);
          norm[3 * (nv + v) + k] = mesh.normals[v](k);
          col[3 * (nv + v) + k] = mesh.colors[v](k);
        }
      for (size_t t = 0; t < mesh.indices.size(); t++) idx[ni + t] = (int32_t)mesh.indices[t];
    }
    nv += (int64_t)mesh.vertices.size();
    ni += (int64_t)mesh.indices.size();
    vert_off[i + 1] = nv;
    idx_off[i + 1] = ni;
  }
  return 0;
}

}  // extern "C"
