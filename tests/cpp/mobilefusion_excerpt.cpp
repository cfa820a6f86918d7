// mobilefusion_excerpt.cpp — the map-thread code path of GCFusion/MobileFusion.cpp written against
// the shim (texturefusion_b200/host/chisel_b200.h) exactly as the reference writes it against its own
// classes: ReIntegrateKeyframe (:114-221), IntegrateFrame (:223-250), RetractObservations (:252-272) and
// the tail of tsdfFusion (:327-385: UpdateMeshes, CompressMeshes, GeneratePatches, CompensateColor,
// UpdateAtlas, DrawMeshes).  The types those functions take from the rest of the reference (Frame,
// KeyFrameDatabase, Sophus::SE3d, cv::Mat, UniGraph) are stubbed with the members the calls touch.
// Compiled with -DTF_WITH_EIGEN against the Eigen stand-in of oracle/eigen_standin (tests only).
//
//   mobilefusion_excerpt <input.bin> <voxel_res> <output.bin> [--bench N] [--export DIR]
//
// Runs: K key-frame groups through ReIntegrateKeyframe(…, 1); the tsdfFusion tail; a loop closure of
// key-frame 0 (Retract + ReIntegrate 0 + ReIntegrate 1 under a corrected pose); the tail again; dumps
// everything observable to <output.bin> for tests/test_shim_gpu.py to compare with the CPU oracle.
// --bench: times N repetitions of one key-frame group through the shim and through the raw C ABI.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <unordered_map>
#include <vector>

#include "chisel_b200.h"

using chisel::ChunkID;
using chisel::ChunkIDList;

// ---- stand-ins for the reference types the calls touch ------------------------------------------
namespace cvstub {
struct Mat {  // cv::Mat: .data, .empty(), .at<T>(i)
  std::vector<unsigned char> buf;
  unsigned char* data = nullptr;
  void assign(const void* p, size_t n) { buf.assign((const unsigned char*)p, (const unsigned char*)p + n); data = buf.data(); }
  bool empty() const { return buf.empty(); }
  template <class T> T& at(int i) { return ((T*)data)[i]; }
  template <class T> const T& at(int i) const { return ((const T*)data)[i]; }
};
}  // namespace cvstub
namespace Sophus {
struct SE3d {  // matrix(), inverse()
  Eigen::Matrix4d M = Eigen::Matrix4d::Identity();
  Eigen::Matrix4d matrix() const { return M; }
  SE3d inverse() const {  // (R^T, -R^T t)
    SE3d r;
    for (int i = 0; i < 3; i++) {
      for (int j = 0; j < 3; j++) r.M(i, j) = M(j, i);
      r.M(i, 3) = -(M(0, i) * M(0, 3) + M(1, i) * M(1, 3) + M(2, i) * M(2, 3));
    }
    return r;
  }
};
}  // namespace Sophus
struct Frame {  // GCSLAM/frame.h:29-161, the members the fusion path reads
  cvstub::Mat refined_depth, rgb, colorValidFlag, observationQualityMap;
  Sophus::SE3d pose_sophus[2];
  int frame_index = 0;
  ChunkIDList validChunks;
  std::vector<void*> validChunksPtr;
  bool tracking_success = true, is_keyframe = false;
  int origin_index = 0;
};
namespace MultiViewGeometry {
struct KeyFrameDatabase { int keyFrameIndex; std::vector<int> corresponding_frames; };
}
struct UniGraph {  // Structure/uni_graph.h: chunk -> node, node -> label (key-frame index)
  std::unordered_map<ChunkID, std::size_t, chisel::ChunkHasher, chisel::ChunkEq> chunks;
  std::vector<std::size_t> labels;
  std::size_t get_label(std::size_t n) const { return labels[n]; }
};

// Chisel::CompensateColor stays the reference's own host code; nothing to do here
void chisel::Chisel::CompensateColor() {}

#define INTEGRATE_ALL 1

struct MobileFusion {
  chisel::ChiselPtr chiselMap;
  chisel::ProjectionIntegrator projectionIntegrator;
  chisel::PinholeCamera cameraModel;
  UniGraph chunkGraph;
  ChunkIDList chunksToUpdate;
  std::vector<float> tsdf_vertices_buffer;
  std::vector<unsigned int> tsdf_indices_buffer;
  unsigned int tsdf_indice_num = 0, tsdf_vertice_num = 0;

  // GCFusion/MobileFusion.cpp:114-221
  void ReIntegrateKeyframe(std::vector<Frame>& frame_list, const MultiViewGeometry::KeyFrameDatabase& kfDatabase,
                           const int integrateFlag) {
    Frame& kf = frame_list[kfDatabase.keyFrameIndex];
    int totalPixelNum = cameraModel.GetWidth() * cameraModel.GetHeight();
    int width = cameraModel.GetWidth();
    int height = cameraModel.GetHeight();
    chisel::Transform lastPose;
    ChunkIDList localChunksIntersecting;
    std::vector<bool> localNeedsUpdateFlag;
    std::vector<bool> localNewChunkFlag;
    if (integrateFlag == 1) {
      lastPose = kf.pose_sophus[0].matrix().cast<float>();
      kf.pose_sophus[1] = kf.pose_sophus[0];
    } else if (integrateFlag == 0) {
      lastPose = kf.pose_sophus[1].matrix().cast<float>();
      localChunksIntersecting = kf.validChunks;
      for (size_t i = 0; i < localChunksIntersecting.size(); i++) {
        localNeedsUpdateFlag.emplace_back(true);
        localNewChunkFlag.emplace_back(false);
      }
    }
    float* depthImageData;
    static unsigned char* colorImageData = new unsigned char[totalPixelNum * 4];
    unsigned char* colorValid = (unsigned char*)kf.colorValidFlag.data;
    depthImageData = (float*)kf.refined_depth.data;
    float* observationQualityPointer = (float*)kf.observationQualityMap.data;
    for (int i = 0; i < height; i++) {
      for (int j = 0; j < width; j++) {
        int pos = i * width + j;
        colorImageData[pos * 4 + 0] = colorValid[pos] > 0 ? kf.rgb.at<unsigned char>(pos * 3 + 0) : 0;
        colorImageData[pos * 4 + 1] = colorValid[pos] > 0 ? kf.rgb.at<unsigned char>(pos * 3 + 1) : 0;
        colorImageData[pos * 4 + 2] = colorValid[pos] > 0 ? kf.rgb.at<unsigned char>(pos * 3 + 2) : 0;
        colorImageData[pos * 4 + 3] = colorValid[pos] > 0 ? 1 : 0;
      }
    }
    if (integrateFlag == 1) {
      chiselMap->PrepareIntersectChunks(projectionIntegrator, depthImageData, lastPose, cameraModel, localChunksIntersecting,
                                        localNeedsUpdateFlag, localNewChunkFlag);
    }
    chiselMap->IntegrateDepthScanColor(projectionIntegrator, depthImageData, colorImageData, lastPose, cameraModel,
                                       localChunksIntersecting, localNeedsUpdateFlag, integrateFlag, kf.frame_index,
                                       observationQualityPointer);
#if INTEGRATE_ALL
    for (size_t i = 0; i < kfDatabase.corresponding_frames.size(); i++) {
      Frame& local_frame = frame_list[kfDatabase.corresponding_frames[i]];
      if (local_frame.refined_depth.empty()) continue;
      if (integrateFlag == 1) {
        lastPose = local_frame.pose_sophus[0].matrix().cast<float>();
        local_frame.pose_sophus[1] = local_frame.pose_sophus[0];
      } else if (integrateFlag == 0) {
        lastPose = local_frame.pose_sophus[1].matrix().cast<float>();
      }
      depthImageData = (float*)local_frame.refined_depth.data;
      chiselMap->IntegrateDepthScanColor(projectionIntegrator, depthImageData, NULL, lastPose, cameraModel,
                                         localChunksIntersecting, localNeedsUpdateFlag, integrateFlag);
    }
#endif
    if (integrateFlag == 1) {
      chiselMap->FinalizeIntegrateChunks(localChunksIntersecting, localNeedsUpdateFlag, localNewChunkFlag, kf.validChunks);
    } else if (integrateFlag == 0) {
      ChunkIDList localValidChunks;
      chiselMap->FinalizeIntegrateChunks(localChunksIntersecting, localNeedsUpdateFlag, localNewChunkFlag, localValidChunks);
      kf.validChunks.clear();
    }
  }

  // GCFusion/MobileFusion.cpp:223-250
  void IntegrateFrame(const Frame& frame_ref) {
    int totalPixelNum = cameraModel.GetWidth() * cameraModel.GetHeight();
    chisel::Transform lastPose;
    lastPose = frame_ref.pose_sophus[0].matrix().cast<float>();
    if (frame_ref.refined_depth.empty()) return;
    float* depthImageData = (float*)frame_ref.refined_depth.data;
    unsigned char* colorImageData;
    if (frame_ref.rgb.empty()) {
      colorImageData = NULL;
    } else {
      colorImageData = new unsigned char[totalPixelNum * 4];
      for (int j = 0; j < totalPixelNum; j++) {
        colorImageData[j * 4 + 0] = frame_ref.rgb.at<unsigned char>(j * 3 + 0);
        colorImageData[j * 4 + 1] = frame_ref.rgb.at<unsigned char>(j * 3 + 1);
        colorImageData[j * 4 + 2] = frame_ref.rgb.at<unsigned char>(j * 3 + 2);
        colorImageData[j * 4 + 3] = 1;
      }
    }
    if (frame_ref.tracking_success && frame_ref.origin_index == 0)
      chiselMap->IntegrateDepthScanColor(projectionIntegrator, depthImageData, colorImageData, lastPose, cameraModel);
    delete[] colorImageData;
  }

  // GCFusion/MobileFusion.cpp:252-272 (the data-cost bookkeeping of TexMap is not part of this excerpt)
  void RetractObservations(chisel::ChunkManager& manager, Frame& kf) {
    int frame_id = kf.frame_index;
    for (size_t i = 0; i < kf.validChunks.size(); i++) {
      if (!manager.HasChunk(kf.validChunks[i])) continue;
      chisel::ChunkPtr chunk = manager.GetChunk(kf.validChunks[i]);
      chunk->observations.erase(frame_id);
    }
  }

  // the tail of tsdfFusion (GCFusion/MobileFusion.cpp:327-385); the MRF is replaced by "the key-frame
  // with the largest observation quality" so that labels are deterministic
  int MeshAndTexture(std::vector<Frame>& frame_list) {
    chiselMap->UpdateMeshes(cameraModel);
    chunksToUpdate.clear();
    const chisel::MeshMap& allMeshes = chiselMap->chunkManager.GetAllMeshes();
    for (auto it : chiselMap->meshesToUpdate) {
      if (!it.second) continue;
      if (allMeshes.find(it.first) == allMeshes.end()) continue;
      chunksToUpdate.emplace_back(it.first);
    }
    chiselMap->CompressMeshes(chiselMap->meshesToUpdate);
    ChunkIDList textured;
    for (const ChunkID& id : chunksToUpdate) {  // stand-in for update_datacost + view_selection
      if (!chiselMap->chunkManager.HasChunk(id)) continue;
      const auto& obs = chiselMap->chunkManager.GetChunk(id)->observations;
      if (obs.empty()) continue;
      int best = -1;
      float bq = -1;
      for (const auto& o : obs)
        if (o.second > bq) { bq = o.second; best = o.first; }
      auto ins = chunkGraph.chunks.emplace(id, chunkGraph.labels.size());
      if (ins.second) chunkGraph.labels.push_back((std::size_t)best);
      else chunkGraph.labels[ins.first->second] = (std::size_t)best;
      textured.push_back(id);
    }
    chunksToUpdate = textured;
    int eot = chiselMap->GeneratePatches(chunksToUpdate, chunkGraph, frame_list, cameraModel);
    if (eot < 0) return -1;
    chiselMap->CompensateColor();
    chiselMap->UpdateAtlas(chunksToUpdate);
    chiselMap->DrawMeshes(tsdf_vertices_buffer.data(), tsdf_indices_buffer.data(), tsdf_indice_num, tsdf_vertice_num);
    return 1;
  }
};

// ---- driver ---------------------------------------------------------------------------------------
static void put(FILE* f, const void* p, size_t n) { fwrite(p, 1, n, f); }
static void put_i64(FILE* f, int64_t v) { put(f, &v, 8); }

static bool id_less(const ChunkID& a, const ChunkID& b) {
  if (a(0) != b(0)) return a(0) < b(0);
  if (a(1) != b(1)) return a(1) < b(1);
  return a(2) < b(2);
}

static void dump_state(FILE* out, MobileFusion& mf, std::vector<Frame>& frames, const std::vector<int>& kf_indices) {
  chisel::Chisel& C = *mf.chiselMap;
  // valid lists of the key-frames
  put_i64(out, (int64_t)kf_indices.size());
  for (int k : kf_indices) {
    put_i64(out, (int64_t)frames[k].validChunks.size());
    for (const ChunkID& id : frames[k].validChunks) { int32_t v[3] = {id(0), id(1), id(2)}; put(out, v, 12); }
  }
  // every chunk of the map: id, voxel planes, observations (sorted by id)
  ChunkIDList ids;
  for (const auto& kv : C.chunkManager.GetChunks()) ids.push_back(kv.first);
  std::sort(ids.begin(), ids.end(), id_less);
  C.chunkManager.SyncToHost(ids);
  put_i64(out, (int64_t)ids.size());
  put_i64(out, (int64_t)C.chunkManager.GetChunkCount());
  for (const ChunkID& id : ids) {
    chisel::ChunkPtr c = C.chunkManager.GetChunk(id);
    int32_t v[3] = {id(0), id(1), id(2)};
    put(out, v, 12);
    put(out, c->voxels.sdf.data(), 2048);
    put(out, c->voxels.weight.data(), 2048);
    put(out, c->colors.colorData.data(), 4096);
    put_i64(out, (int64_t)c->observations.size());
    for (const auto& o : c->observations) { int32_t k = o.first; put(out, &k, 4); put(out, &o.second, 4); }
  }
  // meshes + patches (sorted by id)
  ChunkIDList mids;
  for (const auto& kv : C.chunkManager.GetAllMeshes()) mids.push_back(kv.first);
  std::sort(mids.begin(), mids.end(), id_less);
  put_i64(out, (int64_t)mids.size());
  for (const ChunkID& id : mids) {
    const chisel::Mesh& m = *C.chunkManager.GetMesh(id);
    int32_t v[3] = {id(0), id(1), id(2)};
    put(out, v, 12);
    put_i64(out, (int64_t)m.vertices.size());
    put_i64(out, (int64_t)m.indices.size());
    for (size_t j = 0; j < m.vertices.size(); j++) {
      float r[9] = {m.vertices[j](0), m.vertices[j](1), m.vertices[j](2), m.normals[j](0), m.normals[j](1), m.normals[j](2),
                    m.colors[j](0),   m.colors[j](1),   m.colors[j](2)};
      put(out, r, 36);
    }
    for (size_t j = 0; j < m.indices.size(); j++) { int32_t i = (int32_t)m.indices[j]; put(out, &i, 4); }
    int32_t adj = 0;
    for (int k = 0; k < 6; k++) adj |= m.adj[k] ? (1 << k) : 0;
    put(out, &adj, 4);
    const chisel::PatchPtr p = m.m_patch;
    int32_t has = (p != nullptr && p->complete()) ? 1 : 0;
    put(out, &has, 4);
    if (has) {
      int64_t texloc = (int64_t)p->texloc;
      int32_t pr[6] = {p->frameid, p->boundingbox.x, p->boundingbox.y, p->boundingbox.width, p->boundingbox.height, p->wrong_mapping ? 1 : 0};
      put(out, &texloc, 8);
      put(out, pr, 24);
      put_i64(out, (int64_t)p->texcoord.size());
      for (size_t j = 0; j < p->texcoord.size(); j++) {
        float r[5] = {p->texcoord[j](0), p->texcoord[j](1), p->texcolor[j](0), p->texcolor[j](1), p->texcolor[j](2)};
        put(out, r, 20);
      }
    }
  }
  // atlas: hot range + its rows from the host mirror; GL buffers
  put_i64(out, (int64_t)C.atlas.hot_start);
  put_i64(out, (int64_t)C.atlas.hot_end);
  if (C.atlas.hot_end > C.atlas.hot_start) put(out, C.atlas.texture_buffer.data + C.atlas.hot_start * 3, (C.atlas.hot_end - C.atlas.hot_start) * 3);
  put_i64(out, (int64_t)mf.tsdf_vertice_num);
  put_i64(out, (int64_t)mf.tsdf_indice_num);
}

int main(int argc, char** argv) {
  if (argc < 4) return 2;
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 2;
  const float res = (float)atof(argv[2]);
  int bench = 0;
  std::string export_dir;
  for (int i = 4; i + 1 < argc; i++) {
    if (std::string(argv[i]) == "--bench") bench = atoi(argv[i + 1]);
    if (std::string(argv[i]) == "--export") export_dir = argv[i + 1];
  }
  int32_t hdr[4];  // W, H, n_frames, group size
  if (fread(hdr, 4, 4, f) != 4) return 2;
  const int W = hdr[0], H = hdr[1], nfr = hdr[2], group = hdr[3];
  const size_t npix = (size_t)W * H;
  float camf[6];
  if (fread(camf, 4, 6, f) != 6) return 2;
  MobileFusion mf;
  // MobileFusion::initChiselMap (GCFusion/MobileFusion.h:205-258)
  mf.cameraModel.SetIntrinsics(camf[0], camf[1], camf[2], camf[3]);
  mf.cameraModel.SetNearPlane(camf[4]);
  mf.cameraModel.SetFarPlane(camf[5]);
  mf.cameraModel.SetWidth(W);
  mf.cameraModel.SetHeight(H);
  try {
    mf.chiselMap = chisel::ChiselPtr(new chisel::Chisel(Eigen::Vector3i(8, 8, 8), res, true, W, H));
    mf.projectionIntegrator.SetCentroids(mf.chiselMap->GetChunkManager().GetCentroids());
    mf.projectionIntegrator.SetTruncator(chisel::TruncatorPtr(new chisel::QuadraticTruncator(0.0019f, 0.00152f, 0.001504f, 6.0f)));
    mf.projectionIntegrator.SetWeighter(chisel::WeighterPtr(new chisel::ConstantWeighter(1)));
    mf.projectionIntegrator.SetCarvingDist(0.05f);
    mf.projectionIntegrator.SetCarvingEnabled(true);
    mf.tsdf_vertices_buffer.resize(12 * 4000000);
    mf.tsdf_indices_buffer.resize(12000000);

    std::vector<Frame> frames(nfr);
    std::vector<MultiViewGeometry::KeyFrameDatabase> kflist;
    std::vector<int> kf_indices;
    std::vector<float> tmp(npix);
    std::vector<Sophus::SE3d> corrected(nfr), drifted(nfr);
    for (int k = 0; k < nfr; k++) {
      Frame& fr = frames[k];
      float pose[2][16];
      if (fread(pose, 4, 32, f) != 32) return 2;  // corrected pose, drifted pose (column-major)
      for (int c = 0; c < 4; c++)
        for (int r = 0; r < 4; r++) {
          corrected[k].M(r, c) = pose[0][c * 4 + r];
          drifted[k].M(r, c) = pose[1][c * 4 + r];
        }
      fr.pose_sophus[0] = fr.pose_sophus[1] = drifted[k];  // what tracking knows before the loop closure
      fr.frame_index = k;
      fr.is_keyframe = (k % group) == 0;
      if (fread(tmp.data(), 4, npix, f) != npix) return 2;
      fr.refined_depth.assign(tmp.data(), npix * 4);
      if (fr.is_keyframe) {
        std::vector<unsigned char> rgb(npix * 3), valid(npix);
        if (fread(rgb.data(), 1, npix * 3, f) != npix * 3) return 2;
        if (fread(valid.data(), 1, npix, f) != npix) return 2;
        if (fread(tmp.data(), 4, npix, f) != npix) return 2;
        fr.rgb.assign(rgb.data(), npix * 3);
        fr.colorValidFlag.assign(valid.data(), npix);
        fr.observationQualityMap.assign(tmp.data(), npix * 4);
        kflist.push_back({k, {}});
        kf_indices.push_back(k);
      } else {
        kflist.back().corresponding_frames.push_back(k);
      }
    }
    fclose(f);

    if (bench > 0) {  // one key-frame group through the shim vs the same calls on the raw C ABI
      using clk = std::chrono::steady_clock;
      mf.ReIntegrateKeyframe(frames, kflist[0], 1);  // the first fusion, under the drifted poses
      double t_shim = 0, t_raw = 0;
      tf_map* m = mf.chiselMap->handle();
      const tf_camera cam = mf.cameraModel.c_camera();
      std::vector<unsigned char> rgba(npix * 4);
      Frame& kf = frames[kflist[0].keyFrameIndex];
      chisel::ChunkSet raw_meshes;  // the raw arm keeps the caller-side state of the protocol too
      ChunkIDList raw_valid;
      std::vector<tf_chunk_id> ids(1 << 18);
      std::vector<uint8_t> isnew(1 << 18), nu(1 << 18);
      std::vector<float> q(1 << 18);
      for (int rep = 0; rep < bench; rep++) {
        auto t0 = clk::now();
        mf.ReIntegrateKeyframe(frames, kflist[0], 0);
        mf.ReIntegrateKeyframe(frames, kflist[0], 1);
        auto t1 = clk::now();
        t_shim += std::chrono::duration<double, std::micro>(t1 - t0).count();
        // the same protocol with raw calls: de-integrate over validChunks, prepare, integrate, remove
        const int64_t nv = (int64_t)kf.validChunks.size();
        for (int64_t i = 0; i < nv; i++) { ids[i] = chisel::to_c(kf.validChunks[i]); nu[i] = 1; }
        auto t2 = clk::now();
        auto run = [&](int flag, int64_t n) {
          for (size_t p = 0; p < npix; p++) {  // the RGBA pack of ReIntegrateKeyframe (:151-162) is caller work in both arms
            const bool ok = kf.colorValidFlag.data[p] > 0;
            for (int c = 0; c < 3; c++) rgba[p * 4 + c] = ok ? kf.rgb.data[p * 3 + c] : 0;
            rgba[p * 4 + 3] = ok ? 1 : 0;
          }
          chisel::Transform T;
          T = kf.pose_sophus[flag ? 0 : 1].matrix().cast<float>();
          tf_pose pose = chisel::to_c(T);
          tf_upload_frame(m, 0x7F000000, (float*)kf.refined_depth.data, rgba.data(), (float*)kf.observationQualityMap.data);
          if (flag) {
            tf_prepare(m, 0x7F000000, &pose, &cam, ids.data(), isnew.data(), (int64_t)ids.size(), &n);
            for (int64_t i = 0; i < n; i++) nu[i] = 0;
          }
          tf_integrate(m, 0x7F000000, 1, &pose, &cam, ids.data(), n, flag, nu.data(), q.data());
          for (int lf : kflist[0].corresponding_frames) {
            T = frames[lf].pose_sophus[flag ? 0 : 1].matrix().cast<float>();
            pose = chisel::to_c(T);
            tf_upload_frame(m, 0x7F000000, (float*)frames[lf].refined_depth.data, nullptr, nullptr);
            tf_integrate(m, 0x7F000000, 0, &pose, &cam, ids.data(), n, flag, nu.data(), nullptr);
          }
          std::vector<tf_chunk_id> garbage;  // FinalizeIntegrateChunks (Structure/Chisel.h:184-216)
          raw_valid.clear();
          for (int64_t i = 0; i < n; i++) {
            if (nu[i]) {
              const ChunkID id = chisel::from_c(ids[i]);
              raw_meshes[id] = true;
              for (int k = 0; k < 6; k++) raw_meshes[id + chisel::neighbourhood[k]] = true;
              raw_valid.push_back(id);
            } else if (flag && isnew[i]) {
              garbage.push_back(ids[i]);
            }
          }
          if (!garbage.empty()) tf_remove_chunks(m, garbage.data(), (int64_t)garbage.size());
          for (const tf_chunk_id& g : garbage) raw_meshes.erase(chisel::from_c(g));
          return n;
        };
        run(0, nv);
        run(1, 0);
        auto t3 = clk::now();
        t_raw += std::chrono::duration<double, std::micro>(t3 - t2).count();
      }
      printf("bench: shim %.1f us / key-frame (de- + re-integration of a %d-frame group), raw C ABI %.1f us, ratio %.3f\n",
             t_shim / bench, group, t_raw / bench, t_shim / t_raw);
      return 0;
    }

    FILE* out = fopen(argv[3], "wb");
    if (!out) return 2;
    // 1. first fusion under the drifted poses
    for (const auto& kfdb : kflist) mf.ReIntegrateKeyframe(frames, kfdb, 1);
    if (mf.MeshAndTexture(frames) < 0) return 3;
    dump_state(out, mf, frames, kf_indices);
    // 2. loop closure of every key-frame: the corrected poses arrive in pose_sophus[0]; pose_sophus[1] is
    //    what the map holds (set by ReIntegrateKeyframe(…, 1), GCFusion/MobileFusion.cpp:133-134)
    for (int k = 0; k < nfr; k++) frames[k].pose_sophus[0] = corrected[k];
    for (const auto& kfdb : kflist) {
      mf.RetractObservations(mf.chiselMap->chunkManager, frames[kfdb.keyFrameIndex]);
      mf.ReIntegrateKeyframe(frames, kfdb, 0);
      mf.ReIntegrateKeyframe(frames, kfdb, 1);
    }
    if (mf.MeshAndTexture(frames) < 0) return 3;
    dump_state(out, mf, frames, kf_indices);
    if (!export_dir.empty()) {  // main.cpp:263-270: PLY of the meshes, OBJ + MTL + image of the textured model
      if (!mf.chiselMap->SaveAllMeshesToPLY(export_dir + "/model.ply")) return 4;
      const std::size_t rows = mf.chiselMap->atlas.hot_end / chisel::Atlas::MAX_PATCH_WIDTH;
      mf.chiselMap->atlas.SaveTexturedModel(export_dir, rows);
    }
    // 3. a plain frame through the fused convenience form (IntegrateFrame)
    mf.IntegrateFrame(frames[1]);
    put_i64(out, mf.chiselMap->last_stats.n_chunks);
    put_i64(out, mf.chiselMap->last_stats.n_updated);
    put_i64(out, mf.chiselMap->chunkManager.GetChunkCount());
    put_i64(out, (int64_t)mf.chiselMap->chunkManager.GetChunks().size());
    // error behaviour: GetChunk on an unknown id throws std::out_of_range like unordered_map::at
    int32_t thrown = 0;
    try { mf.chiselMap->chunkManager.GetChunk(ChunkID(100000, 0, 0)); } catch (const std::out_of_range&) { thrown = 1; }
    put(out, &thrown, 4);
    fclose(out);
    return 0;
  } catch (const std::exception& e) {
    fprintf(stderr, "mobilefusion_excerpt: %s\n", e.what());
    return 1;
  }
}
