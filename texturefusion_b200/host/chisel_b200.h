// chisel_b200.h — C++ drop-in shim: the reference's fusion classes over libtexfusion_b200.so.
//
// Keeps the names, argument lists and error behaviour of the classes that
// GCFusion/MobileFusion.{h,cpp} drives on the map thread:
//   chisel::Chisel                Structure/Chisel.h:46-493
//   chisel::ChunkManager          Structure/ChunkManager.h:119-207
//   chisel::ProjectionIntegrator  3rd_party/open_chisel/utils/ProjectionIntegrator.h:42-94
//   chisel::PinholeCamera         3rd_party/open_chisel/camera/PinholeCamera.h:33-75
//   chisel::Atlas / Patch         Structure/Atlas.h:43-75, Structure/Patch.h:51-94
// so that MobileFusion::ReIntegrateKeyframe / IntegrateFrame / tsdfFusion compile against it
// unchanged (INTEGRATION.md lists the include swap).  Everything voxel-sized happens on the GPU
// behind the C ABI (include/texfusion.h); this header only keeps what the reference keeps on the
// host: meshesToUpdate, per-chunk observations, validChunks lists, patch bookkeeping.
//
// Inside the reference tree define TF_WITH_EIGEN (Eigen types are used as they are); without it
// a minimal stand-in for Eigen::Vector3i / Affine3f is provided so this header builds alone.
#pragma once

#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "texfusion.h"

#ifdef TF_WITH_EIGEN
#include <Eigen/Core>
#include <Eigen/Geometry>
#endif

namespace chisel {

#ifdef TF_WITH_EIGEN
typedef Eigen::Vector3i ChunkID;
typedef Eigen::Vector3f Vec3;
typedef Eigen::Vector2f Vec2;
typedef Eigen::Affine3f Transform;
typedef std::vector<ChunkID, Eigen::aligned_allocator<ChunkID>> ChunkIDList;
inline const float* pose_data(const Transform& t) { return t.matrix().data(); }  // column-major 4x4
#else
struct ChunkID {
  int32_t v[3];
  ChunkID() : v{0, 0, 0} {}
  ChunkID(int x, int y, int z) : v{x, y, z} {}
  int32_t& operator()(int i) { return v[i]; }
  int32_t operator()(int i) const { return v[i]; }
  ChunkID operator+(const ChunkID& o) const { return ChunkID(v[0] + o.v[0], v[1] + o.v[1], v[2] + o.v[2]); }
  bool operator==(const ChunkID& o) const { return v[0] == o.v[0] && v[1] == o.v[1] && v[2] == o.v[2]; }
};
struct Vec2 { float x, y; };
struct Transform {  // Eigen::Affine3f layout: column-major 4x4, camera -> world
  float m[16];
  Transform() { std::memset(m, 0, sizeof(m)); m[0] = m[5] = m[10] = m[15] = 1.0f; }
};
typedef std::vector<ChunkID> ChunkIDList;
inline const float* pose_data(const Transform& t) { return t.m; }
#endif

// Structure/ChunkManager.h:44-53
struct ChunkHasher {
  std::size_t operator()(const ChunkID& k) const {
    return ((std::size_t)(int64_t)k(0) * (std::size_t)73856093) ^ ((std::size_t)(int64_t)k(1) * (std::size_t)19349663) ^
           ((std::size_t)(int64_t)k(2) * (std::size_t)83492791);
  }
};
typedef std::unordered_map<ChunkID, bool, ChunkHasher> ChunkSet;

inline tf_chunk_id to_c(const ChunkID& id) { return tf_chunk_id{id(0), id(1), id(2)}; }
inline tf_pose to_c(const Transform& t) {
  tf_pose p;
  std::memcpy(p.m, pose_data(t), sizeof(p.m));
  return p;
}

class TexFusionError : public std::runtime_error {
 public:
  TexFusionError(int code, const std::string& what) : std::runtime_error(what), code(code) {}
  int code;
};

// 3rd_party/open_chisel/camera/PinholeCamera.h:33-75.  The getters return int, as in the reference.
class PinholeCamera {
 public:
  void SetIntrinsics(float ifx, float ify, float icx, float icy) { fx = ifx; fy = ify; cx = icx; cy = icy; }
  int GetWidth() const { return width; }
  int GetHeight() const { return height; }
  int GetFx() const { return fx; }
  int GetFy() const { return fy; }
  int GetCx() const { return cx; }
  int GetCy() const { return cy; }
  void SetWidth(int v) { width = v; }
  void SetHeight(int v) { height = v; }
  float GetNearPlane() const { return nearPlane; }
  float GetFarPlane() const { return farPlane; }
  void SetNearPlane(float v) { nearPlane = v; }
  void SetFarPlane(float v) { farPlane = v; }
  tf_camera c_camera() const { return tf_camera{fx, fy, cx, cy, width, height, nearPlane, farPlane}; }

 protected:
  float fx = 525, fy = 525, cx = 319.5f, cy = 239.5f;
  int width = 640, height = 480;
  float nearPlane = 0.01f, farPlane = 5.0f;
};

// QuadraticTruncator / ConstantWeighter parameter carriers (GCFusion/MobileFusion.h:215-228)
struct QuadraticTruncator {
  QuadraticTruncator(float q = 0.0019f, float l = 0.00152f, float c = 0.001504f, float s = 6.0f)
      : quadraticTerm(q), linearTerm(l), constantTerm(c), scalingFactor(s) {}
  float quadraticTerm, linearTerm, constantTerm, scalingFactor;
};
typedef std::shared_ptr<QuadraticTruncator> TruncatorPtr;
struct ConstantWeighter {
  explicit ConstantWeighter(float w = 1.0f) : weight(w) {}
  float weight;
};
typedef std::shared_ptr<ConstantWeighter> WeighterPtr;

// ProjectionIntegrator: on the GPU the centroid buffers live in shared memory, so only the
// configuration survives (SetTruncator / SetWeighter / carving flags, GCFusion/MobileFusion.h:243-251).
class ProjectionIntegrator {
 public:
  const TruncatorPtr& GetTruncator() const { return truncator; }
  void SetTruncator(const TruncatorPtr& v) { truncator = v; }
  const WeighterPtr& GetWeighter() const { return weighter; }
  void SetWeighter(const WeighterPtr& v) { weighter = v; }
  void SetCarvingDist(float d) { carvingDist = d; }
  void SetCarvingEnabled(bool e) { enableVoxelCarving = e; }
  template <class L> void SetCentroids(const L&) {}
  tf_truncation c_truncation() const {
    QuadraticTruncator t = truncator ? *truncator : QuadraticTruncator();
    return tf_truncation{t.quadraticTerm, t.linearTerm, t.constantTerm, t.scalingFactor, weighter ? weighter->weight : 1.0f};
  }

 protected:
  TruncatorPtr truncator;
  WeighterPtr weighter;
  float carvingDist = 0;
  bool enableVoxelCarving = false;
};

// What callers touch of chisel::Chunk: ID, observations, and (for the CPU mesher,
// Structure/ChunkManager.cpp:614-626) the voxel planes, fetched on demand.
struct Chunk {
  ChunkID ID;
  std::map<int, float> observations;
  struct { std::vector<float> sdf, weight; } voxels;
  struct { std::vector<uint16_t> colorData; } colors;
  const ChunkID& GetID() const { return ID; }
};
typedef std::shared_ptr<Chunk> ChunkPtr;

class ChunkManager {
 public:
  ChunkManager() = default;
  void attach(tf_map* m, float res) { map = m; voxelResolutionMeters = res; }
  float GetResolution() const { return voxelResolutionMeters; }
  bool HasChunk(const ChunkID& id) const {
    int rc = tf_has_chunk(map, to_c(id));
    if (rc < 0) throw TexFusionError(rc, tf_last_error(map));
    return rc == 1;
  }
  // Structure/ChunkManager.h:137-139: unordered_map::at -> std::out_of_range for unknown ids
  ChunkPtr GetChunk(const ChunkID& id) {
    if (!HasChunk(id)) throw std::out_of_range("ChunkManager::GetChunk");
    ChunkPtr& c = host[id];
    if (!c) { c = std::make_shared<Chunk>(); c->ID = id; }
    return c;
  }
  bool RemoveChunk(const ChunkID& id) {
    if (!HasChunk(id)) return false;
    tf_chunk_id c = to_c(id);
    check(tf_remove_chunks(map, &c, 1));
    host.erase(id);
    return true;
  }
  int64_t GetChunkCount() const { return tf_chunk_count(map); }
  // chunk->voxels.sdf / weight, chunk->colors.colorData for a set of chunks (device -> host)
  void SyncToHost(const ChunkIDList& ids) {
    const size_t n = ids.size();
    if (!n) return;
    std::vector<tf_chunk_id> cid(n);
    std::vector<float> sdf(n * 512), w(n * 512);
    std::vector<uint16_t> col(n * 2048);
    for (size_t i = 0; i < n; i++) cid[i] = to_c(ids[i]);
    check(tf_download_chunks(map, cid.data(), (int64_t)n, sdf.data(), w.data(), col.data()));
    for (size_t i = 0; i < n; i++) {
      ChunkPtr c = GetChunk(ids[i]);
      c->voxels.sdf.assign(sdf.begin() + i * 512, sdf.begin() + (i + 1) * 512);
      c->voxels.weight.assign(w.begin() + i * 512, w.begin() + (i + 1) * 512);
      c->colors.colorData.assign(col.begin() + i * 2048, col.begin() + (i + 1) * 2048);
    }
  }
  void check(int rc) const { if (rc < 0) throw TexFusionError(rc, tf_last_error(map)); }
  void Reset() { host.clear(); }
  void forget(const ChunkID& id) { host.erase(id); }

 private:
  tf_map* map = nullptr;
  float voxelResolutionMeters = 0;
  std::unordered_map<ChunkID, ChunkPtr, ChunkHasher> host;
};

// Structure/Patch.h:51-94 — the fields the fusion path reads and writes.
struct Patch {
  int frameid = -1;
  std::size_t texloc = 0;
  bool has_image = false, wrong_mapping = false;
  int box[4] = {0, 0, 0, 0};  // boundingbox x, y, width, height (cv::Rect)
  void SetFrameid(int f) { frameid = f; }
  void SetImage(int x, int y, int w, int h) { box[0] = x; box[1] = y; box[2] = w; box[3] = h; has_image = true; }
  void clear() { frameid = -1; has_image = false; }
  bool complete() const { return has_image && frameid >= 0; }
};
typedef std::shared_ptr<Patch> PatchPtr;

// Structure/Atlas.{h,cpp}: the 13824 x 13824 x 3 texture lives in HBM; rows are fetched with
// DownloadRows for the GL upload (GCFusion/MobileFusion.h:404-427).
class Atlas {
 public:
  static const std::size_t MAX_PATCH_WIDTH = 96 * 72 * 2, MAX_PATCH_HEIGHT = 72 * 96 * 2;
  std::size_t PATCH_WIDTH = 0, PATCH_HEIGHT = 0, hot_start = 0, hot_end = 0;
  void attach(tf_map* m) {
    map = m;
    int32_t w, h;
    tf_atlas_patch_size(map, &w, &h);
    PATCH_WIDTH = (std::size_t)w;
    PATCH_HEIGHT = (std::size_t)h;
  }
  bool HasPatch(const ChunkID& id) const { return patches.count(id) != 0; }
  PatchPtr GetPatch(const ChunkID& id) { return patches.at(id); }
  // Atlas::AddPatch (Structure/Atlas.cpp:43-64); throws std::overflow_error when the atlas is full
  PatchPtr AddPatch(const ChunkID& id) {
    auto it = patches.find(id);
    if (it != patches.end()) { it->second->clear(); return it->second; }
    uint64_t loc = 0;
    int rc = tf_atlas_alloc_slot(map, to_c(id), &loc);
    if (rc == TF_ERR_ATLAS_FULL) throw std::overflow_error("No enough space for texture storage.");
    if (rc < 0) throw TexFusionError(rc, tf_last_error(map));
    PatchPtr p = std::make_shared<Patch>();
    p->texloc = (std::size_t)loc;
    patches[id] = p;
    return p;
  }
  Vec2 GetTexLoc(const ChunkID& id) {
    std::size_t k = GetPatch(id)->texloc;
    Vec2 r;
#ifdef TF_WITH_EIGEN
    r = Vec2((float)(k % MAX_PATCH_WIDTH), (float)(k / MAX_PATCH_WIDTH));
#else
    r.x = (float)(k % MAX_PATCH_WIDTH);
    r.y = (float)(k / MAX_PATCH_WIDTH);
#endif
    return r;
  }
  // Atlas::UpdateBuffer for a list of chunks in one launch (Chisel::UpdateAtlas, Structure/Chisel.cpp:191-196)
  void UpdateBuffers(const ChunkIDList& ids) {
    std::vector<tf_patch_desc> d;
    for (const ChunkID& id : ids) {
      auto it = patches.find(id);
      if (it == patches.end() || !it->second->complete()) continue;
      const Patch& p = *it->second;
      d.push_back(tf_patch_desc{(uint64_t)p.texloc, p.frameid, p.box[0], p.box[1], p.box[2], p.box[3]});
    }
    int rc = tf_atlas_update(map, d.data(), (int64_t)d.size());
    if (rc < 0) throw TexFusionError(rc, tf_last_error(map));
  }
  void UpdateBuffer(const ChunkID& id) { ChunkIDList l(1, id); UpdateBuffers(l); }
  void DownloadRows(std::size_t start, std::size_t end, uint8_t* rgb) {
    int rc = tf_atlas_download(map, start, end, rgb);
    if (rc < 0) throw TexFusionError(rc, tf_last_error(map));
  }

 private:
  tf_map* map = nullptr;
  std::unordered_map<ChunkID, PatchPtr, ChunkHasher> patches;
};

class Chisel {
 public:
  // Structure/Chisel.cpp:38-41.  The truncation parameters are those MobileFusion::initChiselMap
  // hands to the integrator (GCFusion/MobileFusion.h:215-228); pass the integrator to override.
  Chisel(const ChunkID& chunkSize, float voxelResolution, bool useColor, int width = 640, int height = 480,
         const ProjectionIntegrator* integrator = nullptr, int device = 0, int64_t max_chunks = 0, int max_frames = 0) {
    if (!(chunkSize(0) == 8 && chunkSize(1) == 8 && chunkSize(2) == 8))
      throw std::invalid_argument("only 8x8x8 chunks (GCFusion/MobileFusion.h:231-233)");
    tf_config cfg;
    std::memset(&cfg, 0, sizeof(cfg));
    cfg.chunk_dim = 8;
    cfg.voxel_res = voxelResolution;
    cfg.use_color = useColor;
    cfg.trunc = integrator ? integrator->c_truncation() : ProjectionIntegrator().c_truncation();
    cfg.device = device;
    cfg.n_ranks = 1;
    cfg.max_chunks = max_chunks;
    cfg.max_frames = max_frames;
    cfg.width = width;
    cfg.height = height;
    int rc = tf_create(&map, &cfg);
    if (rc < 0) throw TexFusionError(rc, tf_last_error(nullptr));
    npix = (size_t)width * height;
    chunkManager.attach(map, voxelResolution);
    atlas.attach(map);
  }
  virtual ~Chisel() { tf_destroy(map); }
  Chisel(const Chisel&) = delete;
  Chisel& operator=(const Chisel&) = delete;

  const ChunkManager& GetChunkManager() const { return chunkManager; }
  ChunkManager& GetMutableChunkManager() { return chunkManager; }
  const ChunkSet& GetMeshesToUpdate() const { return meshesToUpdate; }
  void Reset() { check(tf_reset(map)); chunkManager.Reset(); meshesToUpdate.clear(); }

  // Structure/Chisel.h:103-140
  void PrepareIntersectChunks(ProjectionIntegrator&, float* depthImage, const Transform& depthExtrinsic,
                              const PinholeCamera& depthCamera, ChunkIDList& chunksIntersecting,
                              std::vector<bool>& needsUpdateFlag, std::vector<bool>& newChunkFlag) {
    chunksIntersecting.clear();
    needsUpdateFlag.clear();
    newChunkFlag.clear();
    check(tf_upload_frame(map, kScratchFrame, depthImage, nullptr, nullptr));
    const tf_pose pose = to_c(depthExtrinsic);
    const tf_camera cam = depthCamera.c_camera();
    int64_t n = 0;
    if (ids_buf.size() < 4096) { ids_buf.resize(4096); flag_buf.resize(4096); }
    int rc = tf_prepare(map, kScratchFrame, &pose, &cam, ids_buf.data(), flag_buf.data(), (int64_t)ids_buf.size(), &n);
    if (rc == TF_ERR_CAPACITY && n > (int64_t)ids_buf.size()) {  // two-call pattern
      ids_buf.resize((size_t)n);
      flag_buf.resize((size_t)n);
      rc = tf_prepare(map, kScratchFrame, &pose, &cam, ids_buf.data(), flag_buf.data(), n, &n);
    }
    check(rc);
    for (int64_t i = 0; i < n; i++) {
      chunksIntersecting.push_back(ChunkID(ids_buf[i].x, ids_buf[i].y, ids_buf[i].z));
      newChunkFlag.push_back(flag_buf[i] != 0);
      needsUpdateFlag.push_back(false);
    }
  }

  // Structure/Chisel.h:218-249
  void IntegrateDepthScanColor(ProjectionIntegrator&, float* depthImage, unsigned char* colorImage,
                               const Transform& depthExtrinsic, const PinholeCamera& depthCamera,
                               ChunkIDList& chunksIntersecting, std::vector<bool>& needsUpdateFlag, int integrate_flag,
                               int keyframeID = -1, float* observationQualityPointer = nullptr) {
    const size_t n = chunksIntersecting.size();
    if (n < 1) return;
    check(tf_upload_frame(map, kScratchFrame, depthImage, colorImage, colorImage ? observationQualityPointer : nullptr));
    const tf_pose pose = to_c(depthExtrinsic);
    const tf_camera cam = depthCamera.c_camera();
    if (ids_buf.size() < n) { ids_buf.resize(n); flag_buf.resize(n); }
    q_buf.resize(n);
    for (size_t i = 0; i < n; i++) {
      ids_buf[i] = to_c(chunksIntersecting[i]);
      flag_buf[i] = needsUpdateFlag[i] ? 1 : 0;
    }
    int rc = tf_integrate(map, kScratchFrame, colorImage != nullptr, &pose, &cam, ids_buf.data(), (int64_t)n, integrate_flag,
                          flag_buf.data(), q_buf.data());
    if (rc == TF_ERR_NOT_FOUND) throw std::out_of_range("ChunkManager::GetChunk");
    check(rc);
    for (size_t i = 0; i < n; i++) {
      needsUpdateFlag[i] = flag_buf[i] != 0;
      if (keyframeID >= 0 && q_buf[i] > 0 && needsUpdateFlag[i])  // Structure/Chisel.h:244-247
        chunkManager.GetChunk(chunksIntersecting[i])->observations[keyframeID] = q_buf[i];
    }
  }

  // Structure/Chisel.h:453-468: Prepare + Integrate(1) + Finalize, fused on the device
  void IntegrateDepthScanColor(ProjectionIntegrator&, float* depthImage, unsigned char* colorImage,
                               const Transform& depthExtrinsic, const PinholeCamera& depthCamera) {
    check(tf_upload_frame(map, kScratchFrame, depthImage, colorImage, nullptr));
    const tf_pose pose = to_c(depthExtrinsic);
    const tf_camera cam = depthCamera.c_camera();
    if (ids_buf.size() < (size_t)kFusedCap) { ids_buf.resize(kFusedCap); flag_buf.resize(kFusedCap); }
    upd_buf.resize(kFusedCap);
    tf_frame_stats st;
    check(tf_integrate_frame(map, kScratchFrame, colorImage != nullptr, &pose, &cam, &st, ids_buf.data(), nullptr,
                             upd_buf.data(), nullptr, kFusedCap));
    last_stats = st;
    for (int64_t i = 0; i < st.n_chunks && i < kFusedCap; i++)
      if (upd_buf[i]) MarkMeshes(ChunkID(ids_buf[i].x, ids_buf[i].y, ids_buf[i].z));
  }

  // Structure/Chisel.h:184-216
  void FinalizeIntegrateChunks(ChunkIDList& chunksIntersecting, std::vector<bool>& needsUpdateFlag,
                               std::vector<bool>& newChunkFlag, ChunkIDList& validChunks) {
    validChunks.clear();
    ChunkIDList garbageChunks;
    for (size_t i = 0; i < chunksIntersecting.size(); i++) {
      const ChunkID& id = chunksIntersecting[i];
      if (needsUpdateFlag[i]) {
        MarkMeshes(id);
        validChunks.push_back(id);
      } else if (newChunkFlag[i]) {
        garbageChunks.push_back(id);
      }
    }
    GarbageCollect(garbageChunks);
  }

  // Structure/Chisel.h:472-477
  void GarbageCollect(const ChunkIDList& chunks) {
    if (chunks.empty()) return;
    std::vector<tf_chunk_id> c(chunks.size());
    for (size_t i = 0; i < chunks.size(); i++) c[i] = to_c(chunks[i]);
    check(tf_remove_chunks(map, c.data(), (int64_t)c.size()));
    for (const ChunkID& id : chunks) {
      chunkManager.forget(id);
      meshesToUpdate.erase(id);
    }
  }

  // Key-frame colour for the atlas: Frame::rgb (+ colorValidFlag) kept in HBM under frame_index.
  void UploadKeyframe(int frame_index, const float* depth, const unsigned char* rgb, const unsigned char* colorValid,
                      const float* quality) {
    check(tf_upload_frame(map, frame_index, depth, nullptr, quality));
    check(tf_upload_keyframe_rgb(map, frame_index, rgb, colorValid));
  }
  // Patch::CalculateTexCoords (Structure/Patch.cpp:40-108) for all patches of one key-frame in one
  // launch.  T_g_l = frame.pose_sophus[0].inverse().matrix().cast<float>() (column-major 16 floats);
  // mesh p owns vertices [offsets[p], offsets[p+1]).  Fills texcoord (2/vertex), texcolor (3/vertex)
  // and per patch {boundingbox x, y, w, h, wrong_mapping, flag}.
  void CalculateTexCoords(int frame_index, const float* T_g_l, const PinholeCamera& camera, int64_t n_patches,
                          const int64_t* offsets, const float* vertices, const float* colors, float* texcoord,
                          float* texcolor, tf_patch_result* results) {
    tf_pose T;
    std::memcpy(T.m, T_g_l, sizeof(T.m));
    const tf_camera cam = camera.c_camera();
    check(tf_patch_texcoords(map, frame_index, &T, &cam, n_patches, offsets, vertices, colors, texcoord, texcolor, results));
  }
  // Structure/Chisel.cpp:191-196
  void UpdateAtlas(ChunkIDList& chunksToUpdate) { atlas.UpdateBuffers(chunksToUpdate); }

  tf_map* handle() { return map; }
  tf_frame_stats last_stats{};

  ChunkManager chunkManager;
  ChunkSet meshesToUpdate;
  Atlas atlas;

 protected:
  void check(int rc) const { if (rc < 0) throw TexFusionError(rc, tf_last_error(map)); }
  void MarkMeshes(const ChunkID& id) {  // Structure/Chisel.h:197-203
    meshesToUpdate[id] = true;
    meshesToUpdate[id + ChunkID(-1, 0, 0)] = true;
    meshesToUpdate[id + ChunkID(1, 0, 0)] = true;
    meshesToUpdate[id + ChunkID(0, -1, 0)] = true;
    meshesToUpdate[id + ChunkID(0, 1, 0)] = true;
    meshesToUpdate[id + ChunkID(0, 0, -1)] = true;
    meshesToUpdate[id + ChunkID(0, 0, 1)] = true;
  }
  static const int32_t kScratchFrame = 0x7F000000;  // frame-store slot for images passed by pointer
  static const int64_t kFusedCap = 1 << 17;
  tf_map* map = nullptr;
  size_t npix = 0;
  std::vector<tf_chunk_id> ids_buf;
  std::vector<uint8_t> flag_buf, upd_buf;
  std::vector<float> q_buf;
};

}  // namespace chisel
