// chisel_b200.h — C++ drop-in shim: the reference's fusion classes over libtexfusion_b200.so.
//
// Keeps the names, argument lists and error behaviour of the classes that
// GCFusion/MobileFusion.{h,cpp} drives on the map thread:
//   chisel::Chisel                Structure/Chisel.h:46-493, Structure/Chisel.cpp
//   chisel::ChunkManager          Structure/ChunkManager.h:119-207, 714-735
//   chisel::ProjectionIntegrator  3rd_party/open_chisel/utils/ProjectionIntegrator.h:42-94
//   chisel::PinholeCamera         3rd_party/open_chisel/camera/PinholeCamera.h:33-75
//   chisel::Mesh                  3rd_party/open_chisel/geometry/Mesh.h:37-84
//   chisel::Atlas / Patch         Structure/Atlas.h:41-75, Structure/Patch.h:51-94
// so that MobileFusion::ReIntegrateKeyframe / IntegrateFrame / RetractObservations / tsdfFusion
// compile against it unchanged (INTEGRATION.md lists the include swap; tests/cpp/mobilefusion_excerpt.cpp
// is that code path compiled and run against this header).  Everything voxel-, pixel- or
// vertex-sized happens on the GPU behind the C ABI (include/texfusion.h); this header keeps what the
// reference keeps on the host: the id set of the map with per-chunk observations, meshesToUpdate,
// validChunks lists, the meshes (filled by the device mesher) and the patch bookkeeping.
// No call in here goes to the device per chunk: ids are tracked from the results of the batched
// calls (tf_prepare / tf_integrate_frame / tf_remove_chunks).
//
// Inside the reference tree define TF_WITH_EIGEN (Eigen types are used as they are) and, for
// Patch::SetImage(cv::Mat&) / Atlas::texture_buffer as cv::Mat, TF_WITH_OPENCV; without them
// minimal stand-ins are provided so this header builds alone.
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <fstream>
#include <iomanip>
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "texfusion.h"

#ifdef TF_WITH_EIGEN
#include <Eigen/Core>
#include <Eigen/Geometry>
#endif
#ifdef TF_WITH_OPENCV
#include <opencv2/core.hpp>
#include <opencv2/imgcodecs.hpp>
#include <opencv2/imgproc.hpp>
#endif

namespace chisel {

#ifdef TF_WITH_EIGEN
typedef Eigen::Vector3i ChunkID;
typedef Eigen::Vector3i Point3;
typedef Eigen::Vector3f Vec3;
typedef Eigen::Vector2f Vec2;
typedef Eigen::Affine3f Transform;
typedef std::vector<ChunkID, Eigen::aligned_allocator<ChunkID>> ChunkIDList;
typedef std::vector<Vec3, Eigen::aligned_allocator<Vec3>> Vec3List;
typedef std::vector<Vec2, Eigen::aligned_allocator<Vec2>> Vec2List;
#else
template <class T, int N>
struct VecN {
  T v[N];
  VecN() { for (int i = 0; i < N; i++) v[i] = T(); }
  VecN(T a, T b) { static_assert(N == 2, ""); v[0] = a; v[1] = b; }
  VecN(T a, T b, T c) { static_assert(N == 3, ""); v[0] = a; v[1] = b; v[2] = c; }
  T& operator()(int i) { return v[i]; }
  const T& operator()(int i) const { return v[i]; }
  T& operator[](int i) { return v[i]; }
  const T& operator[](int i) const { return v[i]; }
  VecN operator+(const VecN& o) const { VecN r; for (int i = 0; i < N; i++) r.v[i] = v[i] + o.v[i]; return r; }
  VecN operator-(const VecN& o) const { VecN r; for (int i = 0; i < N; i++) r.v[i] = v[i] - o.v[i]; return r; }
  VecN& operator+=(const VecN& o) { for (int i = 0; i < N; i++) v[i] += o.v[i]; return *this; }
  bool operator==(const VecN& o) const { for (int i = 0; i < N; i++) if (!(v[i] == o.v[i])) return false; return true; }
};
typedef VecN<int32_t, 3> ChunkID;
typedef VecN<int32_t, 3> Point3;
typedef VecN<float, 3> Vec3;
typedef VecN<float, 2> Vec2;
struct Transform {  // Eigen::Affine3f layout: column-major 4x4, camera -> world
  float m[16];
  Transform() { std::memset(m, 0, sizeof(m)); m[0] = m[5] = m[10] = m[15] = 1.0f; }
};
typedef std::vector<ChunkID> ChunkIDList;
typedef std::vector<Vec3> Vec3List;
typedef std::vector<Vec2> Vec2List;
#endif

// Structure/ChunkManager.h:44-53
struct ChunkHasher {
  std::size_t operator()(const ChunkID& k) const {
    return ((std::size_t)(int64_t)k(0) * (std::size_t)73856093) ^ ((std::size_t)(int64_t)k(1) * (std::size_t)19349663) ^
           ((std::size_t)(int64_t)k(2) * (std::size_t)83492791);
  }
};
struct ChunkEq {
  bool operator()(const ChunkID& a, const ChunkID& b) const { return a(0) == b(0) && a(1) == b(1) && a(2) == b(2); }
};
typedef std::unordered_map<ChunkID, bool, ChunkHasher, ChunkEq> ChunkSet;
const ChunkID neighbourhood[6] = {ChunkID(-1, 0, 0), ChunkID(1, 0, 0), ChunkID(0, -1, 0),
                                  ChunkID(0, 1, 0),  ChunkID(0, 0, -1), ChunkID(0, 0, 1)};

inline tf_chunk_id to_c(const ChunkID& id) { return tf_chunk_id{id(0), id(1), id(2)}; }
inline ChunkID from_c(const tf_chunk_id& id) { return ChunkID(id.x, id.y, id.z); }
inline tf_pose to_c(const Transform& t) {
  tf_pose p;
#ifdef TF_WITH_EIGEN
  const Eigen::Matrix4f mat = t.matrix();
  std::memcpy(p.m, mat.data(), sizeof(p.m));
#else
  std::memcpy(p.m, t.m, sizeof(p.m));
#endif
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
typedef QuadraticTruncator Truncator;
typedef std::shared_ptr<QuadraticTruncator> TruncatorPtr;
struct ConstantWeighter {
  explicit ConstantWeighter(float w = 1.0f) : weight(w) {}
  float weight;
};
typedef ConstantWeighter Weighter;
typedef std::shared_ptr<ConstantWeighter> WeighterPtr;

// ProjectionIntegrator: on the GPU the centroid buffers live in shared memory, so only the
// configuration survives (SetTruncator / SetWeighter / carving flags, GCFusion/MobileFusion.h:243-251).
// The reference reads truncator and weighter on every voxelUpdateSIMD call
// (ProjectionIntegrator.cpp:90-92); Chisel forwards them to the library on every call that takes
// the integrator (tf_set_truncation when they changed).
class ProjectionIntegrator {
 public:
  const TruncatorPtr& GetTruncator() const { return truncator; }
  void SetTruncator(const TruncatorPtr& v) { truncator = v; }
  const WeighterPtr& GetWeighter() const { return weighter; }
  void SetWeighter(const WeighterPtr& v) { weighter = v; }
  void SetCarvingDist(float d) { carvingDist = d; }
  void SetCarvingEnabled(bool e) { enableVoxelCarving = e; }
  float GetCarvingDist() const { return carvingDist; }
  bool IsCarvingEnabled() const { return enableVoxelCarving; }
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

struct Patch;
typedef std::shared_ptr<Patch> PatchPtr;

// 3rd_party/open_chisel/geometry/Mesh.h:37-84 (GRID_EACH_DIM 8)
typedef std::size_t VertIndex;
typedef std::vector<VertIndex> VertIndexList;
struct Mesh {
  Vec3List vertices, normals, colors;
  VertIndexList indices;
  Point3 chunkID;
  bool simplified = false;
  bool adj[6] = {false, false, false, false, false, false};
  bool outlier_checked = false;
  Vec3 origin;
  float grid_resolution = 0;
  PatchPtr m_patch;
  bool HasVertices() const { return !vertices.empty(); }
  void Clear() {  // Mesh.h:49-63
    vertices.clear(); normals.clear(); colors.clear(); indices.clear();
    std::memset(adj, 0, sizeof(adj));
    simplified = false;
  }
  // Mesh.cpp:39-82: which faces of the chunk the mesh touches (8-cell grid per axis)
  int GetIndice(const Vec3& vert) {
    int pos[3];
    for (int j = 0; j < 3; j++) pos[j] = (int)std::floor((vert(j) - origin(j)) / grid_resolution);
    if (pos[0] >= 8) adj[1] = true;
    if (pos[1] >= 8) adj[3] = true;
    if (pos[2] >= 8) adj[5] = true;
    if (pos[0] <= 0) adj[0] = true;
    if (pos[1] <= 0) adj[2] = true;
    if (pos[2] <= 0) adj[4] = true;
    return pos[0] + pos[1] * 8 + pos[2] * 64;
  }
  void SimplifyByClustering(float resolution, const Vec3& chunkOri) {
    if (simplified) return;
    origin = chunkOri;
    grid_resolution = resolution;
    for (std::size_t i = 0; i < vertices.size(); i++) GetIndice(vertices[i]);
    simplified = true;
  }
};
typedef std::shared_ptr<Mesh> MeshPtr;
typedef std::unordered_map<ChunkID, MeshPtr, ChunkHasher, ChunkEq> MeshMap;

// What callers touch of chisel::Chunk: ID, origin, observations, and (for code that still reads
// voxels on the host) the voxel planes, fetched on demand by ChunkManager::SyncToHost.
struct Chunk {
  ChunkID ID;
  Vec3 origin;
  std::map<int, float> observations;
  struct { std::vector<float> sdf, weight; } voxels;
  struct { std::vector<uint16_t> colorData; } colors;
  const ChunkID& GetID() const { return ID; }
  const Vec3& GetOrigin() const { return origin; }
};
typedef std::shared_ptr<Chunk> ChunkPtr;
typedef std::unordered_map<ChunkID, ChunkPtr, ChunkHasher, ChunkEq> ChunkMap;

class ChunkManager {
 public:
  ChunkManager() = default;
  void attach(tf_map* m, float res) {
    map = m;
    voxelResolutionMeters = res;
    chunkSize = ChunkID(8, 8, 8);
    centroids.clear();  // ChunkManager::CacheCentroids (Structure/ChunkManager.cpp:49-62)
    const float half = res * 0.5f;
    for (int z = 0; z < 8; z++)
      for (int y = 0; y < 8; y++)
        for (int x = 0; x < 8; x++) centroids.push_back(Vec3((float)x * res + half, (float)y * res + half, (float)z * res + half));
  }
  float GetResolution() const { return voxelResolutionMeters; }
  const ChunkID& GetChunkSize() const { return chunkSize; }
  const Vec3List& GetCentroids() const { return centroids; }
  const ChunkMap& GetChunks() const { return host; }

  // The id set of the map is mirrored on the host (note_exists / forget, driven by the results of
  // the batched device calls): no device round trip per id.
  bool HasChunk(const ChunkID& id) const { return host.find(id) != host.end(); }
  bool HasChunk(int x, int y, int z) const { return HasChunk(ChunkID(x, y, z)); }
  // Structure/ChunkManager.h:137-139: unordered_map::at -> std::out_of_range for unknown ids
  ChunkPtr GetChunk(const ChunkID& id) const {
    auto it = host.find(id);
    if (it == host.end()) throw std::out_of_range("ChunkManager::GetChunk");
    return it->second;
  }
  bool RemoveChunk(const ChunkID& id) {  // :151-161
    if (!HasChunk(id)) return false;
    tf_chunk_id c = to_c(id);
    check(tf_remove_chunks(map, &c, 1));
    forget(id);
    return true;
  }
  int64_t GetChunkCount() const { return tf_chunk_count(map); }

  // meshes (Structure/ChunkManager.h:714-725)
  const MeshMap& GetAllMeshes() const { return allMeshes; }
  MeshMap& GetAllMutableMeshes() { return allMeshes; }
  bool HasMesh(const ChunkID& id) const { return allMeshes.find(id) != allMeshes.end(); }
  const MeshPtr& GetMesh(const ChunkID& id) const { return allMeshes.at(id); }
  MeshPtr& GetMutableMesh(const ChunkID& id) { return allMeshes.at(id); }

  // ChunkManager::RecomputeMeshes (Structure/ChunkManager.cpp:232-264): marching cubes for every
  // flagged chunk that is in the map — one tf_mesh_chunks call, on the voxels where they lie in HBM.
  void RecomputeMeshes(const ChunkSet& chunkMeshes, const PinholeCamera&) {
    if (chunkMeshes.empty()) return;
    std::vector<tf_chunk_id> ids;
    std::vector<MeshPtr> meshes;
    for (const auto& kv : chunkMeshes) {
      if (!kv.second || !HasChunk(kv.first)) continue;
      MeshPtr mesh = HasMesh(kv.first) ? allMeshes[kv.first] : std::make_shared<Mesh>();
      mesh->chunkID = kv.first;
      ids.push_back(to_c(kv.first));
      meshes.push_back(mesh);
    }
    const int64_t n = (int64_t)ids.size();
    if (!n) return;
    mesh_voff.resize(n + 1);
    mesh_ioff.resize(n + 1);
    if (mesh_v.size() < 3 * 4096) { mesh_v.resize(3 * 4096); mesh_n.resize(3 * 4096); mesh_c.resize(3 * 4096); mesh_i.resize(3 * 4096); }
    int rc = tf_mesh_chunks(map, ids.data(), n, mesh_voff.data(), mesh_ioff.data(), mesh_v.data(), mesh_n.data(), mesh_c.data(),
                            mesh_i.data(), (int64_t)mesh_v.size() / 3, (int64_t)mesh_i.size());
    if (rc == TF_ERR_CAPACITY) {  // the offsets hold the required sizes: grow (with slack) and repeat
      mesh_v.resize((size_t)(mesh_voff[n] * 3 * 3 / 2)); mesh_n.resize(mesh_v.size()); mesh_c.resize(mesh_v.size());
      mesh_i.resize((size_t)(mesh_ioff[n] * 3 / 2));
      rc = tf_mesh_chunks(map, ids.data(), n, mesh_voff.data(), mesh_ioff.data(), mesh_v.data(), mesh_n.data(), mesh_c.data(),
                          mesh_i.data(), (int64_t)mesh_v.size() / 3, (int64_t)mesh_i.size());
    }
    check(rc);
    for (int64_t k = 0; k < n; k++) {
      Mesh& mesh = *meshes[k];
      mesh.Clear();
      const int64_t v0 = mesh_voff[k], v1 = mesh_voff[k + 1], i0 = mesh_ioff[k], i1 = mesh_ioff[k + 1];
      mesh.vertices.reserve(v1 - v0); mesh.normals.reserve(v1 - v0); mesh.colors.reserve(v1 - v0); mesh.indices.reserve(i1 - i0);
      for (int64_t v = v0; v < v1; v++) {
        mesh.vertices.push_back(Vec3(mesh_v[3 * v], mesh_v[3 * v + 1], mesh_v[3 * v + 2]));
        mesh.normals.push_back(Vec3(mesh_n[3 * v], mesh_n[3 * v + 1], mesh_n[3 * v + 2]));
        mesh.colors.push_back(Vec3(mesh_c[3 * v], mesh_c[3 * v + 1], mesh_c[3 * v + 2]));
      }
      for (int64_t i = i0; i < i1; i++) mesh.indices.push_back((VertIndex)mesh_i[i]);
      if (!mesh.vertices.empty()) allMeshes[from_c(ids[k])] = meshes[k];  // :261-263
    }
  }

  // chunk->voxels.sdf / weight, chunk->colors.colorData for a set of chunks (device -> host), for
  // host code that still reads voxels (the reference's own mesher no longer needs it)
  void SyncToHost(const ChunkIDList& ids) {
    const size_t n = ids.size();
    if (!n) return;
    std::vector<tf_chunk_id> cid(n);
    std::vector<float> sdf(n * 512), w(n * 512);
    std::vector<uint16_t> col(n * 2048);
    for (size_t i = 0; i < n; i++) cid[i] = to_c(ids[i]);
    check(tf_download_chunks(map, cid.data(), (int64_t)n, sdf.data(), w.data(), col.data()));
    for (size_t i = 0; i < n; i++) {
      ChunkPtr c = note_exists(ids[i]);  // (tf_download_chunks has just vouched for the id)
      c->voxels.sdf.assign(sdf.begin() + i * 512, sdf.begin() + (i + 1) * 512);
      c->voxels.weight.assign(w.begin() + i * 512, w.begin() + (i + 1) * 512);
      c->colors.colorData.assign(col.begin() + i * 2048, col.begin() + (i + 1) * 2048);
    }
  }
  void check(int rc) const { if (rc < 0) throw TexFusionError(rc, tf_last_error(map)); }
  void Reset() { host.clear(); allMeshes.clear(); }  // Structure/ChunkManager.cpp:272-275

  // host mirror of the id set
  ChunkPtr note_exists(const ChunkID& id) {
    ChunkPtr& c = host[id];
    if (!c) {
      c = std::make_shared<Chunk>();
      c->ID = id;
      c->origin = Vec3((float)(8 * id(0)) * voxelResolutionMeters, (float)(8 * id(1)) * voxelResolutionMeters,
                       (float)(8 * id(2)) * voxelResolutionMeters);  // Chunk.cpp:52
    }
    return c;
  }
  void forget(const ChunkID& id) {  // RemoveChunk also drops the mesh (:155-157)
    host.erase(id);
    allMeshes.erase(id);
  }

 private:
  tf_map* map = nullptr;
  float voxelResolutionMeters = 0;
  ChunkID chunkSize;
  Vec3List centroids;
  ChunkMap host;
  MeshMap allMeshes;
  std::vector<int64_t> mesh_voff, mesh_ioff;
  std::vector<float> mesh_v, mesh_n, mesh_c;
  std::vector<int32_t> mesh_i;
};

// cv::Rect stand-in
#ifdef TF_WITH_OPENCV
typedef cv::Rect Box;
#else
struct Box {
  int x = 0, y = 0, width = 0, height = 0;
  Box() = default;
  Box(int x_, int y_, int w_, int h_) : x(x_), y(y_), width(w_), height(h_) {}
};
#endif

// Structure/Patch.h:51-94 — the fields the fusion path, DrawMeshes and CompensateColor read and write.
struct Patch {
  explicit Patch(MeshPtr meshit = nullptr) : mesh(meshit) { clear(); texloc = 0; }
  int frameid = -1;
  MeshPtr mesh;
  std::size_t texloc = 0;
#ifdef TF_WITH_OPENCV
  cv::Mat image;  // view into the key-frame's rgb, like the reference; the atlas copy itself runs on the device
#endif
  bool has_image = false, has_adjusted = false, has_updated = false, wrong_mapping = false;
  Box boundingbox;
  Vec2List texcoord;
  Vec3List texcolor;
  Vec2 ratio;
  Vec3List paras, labs, labt;
  std::vector<size_t> caution;
  void SetMesh(MeshPtr m) { mesh = m; }
  void SetFrameid(int f) { frameid = f; }
  ChunkID GetChunkid() const { return mesh->chunkID; }
  float GetWidth() const { return (float)boundingbox.width; }
  float GetHeight() const { return (float)boundingbox.height; }
  int GetCoordsNum() const { return (int)texcoord.size(); }
#ifdef TF_WITH_OPENCV
  void SetImage(cv::Mat& view) { image = view(boundingbox); has_image = true; }  // Structure/Patch.cpp:172-175
#endif
  template <class Image> void SetImage(Image&) { has_image = true; }  // (the crop is boundingbox of key-frame `frameid` in HBM)
  void clear() {  // Structure/Patch.cpp:177-189
    frameid = -1;
#ifdef TF_WITH_OPENCV
    image.release();
#endif
    has_image = false;
    has_adjusted = false;
    texcoord.clear(); texcolor.clear(); labs.clear(); paras.clear();
    ratio = Vec2(1, 1);
  }
  bool complete() const {  // :191-196
    if (mesh == nullptr || mesh->vertices.empty() || !mesh->simplified) return false;
    if (!has_image || texcoord.empty() || frameid < 0) return false;
    return true;
  }
};

// Structure/Atlas.{h,cpp}: the 13824 x 13824 x 3 texture lives in HBM.  texture_buffer is the host
// mirror the GL upload reads (GCFusion/MobileFusion.h:404-427); only the hot rows are kept current
// (SyncHotRange, called by Chisel::UpdateAtlas).
class Atlas {
 public:
  static const std::size_t MAX_PATCH_WIDTH = 96 * 72 * 2, MAX_PATCH_HEIGHT = 72 * 96 * 2;
  std::size_t loc_next = 0, PATCH_WIDTH = 0, PATCH_HEIGHT = 0, hot_start = 0, hot_end = 0;
#ifdef TF_WITH_OPENCV
  cv::Mat texture_buffer;
#else
  struct HostTexture { unsigned char* data = nullptr; int rows = 0, cols = 0; } texture_buffer;
#endif
  ~Atlas() { tf_host_free(host_pixels); }
  void attach(tf_map* m, ChunkManager* cm) {
    map = m;
    manager = cm;
    int32_t w, h;
    tf_atlas_patch_size(map, &w, &h);
    PATCH_WIDTH = (std::size_t)w;
    PATCH_HEIGHT = (std::size_t)h;
  }
  bool HasPatch(const ChunkID& id) const { return manager->HasMesh(id); }                     // Atlas.h:57
  PatchPtr GetPatch(const ChunkID& id) { return manager->GetMutableMesh(id)->m_patch; }       // :58-60
  // Atlas::AddPatch (Structure/Atlas.cpp:43-64); throws std::overflow_error when the atlas is full
  PatchPtr AddPatch(MeshPtr mesh) {
    if (mesh->m_patch == nullptr) {
      uint64_t loc = 0;
      const int rc = tf_atlas_alloc_slot(map, to_c(mesh->chunkID), &loc);
      if (rc == TF_ERR_ATLAS_FULL) throw std::overflow_error("No enough space for texture storage.");
      if (rc < 0) throw TexFusionError(rc, tf_last_error(map));
      mesh->m_patch = std::make_shared<Patch>(mesh);
      mesh->m_patch->texloc = (std::size_t)loc;
      loc_next = std::max<std::size_t>(loc_next, (std::size_t)loc);  // (the allocator state itself lives in the library)
    } else {
      mesh->m_patch->clear();
    }
    return mesh->m_patch;
  }
  Vec2 GetTexLoc(const ChunkID& id) {
    const std::size_t k = GetPatch(id)->texloc;
    return Vec2((float)(k % MAX_PATCH_WIDTH), (float)(k / MAX_PATCH_WIDTH));
  }
  // Atlas::UpdateBuffer (Structure/Atlas.cpp:71-91) for a list of chunks in ONE launch
  void UpdateBuffers(const ChunkIDList& ids) {
    descs.clear();
    for (const ChunkID& id : ids) {
      if (!HasPatch(id)) continue;
      PatchPtr p = GetPatch(id);
      if (p == nullptr || !p->complete()) continue;
      const Box& b = p->boundingbox;
      if ((std::size_t)b.width > PATCH_WIDTH) p->ratio(0) = float(PATCH_WIDTH) / b.width;     // :77-80
      if ((std::size_t)b.height > PATCH_HEIGHT) p->ratio(1) = float(PATCH_HEIGHT) / b.height;
      descs.push_back(tf_patch_desc{(uint64_t)p->texloc, p->frameid, b.x, b.y, b.width, b.height});
    }
    const int rc = tf_atlas_update(map, descs.data(), (int64_t)descs.size());
    if (rc < 0) throw TexFusionError(rc, tf_last_error(map));
  }
  void UpdateBuffer(const ChunkID& id) { ChunkIDList l(1, id); UpdateBuffers(l); }
  // rows [hot_start, hot_end) of the device atlas -> texture_buffer (page-locked, allocated on first use)
  void SyncHotRange() {
    if (hot_end <= hot_start) return;
    if (!host_pixels) {
      host_pixels = (unsigned char*)tf_host_alloc(MAX_PATCH_WIDTH * MAX_PATCH_HEIGHT * 3);
      if (!host_pixels) throw TexFusionError(TF_ERR_CUDA, "tf_host_alloc(atlas mirror) failed");
      std::memset(host_pixels, 0, MAX_PATCH_WIDTH * MAX_PATCH_HEIGHT * 3);
#ifdef TF_WITH_OPENCV
      texture_buffer = cv::Mat((int)MAX_PATCH_HEIGHT, (int)MAX_PATCH_WIDTH, CV_8UC3, host_pixels);
#else
      texture_buffer.data = host_pixels;
      texture_buffer.rows = (int)MAX_PATCH_HEIGHT;
      texture_buffer.cols = (int)MAX_PATCH_WIDTH;
#endif
    }
    const std::size_t end = std::min(hot_end, MAX_PATCH_WIDTH * MAX_PATCH_HEIGHT);
    const int rc = tf_atlas_download(map, hot_start, end, host_pixels + hot_start * 3);
    if (rc < 0) throw TexFusionError(rc, tf_last_error(map));
  }
  // The hot rows straight into a mapped CUDA-GL pixel-unpack buffer (device pointer from
  // cudaGraphicsResourceGetMappedPointer): replaces glBufferDataARB(&texture_buffer.data[hot_start*3]) of
  // GCFusion/MobileFusion.h:406-412 without the detour over the host.
  void CopyHotRangeToDevice(void* mapped_pbo) {
    if (hot_end <= hot_start) return;
    const std::size_t end = std::min(hot_end, MAX_PATCH_WIDTH * MAX_PATCH_HEIGHT);
    const int rc = tf_atlas_copy_to_device(map, hot_start, end, mapped_pbo);
    if (rc < 0) throw TexFusionError(rc, tf_last_error(map));
  }
  // Atlas::SaveTexturedModel (Structure/Atlas.cpp:93-179): texture_model.obj / .mtl + the atlas image.
  // With OpenCV the image is texture_material.png like the reference; without, a binary PPM
  // (texture_material.ppm, same pixels) — the .mtl names the file that was written.  `rows` limits the
  // image to its first rows (0 = all 13824; every slot in use lies above loc_next's row + PATCH_HEIGHT).
  void SaveTexturedModel(const std::string& basepath, std::size_t rows = 0) {
    if (rows == 0 || rows > MAX_PATCH_HEIGHT) rows = MAX_PATCH_HEIGHT;
    const std::size_t keep_start = hot_start, keep_end = hot_end;
    hot_start = 0;
    hot_end = rows * MAX_PATCH_WIDTH;
    SyncHotRange();  // the whole image (or its first rows) from HBM into the host mirror
    hot_start = keep_start;
    hot_end = keep_end;
#ifdef TF_WITH_OPENCV
    const std::string tex_name = "texture_material.png";
    cv::Mat texture_bgr;
    cv::cvtColor(texture_buffer(cv::Rect(0, 0, (int)MAX_PATCH_WIDTH, (int)rows)), texture_bgr, cv::COLOR_RGB2BGR);
    cv::imwrite(basepath + "/" + tex_name, texture_bgr);
#else
    const std::string tex_name = "texture_material.ppm";
    {
      std::ofstream img((basepath + "/" + tex_name).c_str(), std::ios::binary);
      img << "P6\n" << MAX_PATCH_WIDTH << " " << rows << "\n255\n";
      img.write((const char*)texture_buffer.data, (std::streamsize)(rows * MAX_PATCH_WIDTH * 3));
    }
#endif
    Vec3List vertices, normals;
    Vec2List texcoords;
    VertIndexList indices;
    std::size_t vts = 0;
    for (auto& it : manager->GetAllMeshes()) {
      MeshPtr mesh = it.second;
      PatchPtr patch = mesh->m_patch;
      if (patch == nullptr || !patch->complete() || patch->texcoord.size() != mesh->vertices.size()) continue;
      for (std::size_t j = 0; j < mesh->indices.size(); j++) indices.emplace_back(mesh->indices[j] + vts);
      const Vec2 loc = GetTexLoc(mesh->chunkID);
      for (std::size_t j = 0; j < mesh->vertices.size(); j++) {
        vertices.emplace_back(mesh->vertices[j]);
        normals.emplace_back(mesh->normals[j]);
        Vec2 tex = loc;
        tex(0) += patch->texcoord[j](0) * patch->ratio(0);
        tex(1) += patch->texcoord[j](1) * patch->ratio(1);
        tex(0) /= MAX_PATCH_WIDTH;
        tex(1) /= MAX_PATCH_HEIGHT;
        texcoords.emplace_back(tex);
        vts++;
      }
    }
    std::ofstream mout((basepath + "/texture_model.obj").c_str());
    mout << "mtllib texture_model.mtl" << '\n' << std::fixed << std::setprecision(6);
    for (std::size_t i = 0; i < vertices.size(); ++i) mout << "v " << vertices[i](0) << " " << vertices[i](1) << " " << vertices[i](2) << '\n';
    for (std::size_t i = 0; i < texcoords.size(); ++i) mout << "vt " << texcoords[i](0) << " " << 1.0f - texcoords[i](1) << '\n';
    for (std::size_t i = 0; i < normals.size(); ++i) mout << "vn " << normals[i](0) << " " << normals[i](1) << " " << normals[i](2) << '\n';
    mout << "s off" << '\n' << "usemtl demo_texture" << '\n';
    for (std::size_t i = 0; i < indices.size() / 3; ++i) {
      mout << "g face" << i << '\n' << "f";
      for (std::size_t k = 0; k < 3; ++k) mout << " " << indices[i * 3 + k] + 1 << "/" << indices[i * 3 + k] + 1 << "/" << indices[i * 3 + k] + 1;
      mout << '\n';
    }
    mout.close();
    std::ofstream out((basepath + "/texture_model.mtl").c_str());
    out << "newmtl demo_texture" << '\n' << "Ka 1.000000 1.000000 1.000000" << '\n' << "Kd 1.000000 1.000000 1.000000" << '\n'
        << "Ks 0.000000 0.000000 0.000000" << '\n' << "Tr 0.000000" << '\n' << "illum 1" << '\n' << "Ns 1.000000" << '\n'
        << "map_Kd " << tex_name << std::endl;
  }

  void DownloadRows(std::size_t start, std::size_t end, uint8_t* rgb) {
    const int rc = tf_atlas_download(map, start, end, rgb);
    if (rc < 0) throw TexFusionError(rc, tf_last_error(map));
  }

 private:
  tf_map* map = nullptr;
  ChunkManager* manager = nullptr;
  unsigned char* host_pixels = nullptr;
  std::vector<tf_patch_desc> descs;
};

class Chisel {
 public:
  // Structure/Chisel.cpp:38-41.  width / height / device / pool sizes are the only additions.
  Chisel(const ChunkID& chunkSize, float voxelResolution, bool useColor, int width = 640, int height = 480,
         const ProjectionIntegrator* integrator = nullptr, int device = 0, int64_t max_chunks = 0, int max_frames = 0,
         int dot3_order = 0) {
    if (!(chunkSize(0) == 8 && chunkSize(1) == 8 && chunkSize(2) == 8))
      throw std::invalid_argument("only 8x8x8 chunks (GCFusion/MobileFusion.h:231-233)");
    tf_config cfg;
    std::memset(&cfg, 0, sizeof(cfg));
    cfg.chunk_dim = 8;
    cfg.voxel_res = voxelResolution;
    cfg.use_color = useColor;
    trunc = integrator ? integrator->c_truncation() : ProjectionIntegrator().c_truncation();
    cfg.trunc = trunc;
    cfg.device = device;
    cfg.n_ranks = 1;
    cfg.max_chunks = max_chunks;
    cfg.max_frames = max_frames;
    cfg.width = width;
    cfg.height = height;
    cfg.dot3_order = dot3_order;
    int rc = tf_create(&map, &cfg);
    if (rc < 0) throw TexFusionError(rc, tf_last_error(nullptr));
    npix = (size_t)width * height;
    chunkManager.attach(map, voxelResolution);
    atlas.attach(map, &chunkManager);
  }
  virtual ~Chisel() { tf_destroy(map); }
  Chisel(const Chisel&) = delete;
  Chisel& operator=(const Chisel&) = delete;

  const ChunkManager& GetChunkManager() const { return chunkManager; }
  ChunkManager& GetMutableChunkManager() { return chunkManager; }
  const ChunkSet& GetMeshesToUpdate() const { return meshesToUpdate; }
  void Reset() { check(tf_reset(map)); chunkManager.Reset(); meshesToUpdate.clear(); }

  // Structure/Chisel.h:103-140
  void PrepareIntersectChunks(ProjectionIntegrator& integrator, float* depthImage, const Transform& depthExtrinsic,
                              const PinholeCamera& depthCamera, ChunkIDList& chunksIntersecting,
                              std::vector<bool>& needsUpdateFlag, std::vector<bool>& newChunkFlag) {
    sync_truncation(integrator);
    chunksIntersecting.clear();
    needsUpdateFlag.clear();
    newChunkFlag.clear();
    check(tf_upload_frame(map, kScratchFrame, depthImage, nullptr, nullptr));
    const tf_pose pose = to_c(depthExtrinsic);
    const tf_camera cam = depthCamera.c_camera();
    int64_t n = 0;
    if (ids_buf.size() < 16384) { ids_buf.resize(16384); flag_buf.resize(16384); }
    int rc = tf_prepare(map, kScratchFrame, &pose, &cam, ids_buf.data(), flag_buf.data(), (int64_t)ids_buf.size(), &n);
    if (rc == TF_ERR_CAPACITY && n > (int64_t)ids_buf.size()) {  // two-call pattern
      ids_buf.resize((size_t)n + (size_t)n / 2);
      flag_buf.resize(ids_buf.size());
      rc = tf_prepare(map, kScratchFrame, &pose, &cam, ids_buf.data(), flag_buf.data(), (int64_t)ids_buf.size(), &n);
    }
    check(rc);
    chunksIntersecting.reserve((size_t)n);
    for (int64_t i = 0; i < n; i++) {
      const ChunkID id = from_c(ids_buf[i]);
      chunksIntersecting.push_back(id);
      newChunkFlag.push_back(flag_buf[i] != 0);
      needsUpdateFlag.push_back(false);
      if (flag_buf[i]) chunkManager.note_exists(id);  // CreateChunk (:133-135); the others are known already
    }
  }

  // Structure/Chisel.h:218-249
  void IntegrateDepthScanColor(ProjectionIntegrator& integrator, float* depthImage, unsigned char* colorImage,
                               const Transform& depthExtrinsic, const PinholeCamera& depthCamera,
                               ChunkIDList& chunksIntersecting, std::vector<bool>& needsUpdateFlag, int integrate_flag,
                               int keyframeID = -1, float* observationQualityPointer = nullptr) {
    const size_t n = chunksIntersecting.size();
    if (n < 1) return;
    sync_truncation(integrator);
    check(tf_upload_frame(map, kScratchFrame, depthImage, colorImage, colorImage ? observationQualityPointer : nullptr));
    const tf_pose pose = to_c(depthExtrinsic);
    const tf_camera cam = depthCamera.c_camera();
    if (ids_buf.size() < n) { ids_buf.resize(n); flag_buf.resize(n); }
    if (q_buf.size() < n) q_buf.resize(n);
    for (size_t i = 0; i < n; i++) {
      ids_buf[i] = to_c(chunksIntersecting[i]);
      flag_buf[i] = needsUpdateFlag[i] ? 1 : 0;
    }
    const bool want_q = keyframeID >= 0 && colorImage && observationQualityPointer;
    int rc = tf_integrate(map, kScratchFrame, colorImage != nullptr, &pose, &cam, ids_buf.data(), (int64_t)n, integrate_flag,
                          flag_buf.data(), want_q ? q_buf.data() : nullptr);
    if (rc == TF_ERR_NOT_FOUND) throw std::out_of_range("ChunkManager::GetChunk");  // chunks.at() (Chisel.h:236)
    check(rc);
    for (size_t i = 0; i < n; i++) {
      needsUpdateFlag[i] = flag_buf[i] != 0;
      if (want_q && q_buf[i] > 0 && needsUpdateFlag[i])  // :244-247 (tf_integrate has vouched for the id)
        chunkManager.note_exists(chunksIntersecting[i])->observations[keyframeID] = q_buf[i];
    }
  }

  // Structure/Chisel.h:453-468: Prepare + Integrate(1) + Finalize, fused on the device
  void IntegrateDepthScanColor(ProjectionIntegrator& integrator, float* depthImage, unsigned char* colorImage,
                               const Transform& depthExtrinsic, const PinholeCamera& depthCamera) {
    sync_truncation(integrator);
    check(tf_upload_frame(map, kScratchFrame, depthImage, colorImage, nullptr));
    const tf_pose pose = to_c(depthExtrinsic);
    const tf_camera cam = depthCamera.c_camera();
    if (ids_buf.size() < fused_cap) { ids_buf.resize(fused_cap); flag_buf.resize(fused_cap); }
    if (upd_buf.size() < fused_cap) upd_buf.resize(fused_cap);
    tf_frame_stats st;
    check(tf_integrate_frame(map, kScratchFrame, colorImage != nullptr, &pose, &cam, &st, ids_buf.data(), flag_buf.data(),
                             upd_buf.data(), nullptr, (int64_t)fused_cap));
    if (st.n_chunks > (int64_t)fused_cap) {  // the map is updated, the lists were truncated: grow for the next frame, then fail loudly
      fused_cap = (size_t)st.n_chunks * 2;
      throw TexFusionError(TF_ERR_CAPACITY, "IntegrateDepthScanColor: chunk list longer than the list buffer");
    }
    last_stats = st;
    // FinalizeIntegrateChunks (:184-216): mark first, then GarbageCollect erases the never-updated new chunks
    for (int64_t i = 0; i < st.n_chunks; i++) {
      if (!upd_buf[i]) continue;
      const ChunkID id = from_c(ids_buf[i]);
      if (flag_buf[i]) chunkManager.note_exists(id);
      MarkMeshes(id);
    }
    for (int64_t i = 0; i < st.n_chunks; i++)
      if (flag_buf[i] && !upd_buf[i]) {  // created and removed on the device within this call
        const ChunkID id = from_c(ids_buf[i]);
        chunkManager.forget(id);
        meshesToUpdate.erase(id);
      }
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

  // Structure/Chisel.h:479-481
  void UpdateMeshes(const PinholeCamera& camera) { chunkManager.RecomputeMeshes(meshesToUpdate, camera); }

  // Structure/Chisel.cpp:112-147 (host-side flags only)
  void CompressMeshes(ChunkSet& chunksToUpdate) {
    const float gridResolution = chunkManager.GetResolution() * (8 / 8);
    MeshMap& allMeshes = chunkManager.GetAllMutableMeshes();
    for (const auto& it : chunksToUpdate) {
      if (!it.second) continue;
      auto mi = allMeshes.find(it.first);
      if (mi == allMeshes.end()) continue;
      mi->second->SimplifyByClustering(gridResolution, chunkManager.GetChunk(it.first)->GetOrigin());
    }
    for (const auto& it : chunksToUpdate) {
      if (!it.second) continue;
      auto mi = allMeshes.find(it.first);
      if (mi == allMeshes.end()) continue;
      MeshPtr meshit = mi->second;
      for (int k = 0; k < 6; k++) {
        auto ai = allMeshes.find(it.first + neighbourhood[k]);
        if (ai == allMeshes.end()) continue;
        MeshPtr meshadj = ai->second;
        if (!meshadj->simplified) continue;
        const int m = ((k % 2 == 0) ? k + 1 : k - 1);
        if (meshit->adj[k] == true && meshadj->adj[m] == false) meshadj->adj[m] = true;
        if (meshit->adj[k] == false && meshadj->adj[m] == true) meshit->adj[k] = true;
      }
    }
    chunksToUpdate.clear();
  }

  // Key-frame planes kept in HBM under frame_index for the atlas and the texcoord kernel
  // (Frame::refined_depth, rgb, colorValidFlag, observationQualityMap).
  void UploadKeyframe(int frame_index, const float* depth, const unsigned char* rgb, const unsigned char* colorValid,
                      const float* quality) {
    check(tf_upload_frame(map, frame_index, depth, nullptr, quality));
    check(tf_upload_keyframe_rgb(map, frame_index, rgb, colorValid));
    keyframes_on_device.insert(frame_index);
  }
  void ReleaseKeyframe(int frame_index) {
    if (keyframes_on_device.erase(frame_index)) tf_release_frame(map, frame_index);
  }

  // Chisel::GeneratePatches (Structure/Chisel.cpp:149-189).  UniGraph: `chunks` (ChunkID -> node) and
  // get_label(node) (Structure/uni_graph.h); FrameT: the reference's Frame (GCSLAM/frame.h: rgb,
  // refined_depth, colorValidFlag, observationQualityMap as cv::Mat-like objects with .data,
  // pose_sophus[0], frame_index).  Patch::CalculateTexCoords (Structure/Patch.cpp:40-108) runs on the
  // device, one tf_patch_texcoords call per key-frame that is some chunk's label.
  template <class UniGraph, class FrameT>
  int GeneratePatches(ChunkIDList& chunksToUpdate, UniGraph& labelset, std::vector<FrameT>& frame_list, PinholeCamera& cameraModel) {
    std::size_t loc_start = Atlas::MAX_PATCH_HEIGHT * Atlas::MAX_PATCH_WIDTH;
    std::size_t loc_end = 0;
    std::map<int, std::vector<PatchPtr>> by_frame;  // label -> patches, in list order
    for (size_t i = 0; i < chunksToUpdate.size(); i++) {
      if (!chunkManager.HasMesh(chunksToUpdate[i])) continue;
      MeshPtr meshit = chunkManager.GetMutableMesh(chunksToUpdate[i]);
      const int frameid = (int)labelset.get_label(labelset.chunks.find(chunksToUpdate[i])->second);
      PatchPtr piece;
      try {
        piece = atlas.AddPatch(meshit);
      } catch (std::exception&) {
        return -1;  // atlas full (:167-173)
      }
      piece->SetFrameid(frameid);
      by_frame[frameid].push_back(piece);
      if (piece->texloc < loc_start) loc_start = piece->texloc;
      if (piece->texloc > loc_end) loc_end = piece->texloc;
    }
    for (auto& kv : by_frame) {
      FrameT& frame = frame_list[kv.first];
      if (!keyframes_on_device.count(kv.first))
        UploadKeyframe(kv.first, (const float*)frame.refined_depth.data, (const unsigned char*)frame.rgb.data,
                       (const unsigned char*)frame.colorValidFlag.data, (const float*)frame.observationQualityMap.data);
      // T_g_l = pos.pose_sophus[0].inverse().matrix().cast<float>() (Structure/Patch.cpp:51)
      const auto T = frame.pose_sophus[0].inverse().matrix().template cast<float>();
      float T_g_l[16];
      for (int c = 0; c < 4; c++)
        for (int r = 0; r < 4; r++) T_g_l[c * 4 + r] = T(r, c);
      CalculateTexCoords(kv.first, T_g_l, cameraModel, kv.second);
      for (PatchPtr& piece : kv.second) piece->SetImage(frame.rgb);
    }
    atlas.hot_start = (loc_start / Atlas::MAX_PATCH_WIDTH) * Atlas::MAX_PATCH_WIDTH;                        // :184-186
    atlas.hot_end = (loc_end / Atlas::MAX_PATCH_WIDTH + atlas.PATCH_HEIGHT) * Atlas::MAX_PATCH_WIDTH;
    return 0;
  }

  // Patch::CalculateTexCoords for all patches that share one key-frame, in one launch
  void CalculateTexCoords(int frame_index, const float* T_g_l, const PinholeCamera& camera, std::vector<PatchPtr>& patches) {
    tc_off.assign(1, 0);
    tc_v.clear();
    tc_c.clear();
    for (const PatchPtr& p : patches) {
      const Mesh& mesh = *p->mesh;
      for (size_t j = 0; j < mesh.vertices.size(); j++)
        for (int k = 0; k < 3; k++) { tc_v.push_back(mesh.vertices[j](k)); tc_c.push_back(mesh.colors[j](k)); }
      tc_off.push_back(tc_off.back() + (int64_t)mesh.vertices.size());
    }
    const int64_t nv = tc_off.back();
    tc_tc.resize((size_t)std::max<int64_t>(nv, 1) * 2);
    tc_col.resize((size_t)std::max<int64_t>(nv, 1) * 3);
    tc_res.resize(patches.size());
    CalculateTexCoords(frame_index, T_g_l, camera, (int64_t)patches.size(), tc_off.data(), tc_v.data(), tc_c.data(), tc_tc.data(),
                       tc_col.data(), tc_res.data());
    for (size_t k = 0; k < patches.size(); k++) {
      Patch& p = *patches[k];
      const int64_t a = tc_off[k], b = tc_off[k + 1];
      p.texcoord.resize((size_t)(b - a));
      p.texcolor.resize((size_t)(b - a));
      for (int64_t j = a; j < b; j++) {
        p.texcoord[(size_t)(j - a)] = Vec2(tc_tc[2 * j], tc_tc[2 * j + 1]);
        p.texcolor[(size_t)(j - a)] = Vec3(tc_col[3 * j], tc_col[3 * j + 1], tc_col[3 * j + 2]);
      }
      p.boundingbox = Box(tc_res[k].x, tc_res[k].y, tc_res[k].w, tc_res[k].h);
      p.wrong_mapping = tc_res[k].wrong_mapping != 0;
    }
  }
  void CalculateTexCoords(int frame_index, const float* T_g_l, const PinholeCamera& camera, int64_t n_patches,
                          const int64_t* offsets, const float* vertices, const float* colors, float* texcoord,
                          float* texcolor, tf_patch_result* results) {
    tf_pose T;
    std::memcpy(T.m, T_g_l, sizeof(T.m));
    const tf_camera cam = camera.c_camera();
    check(tf_patch_texcoords(map, frame_index, &T, &cam, n_patches, offsets, vertices, colors, texcoord, texcolor, results));
  }

  // Structure/Chisel.cpp:191-196 (+ the hot rows of the host mirror for the GL upload)
  void UpdateAtlas(ChunkIDList& chunksToUpdate) {
    atlas.UpdateBuffers(chunksToUpdate);
    if (mirror_hot_rows) atlas.SyncHotRange();
  }

  // Chisel::CompensateColor (Structure/Chisel.cpp:198-281) is per-vertex host post-processing with
  // Eigen's SelfAdjointEigenSolver on the fields kept above (texcolor, mesh->colors, labs, labt,
  // has_adjusted); it is not part of the device path: keep the reference's definition.
  void CompensateColor();

  // Chisel::DrawMeshes (Structure/Chisel.cpp:283-355): 12 floats per vertex for the GL buffers
  void DrawMeshes(float* vertices, unsigned int* indices, unsigned int& tsdf_indice_num, unsigned int& tsdf_vertice_num) {
    unsigned int vert_num = 0, index_num = 0;
    for (auto& it : chunkManager.GetAllMeshes()) {
      MeshPtr mesh = it.second;
      PatchPtr patch = mesh->m_patch;
      if (patch == nullptr || !patch->complete()) continue;
      // (a mesh that was re-meshed after its patch was computed has texcoords of the old vertex count;
      //  the reference indexes them anyway — out of bounds — this shim skips the chunk until it is textured again)
      if (patch->texcoord.size() != mesh->vertices.size()) continue;
      for (std::size_t j = 0; j < mesh->indices.size(); j++) indices[index_num++] = (unsigned int)mesh->indices[j] + vert_num;
      const Vec2 loc = atlas.GetTexLoc(mesh->chunkID);
      for (std::size_t j = 0; j < mesh->vertices.size(); j++) {
        float* cur = &vertices[12 * (std::size_t)vert_num];
        const Vec3 &vert = mesh->vertices[j], &color = mesh->colors[j], &normal = mesh->normals[j];
        Vec2 tc = patch->texcoord[j];
        if (patch->ratio(0) < 1) tc(0) *= patch->ratio(0);
        if (patch->ratio(1) < 1) tc(1) *= patch->ratio(1);
        tc(0) += loc(0);
        tc(1) += loc(1);
        cur[0] = vert(0), cur[1] = vert(1), cur[2] = vert(2), cur[3] = 50;
        int rgb_value = int(color(0) * 255);
        rgb_value = (rgb_value << 8) + int(color(1) * 255);
        rgb_value = (rgb_value << 8) + int(color(2) * 255);
        cur[4] = (float)rgb_value;
        if (patch->has_adjusted && !patch->labs.empty()) {
          int ad = int((patch->labs[j](0) - patch->texcolor[j](0)) * 255) + 255;
          ad = (ad << 9) + int((patch->labs[j](1) - patch->texcolor[j](1)) * 255) + 255;
          ad = (ad << 9) + int((patch->labs[j](2) - patch->texcolor[j](2)) * 255) + 255;
          cur[5] = (float)ad;
        } else {
          cur[5] = 0;
        }
        cur[6] = tc(0) / Atlas::MAX_PATCH_WIDTH;
        cur[7] = tc(1) / Atlas::MAX_PATCH_HEIGHT;
        cur[8] = normal(0), cur[9] = normal(1), cur[10] = normal(2);
        cur[11] = patch->wrong_mapping ? 1.0f : 0.0f;
        vert_num++;
      }
      patch->has_updated = true;
    }
    tsdf_indice_num = index_num;
    tsdf_vertice_num = vert_num;
  }

  // Chisel::SaveAllMeshesToPLY (Structure/Chisel.cpp:357-379) + SaveMeshPLYASCII
  // (3rd_party/open_chisel/io/PLY.cpp:27-80): one vertex per index, ASCII, colours as uchar
  bool SaveAllMeshesToPLY(const std::string& filename) {
    std::size_t v = 0;
    for (const auto& it : chunkManager.GetAllMeshes()) v += it.second->indices.size();
    std::ofstream stream(filename.c_str());
    if (!stream) return false;
    stream << "ply" << std::endl << "format ascii 1.0" << std::endl << "element vertex " << v << std::endl;
    stream << "property float x" << std::endl << "property float y" << std::endl << "property float z" << std::endl;
    if (v > 0) stream << "property uchar red" << std::endl << "property uchar green" << std::endl << "property uchar blue" << std::endl;
    stream << "element face " << v / 3 << std::endl << "property list uchar int vertex_index" << std::endl << "end_header" << std::endl;
    for (const auto& it : chunkManager.GetAllMeshes()) {
      const Mesh& m = *it.second;
      for (std::size_t i = 0; i < m.indices.size(); i++) {
        const Vec3 &vert = m.vertices[m.indices[i]], &color = m.colors[m.indices[i]];
        stream << vert(0) << " " << vert(1) << " " << vert(2) << " " << static_cast<int>(color(0) * 255.0f) << " "
               << static_cast<int>(color(1) * 255.0f) << " " << static_cast<int>(color(2) * 255.0f) << std::endl;
      }
    }
    for (std::size_t i = 0; i + 2 < v; i += 3) stream << "3 " << i << " " << i + 1 << " " << i + 2 << " " << std::endl;
    return (bool)stream;
  }

  tf_map* handle() { return map; }
  tf_frame_stats last_stats{};
  bool mirror_hot_rows = true;  // UpdateAtlas keeps atlas.texture_buffer[hot range] current for the GL thread

  ChunkManager chunkManager;
  ChunkSet meshesToUpdate;
  Atlas atlas;

 protected:
  void check(int rc) const { if (rc < 0) throw TexFusionError(rc, tf_last_error(map)); }
  void sync_truncation(const ProjectionIntegrator& integrator) {
    const tf_truncation t = integrator.c_truncation();
    if (std::memcmp(&t, &trunc, sizeof(t)) != 0) {
      check(tf_set_truncation(map, &t));
      trunc = t;
    }
  }
  void MarkMeshes(const ChunkID& id) {  // Structure/Chisel.h:197-203
    meshesToUpdate[id] = true;
    for (int k = 0; k < 6; k++) meshesToUpdate[id + neighbourhood[k]] = true;
  }
  static const int32_t kScratchFrame = 0x7F000000;  // frame-store slot for images passed by pointer
  tf_map* map = nullptr;
  tf_truncation trunc{};
  size_t npix = 0;
  size_t fused_cap = 1 << 17;
  std::vector<tf_chunk_id> ids_buf;
  std::vector<uint8_t> flag_buf, upd_buf;
  std::vector<float> q_buf;
  std::unordered_set<int> keyframes_on_device;
  std::vector<int64_t> tc_off;
  std::vector<float> tc_v, tc_c, tc_tc, tc_col;
  std::vector<tf_patch_result> tc_res;
};
typedef std::shared_ptr<Chisel> ChiselPtr;

}  // namespace chisel
