/*
 * texfusion.h — C ABI of libtexfusion_b200.so
 *
 * B200-native (sm_100a) replacement for the per-frame fusion hot path of
 * THU-luvision/TextureFusion (FlashFusion).  The reference has no FFI layer: its boundary
 * is the public C++ surface of chisel::Chisel / ChunkManager / ProjectionIntegrator / Atlas
 * called from GCFusion/MobileFusion.cpp.  Each entry point below names the reference
 * interface (file:line, relative to the reference tree) whose work it takes over; the C++
 * shim in texturefusion_b200/host/ keeps those class/method names on top of this ABI
 * (see INTEGRATION.md).
 *
 * Conventions
 *   - plain pointers + sizes, POD structs, no C++/torch types, no exceptions.
 *   - every function returns 0 on success, <0 on error (tf_last_error gives the text);
 *     functions that are pure queries say so.
 *   - a tf_map is not thread-safe; one caller thread at a time (the reference's map
 *     thread, GCFusion/MobileFusion.cpp:99-112).  Kernels and read-backs of one map run on one
 *     CUDA stream, frame uploads on a second (copy) stream; a call returns once the results
 *     it hands to the host are visible.
 *   - tf_upload_frame / tf_upload_keyframe_rgb are asynchronous: pageable sources are staged
 *     before the call returns; page-locked sources (tf_host_alloc) are DMA'd directly and
 *     must stay untouched until a call that consumes that frame has returned, or tf_sync.
 *     A caller can therefore upload frame i+1 and then fuse frame i: the copy overlaps the
 *     kernels.
 *   - output arrays that are page-locked (tf_host_alloc / cudaHostAlloc / cudaHostRegister)
 *     and 16-byte aligned are written by the device directly; any other pointer goes through
 *     an internal staging buffer and a host copy (same results, a few microseconds more).
 *   - environment: TEXFUSION_B200_GRAPH=0 (plain kernel launches instead of a CUDA graph per
 *     frame), TEXFUSION_B200_PDL=0 (no programmatic dependent launch).
 *   - there is NO CPU fallback: without a CUDA device tf_create fails with TF_ERR_CUDA.
 */
#ifndef TEXFUSION_B200_H
#define TEXFUSION_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TF_ABI_VERSION 2

enum {
  TF_OK = 0,
  TF_ERR_INVALID = -1,    /* bad argument                                   */
  TF_ERR_CUDA = -2,       /* CUDA runtime error / no device                 */
  TF_ERR_CAPACITY = -3,   /* chunk pool, candidate grid or out buffer full  */
  TF_ERR_NOT_FOUND = -4,  /* unknown frame_index or chunk id (ref: unordered_map::at throws,
                             Structure/ChunkManager.h:137-139)              */
  TF_ERR_ATLAS_FULL = -5  /* ref: std::overflow_error in Atlas::AddPatch, Structure/Atlas.cpp:52-53 */
};

typedef struct tf_map tf_map; /* opaque: one volumetric map on one GPU (one rank) */

/* == Eigen::Vector3i layout; chisel::ChunkID (3rd_party/open_chisel/geometry/Geometry.h) */
typedef struct { int32_t x, y, z; } tf_chunk_id;

/* chisel::PinholeCamera (3rd_party/open_chisel/camera/PinholeCamera.h:33-75).  Pass the
 * float intrinsics given to SetIntrinsics; the library applies the int truncation of
 * GetFx/GetFy/GetCx/GetCy (:46-49) itself.  width / height must be the frame size the map was
 * created with and near_plane must be >= 0 (else TF_ERR_INVALID). */
typedef struct {
  float fx, fy, cx, cy;
  int32_t width, height;
  float near_plane, far_plane;
} tf_camera;

/* QuadraticTruncator(quad, lin, cst, scale) + ConstantWeighter(weight)
 * (GCFusion/MobileFusion.h:215-228). */
typedef struct { float quad, lin, cst, scale, weight; } tf_truncation;

/* chisel::Transform == Eigen::Affine3f, column-major 4x4, camera -> world
 * (pose_sophus[k].matrix().cast<float>(), GCFusion/MobileFusion.cpp:133,136). */
typedef struct { float m[16]; } tf_pose;

typedef struct {
  int32_t chunk_dim;      /* voxels per chunk edge; only 8 is supported (GCFusion/MobileFusion.h:231-233) */
  float voxel_res;        /* metres */
  int32_t use_color;      /* chisel::Chisel ctor arg (GCFusion/MobileFusion.h:239-241) */
  tf_truncation trunc;
  int32_t device;         /* CUDA ordinal */
  int32_t n_ranks, rank;  /* chunk sharding: this map owns chunks with owner(id) == rank */
  int64_t max_chunks;     /* chunk-pool capacity (8 KiB each); 0 = default */
  int32_t max_frames;     /* frame-store slots, allocated up front: 4 B (16 B with use_color) per
                             pixel and slot; 0 = default (32) */
  int32_t width, height;  /* frame size of the store; 0 = 640x480 */
  int32_t dot3_order;     /* association of the reference's 3-term Eigen products (Rt * v at
                             ProjectionIntegrator.cpp:88-89, Structure/Chisel.cpp:67-69,
                             Structure/ChunkManager.h:429-453,521-524), which depends on the Eigen the
                             reference was built with: 0 = Eigen >= 3.3, x0 + (x1 + x2);
                             1 = Eigen 3.2, (x0 + x1) + x2.  tools/ref_golden/probe_eigen_order.cpp
                             prints the value for an installed Eigen. */
} tf_config;

/* chisel::Chisel::Chisel + ChunkManager ctor (Structure/Chisel.cpp:38-41). */
int tf_create(tf_map** out, const tf_config* cfg);
void tf_destroy(tf_map* m);
/* text of the last error on this map (m == NULL: last tf_create failure). Query. */
const char* tf_last_error(const tf_map* m);
/* Chisel::Reset (Structure/Chisel.cpp:47-50): drop all chunks, keep frames and atlas. */
int tf_reset(tf_map* m);
/* ProjectionIntegrator::SetTruncator / SetWeighter (GCFusion/MobileFusion.h:245-249): the reference
 * reads them on every voxelUpdateSIMD call (ProjectionIntegrator.cpp:90-92), so they can change
 * between calls; takes effect from the next prepare / integrate call. */
int tf_set_truncation(tf_map* m, const tf_truncation* t);

/* pinned host memory for frame / result buffers */
void* tf_host_alloc(size_t bytes);
void tf_host_free(void* p);

/* ---- frame store ---------------------------------------------------------------------
 * Frame::refined_depth / RGBA packed by ReIntegrateKeyframe / observationQualityMap
 * (GCSLAM/frame.h:35-66, GCFusion/MobileFusion.cpp:147-162).  depth: float32 metres,
 * width*height; rgba: 4 bytes per pixel, A = colour-valid (1/0); quality: float32.
 * rgba/quality may be NULL (depth-only local frame, GCFusion/MobileFusion.cpp:198-202).
 * Re-uploading a frame_index replaces it: planes that are not passed are absent afterwards.
 * When the store is full the least-recently-used frame that is not pinned (see
 * tf_upload_keyframe_rgb) is replaced; TF_ERR_CAPACITY if every slot is pinned. */
int tf_upload_frame(tf_map* m, int32_t frame_index, const float* depth,
                    const uint8_t* rgba_or_null, const float* quality_or_null);
/* Key-frame colour as the reference stores it: Frame::rgb (8UC3) + colorValidFlag (8U).
 * Keeps rgb for the atlas (Patch::SetImage, Structure/Patch.cpp:172-175) and packs the
 * RGBA plane on the device — replaces the scalar pack loop GCFusion/MobileFusion.cpp:151-162
 * (color_valid == NULL: alpha = 1 everywhere, as IntegrateFrame :237-242).  The slot is pinned —
 * never evicted — until tf_release_frame: the reference keeps every key-frame's rgb for
 * Atlas::UpdateBuffer, so size tf_config.max_frames for the key-frames that stay texturable. */
int tf_upload_keyframe_rgb(tf_map* m, int32_t frame_index, const uint8_t* rgb,
                           const uint8_t* color_valid_or_null);
int tf_release_frame(tf_map* m, int32_t frame_index);
/* Device addresses of a frame's planes (allocating the slot if needed) so that a
 * collective (NCCL broadcast over NVLink) can land directly in the store.  Any out
 * pointer may be NULL.  has_color != 0 marks rgba/quality as present. */
int tf_frame_device_ptrs(tf_map* m, int32_t frame_index, int has_color, void** depth,
                         void** rgba, void** quality);

/* ---- multi-GPU: frame broadcast (no reference counterpart) --------------------
 * One process per GPU, each with a tf_map created with the same n_ranks and its own rank; chunks
 * are owned by rank (tf_config.n_ranks / rank), so the only per-frame exchange is the frame itself.
 * tf_comm_unique_id: ncclGetUniqueId on one rank; the caller hands the 128 bytes to the others
 * (MPI, torch.distributed, a file, ...).  tf_comm_init: ncclCommInitRank on every rank (collective).
 * NCCL is loaded at run time (libnccl.so.2; TEXFUSION_B200_NCCL overrides the path).
 * tf_broadcast_frame (collective, asynchronous): the root has uploaded frame_index
 * (tf_upload_frame); its depth (+ rgba + quality when has_color) planes land in every other rank's
 * frame store with ONE ncclBroadcast queued on the upload stream, i.e. behind the root's own
 * host-to-device copy and next to the kernels of the frame being fused.  Consumers of the frame wait
 * for it on the device; tf_wait_upload waits on the host. */
int tf_comm_unique_id(uint8_t id_out[128]);
int tf_comm_init(tf_map* m, const uint8_t id[128]);
int tf_broadcast_frame(tf_map* m, int32_t frame_index, int has_color, int root);

/* ---- per-frame hot path -------------------------------------------------------------- */

/* Chisel::PrepareIntersectChunks (Structure/Chisel.h:103-140): depth bbox
 * (ChunkManager::GetBoundaryChunkID, Structure/ChunkManager.h:366-378), coarse/fine
 * culling (GetChunkIDsObservedByCamera, :380-559) and HasChunk/CreateChunk for every hit.
 * ids_out is in the reference's traversal order; is_new_out[i] in {0,1}.  If more than cap
 * chunks are hit, *n_out is the required size and TF_ERR_CAPACITY is returned (the map is
 * then unchanged).  With n_ranks > 1 only chunks owned by this rank are listed. */
int tf_prepare(tf_map* m, int32_t frame_index, const tf_pose* pose, const tf_camera* cam,
               tf_chunk_id* ids_out, uint8_t* is_new_out, int64_t cap, int64_t* n_out);

/* Chisel::IntegrateDepthScanColor, list form (Structure/Chisel.h:218-249) ->
 * ProjectionIntegrator::voxelUpdateSIMD (3rd_party/open_chisel/utils/ProjectionIntegrator.cpp:67-426).
 * flag 1 = integrate, 0 = de-integrate.  needs_update_inout[i] |= updated.
 * quality_out_or_null[i] = the chunk's raw chunkObservationQuality (only meaningful when
 * use_color and the frame has a quality plane); the caller applies
 * `keyframeID >= 0 && q > 0 && needsUpdate` (Structure/Chisel.h:244-247). */
int tf_integrate(tf_map* m, int32_t frame_index, int use_color, const tf_pose* pose,
                 const tf_camera* cam, const tf_chunk_id* ids, int64_t n, int flag,
                 uint8_t* needs_update_inout, float* quality_out_or_null);

/* One group of frames applied, in order, to ONE chunk list with each chunk's voxels held on
 * chip across the group (a key-frame followed by its local depth frames,
 * GCFusion/MobileFusion.cpp:176-203).  Same result as n_frames tf_integrate calls. */
typedef struct {
  int32_t frame_index;
  int32_t use_color;
  int32_t flag; /* 1 integrate, 0 de-integrate */
  int32_t reserved;
  tf_pose pose;
} tf_group_frame;
int tf_integrate_group(tf_map* m, const tf_group_frame* frames, int32_t n_frames,
                       const tf_camera* cam, const tf_chunk_id* ids, int64_t n,
                       uint8_t* needs_update_inout, float* quality_out_or_null /* frame 0 */);

/* Device half of Chisel::FinalizeIntegrateChunks / GarbageCollect
 * (Structure/Chisel.h:184-216,472-477): ChunkManager::RemoveChunk for each id. */
int tf_remove_chunks(tf_map* m, const tf_chunk_id* ids, int64_t n);

/* Chisel::IntegrateDepthScanColor, convenience form (Structure/Chisel.h:453-468) =
 * Prepare + Integrate(flag 1) + Finalize, as MobileFusion::IntegrateFrame uses it
 * (GCFusion/MobileFusion.cpp:223-250) — fused: no host round trip between the stages.
 * Outputs (any may be NULL; cap = capacity of each array): the intersecting list, its
 * new/updated flags and the raw quality sums.  Never-updated new chunks are already
 * removed on return; the caller derives meshesToUpdate / validChunks from `updated`. */
typedef struct {
  int64_t n_chunks;        /* |chunksIntersecting| */
  int64_t n_new;           /* created by this frame */
  int64_t n_updated;       /* needsUpdate == true */
  int64_t n_removed;       /* new && !updated -> garbage collected */
  int64_t voxel_updates;   /* 512 * n_chunks */
} tf_frame_stats;
int tf_integrate_frame(tf_map* m, int32_t frame_index, int use_color, const tf_pose* pose,
                       const tf_camera* cam, tf_frame_stats* stats_out, tf_chunk_id* ids_out,
                       uint8_t* is_new_out, uint8_t* updated_out, float* quality_out, int64_t cap);

/* The same call in two halves, for a streaming caller: _begin queues the frame's kernels and returns at
 * once; between _begin and _end only frame-store calls of OTHER frames are allowed (tf_upload_frame,
 * tf_broadcast_frame, tf_wait_upload: the next frame's ingest then overlaps this frame's kernels on
 * the host side too); _end waits, fills the output arrays handed to _begin and returns the frame's
 * status.  tf_integrate_frame == _begin + _end. */
int tf_integrate_frame_begin(tf_map* m, int32_t frame_index, int use_color, const tf_pose* pose, const tf_camera* cam,
                             tf_chunk_id* ids_out, uint8_t* is_new_out, uint8_t* updated_out, float* quality_out,
                             int64_t cap);
int tf_integrate_frame_end(tf_map* m, tf_frame_stats* stats_out);

/* One step of the streaming loop above as ONE call (what a caller in a language with an expensive foreign-call
 * boundary wants): tf_integrate_frame_begin(frame) | tf_upload_frame(next) (when next_depth is given) |
 * tf_broadcast_frame(next) (maps with a communicator, when next_index >= 0) | tf_integrate_frame_end | tf_wait_upload(wait_index)
 * (when wait_index >= 0).  Same results and errors as the separate calls; the first error ends the step. */
typedef struct {
  int32_t frame_index, use_color;   /* the frame to fuse (its ingest was queued by an earlier step) */
  tf_pose pose;
  tf_chunk_id* ids_out;             /* as tf_integrate_frame */
  uint8_t* is_new_out;
  uint8_t* updated_out;
  float* quality_out;
  int64_t cap;
  int32_t next_index;               /* frame to ingest while `frame_index` is being fused; < 0: none */
  int32_t next_has_color;           /* (sharded maps) planes tf_broadcast_frame moves */
  const float* next_depth;          /* host planes of that frame on the ingest rank, NULL on the others */
  const uint8_t* next_rgba;
  const float* next_quality;
  int32_t broadcast_root;
  int32_t wait_index;               /* frame whose ingest has completed when the call returns; < 0: none */
} tf_stream_step_args;
int tf_stream_step(tf_map* m, const tf_camera* cam, const tf_stream_step_args* step, tf_frame_stats* stats_out);

/* Loop-closure path (GCFusion/MobileFusion.cpp:301-310): one item = one
 * ReIntegrateKeyframe call (:114-221).  flag 0: de-integrate the key-frame and its local
 * frames with their old poses over `ids` (= kf.validChunks).  flag 1: Prepare with the
 * key-frame's depth under its new pose, integrate the group, Finalize; the new valid list
 * is written to valid_out (cap entries) / *n_valid_out.
 * The items are applied in order but queued without intermediate host synchronisation (one per 32
 * re-integrations); outputs are valid when the call returns.  Errors are therefore reported for
 * the batch as a whole: a de-integration id that is not in the map is skipped and the call returns
 * TF_ERR_NOT_FOUND after the remaining items have been applied.  A flag-1 item whose three output
 * pointers are all NULL only fuses its frames (a sequence of such single-frame items streams a
 * recording into the map without a host round trip per frame). */
typedef struct {
  int32_t flag;
  int32_t n_frames;               /* 1 key-frame + local frames */
  const tf_group_frame* frames;   /* frames[0] is the key-frame */
  const tf_chunk_id* ids;         /* flag 0: chunk list to de-integrate */
  int64_t n_ids;
  tf_chunk_id* valid_out;         /* flag 1 */
  float* quality_out;             /* flag 1: raw quality sum per valid_out entry (or NULL) */
  int64_t cap;
  int64_t* n_valid_out;
} tf_batch_item;
int tf_integrate_batch(tf_map* m, const tf_batch_item* items, int64_t n_items,
                       const tf_camera* cam);

/* ---- chunk map queries (ChunkManager::HasChunk/GetChunks, Structure/ChunkManager.h:126-139) */
int tf_has_chunk(tf_map* m, tf_chunk_id id);  /* 1 / 0, <0 on error */
int64_t tf_chunk_count(tf_map* m);
int tf_list_chunks(tf_map* m, tf_chunk_id* out, int64_t cap, int64_t* n_out);
/* Read-back for the CPU mesher (Structure/ChunkManager.cpp:614-626) and for parity:
 * sdf[512], weight[512], color[2048] per chunk in the reference layout
 * (voxel = (z*8+y)*8+x, 3rd_party/open_chisel/geometry/Chunk.h:91-93). */
int tf_download_chunks(tf_map* m, const tf_chunk_id* ids, int64_t n, float* sdf,
                       float* weight, uint16_t* color);

/* ---- meshing ------------------------------------------------------------
 * ChunkManager::RecomputeMeshes -> GenerateMeshEfficient (Structure/ChunkManager.cpp:232-264, 595-1002;
 * gradient normals :277-455; triangle table 3rd_party/open_chisel/marching_cubes/MarchingCubes.cpp) for a
 * list of chunks, on the voxels where they lie in HBM — replaces tf_download_chunks + the CPU mesher
 * in the per-key-frame loop (Chisel::UpdateMeshes, GCFusion/MobileFusion.cpp:327).
 * Chunk i owns vertices [vert_off[i], vert_off[i+1]) and indices [idx_off[i], idx_off[i+1])
 * (vert_off / idx_off have n + 1 entries); vertices / normals / colors are xyz float triples in the
 * order of mesh->vertices / normals / colors, indices are local to the chunk's vertex block like
 * mesh->indices.  Ids that are not in the map give empty meshes (:239-241).  With all four output
 * arrays NULL only the offsets are computed (size query); TF_ERR_CAPACITY if vert_cap (vertices)
 * or idx_cap (indices) is too small — the offsets then hold the required sizes. */
int tf_mesh_chunks(tf_map* m, const tf_chunk_id* ids, int64_t n, int64_t* vert_off, int64_t* idx_off,
                   float* vertices, float* normals, float* colors, int32_t* indices, int64_t vert_cap,
                   int64_t idx_cap);

/* ---- texture atlas (Structure/Atlas.{h,cpp}) ---------------------------------------- */
/* Atlas::AddPatch placement (Structure/Atlas.cpp:43-64): first call for an id hands out
 * loc_next and advances it; later calls return the same texloc. */
int tf_atlas_alloc_slot(tf_map* m, tf_chunk_id id, uint64_t* texloc_out);
/* Atlas::UpdateBuffer (Structure/Atlas.cpp:71-91) for a batch of patches: crop
 * (x,y,w,h) of key-frame frame_index's rgb (cv::Rect semantics, Patch::SetImage) is copied
 * to the slot at texloc, or bilinearly resized (cv::resize INTER_LINEAR) to the slot size
 * when it does not fit. */
typedef struct {
  uint64_t texloc;
  int32_t frame_index;
  int32_t x, y, w, h;
} tf_patch_desc;
int tf_atlas_update(tf_map* m, const tf_patch_desc* patches, int64_t n);
/* texture_buffer rows for the GL upload (GCFusion/MobileFusion.h:404-427): bytes
 * [hot_start*3, hot_end*3) of the 13824x13824x3 atlas, hot_* in pixels as Atlas::hot_start. */
int tf_atlas_download(tf_map* m, uint64_t hot_start, uint64_t hot_end, uint8_t* rgb_out);
/* The same bytes copied device to device into dst_device (device memory of any GPU of the process),
 * complete on return: the CUDA side of a CUDA-GL interop upload.  The caller registers its pixel-unpack
 * buffer once (cudaGraphicsGLRegisterBuffer), maps it, passes the mapped pointer here and unmaps it before
 * glTexSubImage2D — replacing glBufferDataARB(&texture_buffer.data[hot_start*3]) of
 * GCFusion/MobileFusion.h:406-412 (binding shown in INTEGRATION.md). */
int tf_atlas_copy_to_device(tf_map* m, uint64_t hot_start, uint64_t hot_end, void* dst_device);
int tf_atlas_patch_size(tf_map* m, int32_t* patch_w, int32_t* patch_h);

/* Patch::CalculateTexCoords (Structure/Patch.cpp:40-108, bilinear :110-146, bilinear_depth
 * :148-170) for a batch of chunk meshes against key-frame frame_index (its rgb and depth must be
 * in the store).  world_to_camera = pose_sophus[0].inverse().matrix().cast<float>() (:51).
 * Patch p owns vertices [vertex_offsets[p], vertex_offsets[p+1]); vertices / colors are xyz / rgb
 * float triples (mesh->vertices, mesh->colors).  Outputs: texcoord (2 floats per vertex, already
 * shifted by the bounding box like :101-103), texcolor (3 floats), and per patch the bounding box
 * (cv::Rect), wrong_mapping (:87-96) and the return flag (-1 if a vertex left the image). */
typedef struct { int32_t x, y, w, h; int32_t wrong_mapping; int32_t flag; } tf_patch_result;
int tf_patch_texcoords(tf_map* m, int32_t frame_index, const tf_pose* world_to_camera, const tf_camera* cam,
                       int64_t n_patches, const int64_t* vertex_offsets, const float* vertices,
                       const float* colors, float* texcoord_out, float* texcolor_out, tf_patch_result* results);

/* ---- frame pre-processing ------------------------------------------------
 * The per-pixel loops that the reference runs between loading a frame and fusing it
 * (main.cpp:117-147), on the planes of the frame store; they replace
 *   BasicAPI::extractNormalMapSIMD      BasicAPI.cpp:849-905   -> tf_pre_normal_map
 *   BasicAPI::refineKeyframesSIMD       BasicAPI.cpp:506-636   -> tf_pre_refine_keyframe
 *   BasicAPI::refineNewframesSIMD       BasicAPI.cpp:378-442   -> tf_pre_refine_newframe
 *   BasicAPI::refineDepthUseNormalSIMD  BasicAPI.cpp:728-780   -> tf_pre_refine_depth_by_normal
 *   BasicAPI::checkColorQuality + estimateColorQuality (:783-847) and the RGBA pack of
 *   GCFusion/MobileFusion.cpp:151-162                          -> tf_pre_color_quality
 * A frame enters with tf_upload_frame(depth only: Frame::refined_depth after framePreprocess); a
 * key-frame then needs ONE more upload, its 3-byte RGB, instead of the RGBA, colour-valid and
 * quality planes.  The calls are asynchronous (queued on the ingest stream behind the upload; the
 * fusion calls wait for them like for an upload).
 * `cam` carries the camera's FLOAT intrinsics as main.cpp passes them (camera.c_fx ...): they are
 * NOT truncated to integers here (near_plane / far_plane are ignored).  Transforms are 3x4 row-major
 * [R | t] in float: ref_to_new = (new.pose_sophus[0].inverse() * ref.pose_sophus[0]).matrix().block<3,4>(0,0)
 * cast to float (BasicAPI.cpp:529-538), new_to_ref the inverse product (:395-399).
 * Per frame the library keeps a normal map (3 planes) and the refinement weights (Frame::weight,
 * zero for a new upload) for the four most recently touched frames.
 * Where the reference leaves memory unwritten (cv::Mat::create: the normal map's border, colour flags
 * of rejected pixels) the planes are 0.  _mm256_rsqrt_ps is reproduced from the instruction's value
 * table (Intel), see texturefusion_b200/csrc/tf_rsqrt_table.h.  Requires width % 8 == 0. */
/* framePreprocess (Tools/DatasetWrapper.hpp:187-263, BasicAPI.cpp:942-1004), optional: the raw 16-bit depth image
 * instead of tf_upload_frame's float plane (values above max_depth * depth_scale are dropped, then
 * metres = value / depth_scale), and cv::bilateralFilter(refined_depth, d, sigma_color, sigma_space) — the reference
 * calls it with (9, 0.03, 10) — restated from OpenCV's float path (4096-bin exp table with interpolation,
 * disc of radius d/2, BORDER_REFLECT_101).  The filter agrees with cv2 to float rounding (accumulation order), not
 * bit for bit: a caller that needs the reference's exact refined depth keeps cv::bilateralFilter on the host and
 * uses tf_upload_frame. */
int tf_pre_upload_depth_u16(tf_map* m, int32_t frame_index, const uint16_t* depth, float depth_scale, float max_depth);
int tf_pre_bilateral(tf_map* m, int32_t frame_index, int32_t d, float sigma_color, float sigma_space);
int tf_pre_normal_map(tf_map* m, int32_t frame_index, const tf_camera* cam);
int tf_pre_refine_keyframe(tf_map* m, int32_t keyframe_index, int32_t new_index, const float* ref_to_new, const tf_camera* cam);
int tf_pre_refine_newframe(tf_map* m, int32_t keyframe_index, int32_t new_index, const float* new_to_ref, const tf_camera* cam);
int tf_pre_refine_depth_by_normal(tf_map* m, int32_t frame_index, const tf_camera* cam);
/* rgb: Frame::rgb (host, H x W x 3).  Fills the key-frame's colour-valid, quality and RGBA planes and keeps the rgb
 * for the atlas (the slot is pinned like after tf_upload_keyframe_rgb). */
int tf_pre_color_quality(tf_map* m, int32_t frame_index, const uint8_t* rgb, const tf_camera* cam);
/* Read-back for the tracker and for parity (blocking; any pointer may be NULL): refined depth, normal map
 * (3 planes: x | y | z), refinement weights, colour-valid flags, quality. */
int tf_pre_download(tf_map* m, int32_t frame_index, float* depth, float* normal, float* weight, uint8_t* color_valid,
                    float* quality);

/* ---- misc ---------------------------------------------------------------------------- */
int tf_sync(tf_map* m);
/* Blocks until the uploads of one stored frame have completed (its page-locked source buffers may
 * then be reused); cheaper than tf_sync.  TF_ERR_NOT_FOUND if the frame is not in the store. */
int tf_wait_upload(tf_map* m, int32_t frame_index);
/* counters since creation: kernels launched by this library, bytes moved each way */
typedef struct {
  int64_t kernel_launches;
  int64_t h2d_bytes, d2h_bytes;
  int64_t frames_integrated;
  int64_t voxel_updates;
  int64_t pool_capacity, pool_used;
} tf_counters;
int tf_get_counters(tf_map* m, tf_counters* out);
/* CUDA stream of the map as a cudaStream_t cast to void* (for event timing by the caller) */
void* tf_stream(tf_map* m);
/* cudaStream_t of the frame uploads: a collective that must follow an upload (frame broadcast to the
 * other ranks) can be queued on it instead of waiting on the host. */
void* tf_copy_stream(tf_map* m);
/* device time (ms) spent in the integrate kernel since the last call with reset != 0;
 * measured with CUDA events on the map's stream when enabled via tf_set_profiling. */
int tf_set_profiling(tf_map* m, int level /* 0 off, 1 integrate kernel, 2 every pipeline stage */);
/* device time (ms) per stage of the fused pipeline since the last reset, level 2 only:
 * [bbox, cull, -, alloc, integrate (+ fused finalize), -] */
int tf_get_stage_times(tf_map* m, int reset, double* ms6);
int tf_get_kernel_time(tf_map* m, int reset, double* integrate_ms, int64_t* integrate_launches,
                       double* integrate_bytes);

/* Test hook: runs both pixel-projection paths of integrate_kernel (tf_device.cuh:
 * project_safe / project_exact) on n caller-provided (c, cz) pairs; accepted[i] != 0 where the pair
 * lies in the operand range the kernel's per-chunk test guarantees for project_safe. */
int tf_debug_project(tf_map* m, const float* c, const float* cz, int64_t n, float f, float ch,
                     int32_t* u_fast, int32_t* u_exact, uint8_t* accepted);
/* Test hook: the running average's quotient as integrate_kernel forms it (div.rn's fast-path
 * sequence inline, IEEE division out of line where its result is not kept) and __fdiv_rn, on n
 * caller-provided (numerator, divisor) pairs; accepted[i] != 0 where the inline result was kept. */
int tf_debug_divide(tf_map* m, const float* num, const float* den, int64_t n, float* q_kernel,
                    float* q_ieee, uint8_t* accepted);

#ifdef __cplusplus
}
#endif
#endif /* TEXFUSION_B200_H */
