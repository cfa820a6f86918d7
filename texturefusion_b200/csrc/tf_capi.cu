// tf_capi.cu — host runtime + C ABI (include/texfusion.h) of libtexfusion_b200.so.
//
// One tf_map = one GPU's share of the volumetric map: a device-resident open-addressing
// hash (chunk id -> slot), an 8 KiB-per-chunk slab pool, a frame store (depth / RGBA /
// quality / rgb planes kept in HBM), per-frame scratch lists and the texture atlas.
// Every stage of a frame runs as a chain of kernels on one stream with device-side work
// counts; the host synchronises once per call, when it needs the results.
#ifdef TF_TIMELINE
#include <chrono>
#endif
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/texfusion.h"
#include "tf_host_math.h"
#include "tf_kernels.cuh"
#include "tf_mesh.cuh"
#include "tf_pre.cuh"

using namespace tfb;

namespace {

thread_local std::string g_create_error;

struct FrameSlot {
  int frame_index = -1;
  bool has_rgba = false, has_quality = false, has_rgb = false;
  uint64_t last_use = 0;
  float* depth = nullptr;
  uchar4* rgba = nullptr;
  float* quality = nullptr;
  unsigned char* rgb = nullptr;
  unsigned char* valid = nullptr;
  cudaEvent_t ready = nullptr;  // recorded on the copy stream after the slot's uploads
  bool pending = false;         // uploads in flight: consumers on the compute stream wait for `ready`
  uint64_t read_ticket = 0;     // ticket of the last queued compute work that reads the slot (0: none since its upload)
  bool pinned = false;          // key-frame rgb kept for the atlas: not evictable until tf_release_frame
};

// ---- NCCL, bound at run time --------------------------------------------------------------------
// The one per-frame collective of the sharded path is the frame broadcast.  The library does not
// link against NCCL (single-GPU users need none): tf_comm_init dlopens libnccl.so.2 — in a process
// that already carries one (e.g. the copy bundled with PyTorch) that is the copy it gets — and
// binds the five entry points it uses.  Types restated from nccl.h (ABI-stable since NCCL 2.0).
typedef struct ncclComm* nccl_comm_t;
struct nccl_unique_id { char internal[128]; };
constexpr int kNcclUint8 = 1;
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(nccl_unique_id*) = nullptr;
  int (*CommInitRank)(nccl_comm_t*, int, nccl_unique_id, int) = nullptr;
  int (*CommDestroy)(nccl_comm_t) = nullptr;
  int (*Broadcast)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  std::string err;
};
NcclApi* nccl_api() {
  static NcclApi api;
  if (api.lib || !api.err.empty()) return &api;
  const char* names[] = {getenv("TEXFUSION_B200_NCCL"), "libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    if (!n) continue;
    api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (api.lib) break;
  }
  if (!api.lib) { api.err = std::string("cannot load NCCL: ") + dlerror(); return &api; }
  api.GetUniqueId = (int (*)(nccl_unique_id*))dlsym(api.lib, "ncclGetUniqueId");
  api.CommInitRank = (int (*)(nccl_comm_t*, int, nccl_unique_id, int))dlsym(api.lib, "ncclCommInitRank");
  api.CommDestroy = (int (*)(nccl_comm_t))dlsym(api.lib, "ncclCommDestroy");
  api.Broadcast = (int (*)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t))dlsym(api.lib, "ncclBroadcast");
  api.GetErrorString = (const char* (*)(int))dlsym(api.lib, "ncclGetErrorString");
  if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.Broadcast || !api.GetErrorString) {
    api.err = "libnccl lacks an expected symbol";
    api.lib = nullptr;
  }
  return &api;
}

struct EventPair {
  cudaEvent_t a, b;
  double bytes;
  int stage;  // 0 bbox, 1 cull, 2 (unused), 3 alloc, 4 integrate (+ fused finalize), 5 (unused)
};
constexpr int kStages = 6;

}  // namespace

// The fused per-frame chain (bbox -> cull+alloc -> integrate+finalize) as an instantiated CUDA
// graph: one graph launch instead of three kernel launches; the kernel arguments of the frame are
// patched in with cudaGraphExecKernelNodeSetParams.
struct FrameGraph {
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  cudaGraphNode_t node[4] = {nullptr, nullptr, nullptr, nullptr};  // bbox, cull, integrate, export
  cudaKernelNodeParams kp[4] = {};
  bool failed = false;  // capture / instantiation not possible: use plain launches
};

struct tf_map {
  tf_config cfg;
  int W = 0, H = 0, npix = 0;
  cudaStream_t stream = nullptr;       // all kernels and read-backs
  cudaStream_t copy_stream = nullptr;  // frame uploads: the next frame's H2D copy overlaps the current frame's kernels
  cudaEvent_t reuse_ev = nullptr;      // orders an overwriting upload behind the kernels that read the slot
  uint64_t compute_ticket = 0, compute_done = 0;  // see guard_overwrite
  cudaEvent_t patch_copy_done = nullptr;  // tf_atlas_update: the descriptor copy has left the page-locked arena
  std::string err;
  unsigned char* slab = nullptr;        // frame store: per slot [depth | rgba | quality | rgb | valid], contiguous so
  size_t slot_stride = 0;               // that one broadcast moves depth (+ rgba + quality) of a frame
  nccl_comm_t comm = nullptr;           // tf_comm_init
  int grid_bbox = 0;
  int sm_count = 0, grid = 0, grid_cull = 0, grid_integrate = 0, grid_integrate_c = 0;

  MapDev md{};
  FrameState* fs = nullptr;
  int hash_cap = 0;

  // per-frame scratch
  int cand_cap = 0, list_cap = 0;
  CullBuffers cb{};  // culling scratch + the frame's chunk list (device arrays, see tf_kernels.cuh)
  float* partial = nullptr;
  unsigned* list_upd = nullptr;
  float* list_q = nullptr;
  int setup_frames = 0;

  // mapped pinned result + output staging
  FrameResultHost* res_h = nullptr;
  FrameResultHost* res_d = nullptr;
  int3* out_ids_h = nullptr;
  unsigned char* out_new_h = nullptr;
  unsigned char* out_upd_h = nullptr;
  float* out_q_h = nullptr;
  // tf_integrate_batch: result arena in mapped page-locked memory (allocated by the first batch)
  int3* arena_ids = nullptr;
  float* arena_q = nullptr;
  unsigned char* arena_upd = nullptr;
  int4* batch_rec = nullptr;
  int3* ids_stage = nullptr;     // page-locked staging of the de-integration lists (pageable sources would
  int64_t ids_stage_used = 0;    // make cudaMemcpyAsync wait for the stream, i.e. serialise host and device)
  int3* st_ids = nullptr;  // device staging of the ordered per-chunk outputs (export_kernel)
  unsigned char* st_new = nullptr;
  unsigned char* st_upd = nullptr;
  float* st_q = nullptr;
  int3* out_ids_d = nullptr;
  unsigned char* out_new_d = nullptr;
  unsigned char* out_upd_d = nullptr;
  float* out_q_d = nullptr;
  unsigned* upd_stage_h = nullptr;  // pinned, list_cap
  float* q_stage_h = nullptr;

  // frame store
  std::vector<FrameSlot> slots;
  std::unordered_map<int, int> frame_to_slot;
  uint64_t use_clock = 0;

  // download staging
  int dl_cap = 0;
  float* dl_sdf = nullptr;
  float* dl_w = nullptr;
  uint2* dl_col = nullptr;
  int* count_d = nullptr;

  // atlas
  unsigned char* atlas = nullptr;
  uint64_t loc_next = 0;
  int patch_w = 0, patch_h = 0;
  std::unordered_map<unsigned long long, uint64_t> patches;

  // grow-only device / page-locked arenas of the entry points that move variable-sized data
  // (meshing, texcoords, chunk listing): no cudaMalloc / cudaFree in steady state
  struct Arena { void* p = nullptr; size_t cap = 0; bool host = false; };
  Arena ar_mesh_ids, ar_mesh_counts, ar_mesh_off, ar_mesh_v, ar_mesh_n, ar_mesh_c, ar_mesh_i, ar_mesh_off_h;
  Arena ar_tc_off, ar_tc_v, ar_tc_c, ar_tc_tc, ar_tc_col, ar_tc_res, ar_list, ar_patch_h, ar_patch_d;

  // frame pre-processing (tf_pre_*): the side planes of the few frames being prepared (normal map, refinement
  // weights; least-recently-used replacement) and the scratch of the in-place key-frame refinement
  struct PreEntry { int frame_index = -1; float* normal = nullptr; float* weight = nullptr; bool has_normal = false; uint64_t last_use = 0; };
  PreEntry pre[4];
  float* pre_snap = nullptr;     // snapshot of the key-frame's depth / staging of a raw 16-bit depth image
  BilateralState* pre_bil = nullptr;
  int* pre_queue = nullptr;      // [npix] + counter + last-block ticket
  unsigned char* pre_waiting = nullptr;

  // host mirrors
  int64_t n_live = 0;
  int pool_next = 0;
  int parity = 0;  // which set of bounding-box accumulators the current frame uses
  tf_counters counters{};

  // a fused frame that has been launched but not collected (tf_integrate_frame_begin / _end)
  struct PendingFrame {
    bool active = false, graph = false;
    unsigned seq = 0;
    int n_frames = 0, ocap = 0;
    bool color[kMaxGroupFrames] = {};
    tf_chunk_id* ids_out = nullptr;
    uint8_t *new_out = nullptr, *upd_out = nullptr;
    float* q_out = nullptr;
    bool direct[4] = {false, false, false, false};
    uint64_t ticket = 0;
    int n_export_blocks = 0;  // blocks of export_kernel whose stamps the host waits for (0: no lists)
  } pend;

  // CUDA graphs of the fused per-frame chain, one per (colour, group size); see fused_group
  FrameGraph graphs[2][2][kMaxGroupFrames + 1];  // [lists exported][colour][group size]
  unsigned seq = 0;  // completion stamp of the last fused frame (FrameResultHost::seq)

  // profiling of the integrate kernel
  int prof = 0;  // 1: integrate kernel only, 2: every stage of the fused pipeline
  std::vector<EventPair> ev_pool, ev_pending;
  double prof_ms = 0, prof_bytes = 0;
  int64_t prof_launches = 0;
  double stage_ms[kStages] = {0, 0, 0, 0, 0, 0};
};

namespace {

#define CUDA_OK(m, call)                                                                       \
  do {                                                                                         \
    cudaError_t e_ = (call);                                                                   \
    if (e_ != cudaSuccess) {                                                                   \
      (m)->err = std::string(#call) + ": " + cudaGetErrorString(e_);                           \
      return TF_ERR_CUDA;                                                                      \
    }                                                                                          \
  } while (0)

// grows (never shrinks) a device or page-locked host arena; the stream is drained before a buffer
// that queued work may still use is released
int ensure_arena_bytes(tf_map* m, tf_map::Arena& a, size_t need, bool host = false) {
  if (need <= a.cap) return TF_OK;
  cudaStreamSynchronize(m->stream);
  if (a.p) { if (a.host) cudaFreeHost(a.p); else cudaFree(a.p); }
  a.p = nullptr;
  a.cap = 0;
  const size_t cap = std::max<size_t>(need + need / 2, 4096);
  const cudaError_t e = host ? cudaHostAlloc(&a.p, cap, cudaHostAllocDefault) : cudaMalloc(&a.p, cap);
  if (e != cudaSuccess) { m->err = std::string("arena allocation: ") + cudaGetErrorString(e); return TF_ERR_CUDA; }
  a.cap = cap;
  a.host = host;
  return TF_OK;
}

int fail(tf_map* m, int code, const std::string& msg) {
  if (m) m->err = msg;
  return code;
}

template <class T>
cudaError_t dmalloc(T** p, size_t n) {
  return cudaMalloc((void**)p, n * sizeof(T));
}

#ifdef TF_TIMELINE
static double g_host_t[16];
#define HT(k) (g_host_t[k] = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count())
#else
#define HT(k) ((void)0)
#endif

// Launch with programmatic stream serialization: the kernel may start while its predecessor on
// the stream is still running and synchronises with it through griddepcontrol.wait.
template <class... KArgs, class... Args>
void launch_pdl(void (*kernel)(KArgs...), int grid, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  // on by default; TEXFUSION_B200_PDL=0 falls back to plain stream order
  static const bool use_pdl = [] { const char* e = getenv("TEXFUSION_B200_PDL"); return !(e && e[0] == '0'); }();
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = use_pdl ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

int check_kernel(tf_map* m, const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(m, TF_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
  m->counters.kernel_launches++;
  return TF_OK;
}

// Every entry point that touches the device first makes the map's device current: a process may
// hold maps on several GPUs (or another library may have switched devices in between).
inline void use_device(const tf_map* m) {
  int cur = -1;
  if (cudaGetDevice(&cur) != cudaSuccess || cur != m->cfg.device) cudaSetDevice(m->cfg.device);
}

// (near_plane >= 0: integrate_kernel relies on masked-out gathers (depth 0) failing `depth > near`)
bool cam_ok(const tf_map* m, const tf_camera* c) { return c && c->width == m->W && c->height == m->H && c->near_plane >= 0.0f; }

// Slot of a stored frame, or -1.  for_compute: work is about to be queued on the compute stream
// that reads the slot, so make that stream wait for the slot's uploads.
int find_slot(tf_map* m, int32_t frame_index, bool for_compute = true) {
  auto it = m->frame_to_slot.find(frame_index);
  if (it == m->frame_to_slot.end()) return -1;
  FrameSlot& fsl = m->slots[it->second];
  fsl.last_use = ++m->use_clock;
  if (for_compute) {
    fsl.read_ticket = ++m->compute_ticket;
    if (fsl.pending) {
      cudaStreamWaitEvent(m->stream, fsl.ready, 0);
      fsl.pending = false;
    }
  }
  return it->second;
}

// Before the copy stream overwrites a slot that queued kernels may still be reading (a re-upload
// of the same index, or a slot taken over from an evicted frame): order the copy behind them.
// (Work up to ticket compute_done is known to have finished — the host has collected its results —
//  so in the streaming loop, where frame i+1 is uploaded while frame i runs, the copy is not held back.)
void guard_overwrite(tf_map* m, FrameSlot& fsl) {
  if (fsl.read_ticket <= m->compute_done) return;
  cudaEventRecord(m->reuse_ev, m->stream);
  cudaStreamWaitEvent(m->copy_stream, m->reuse_ev, 0);
  fsl.read_ticket = 0;
}

// Slot for frame_index: existing, free, or the least-recently-used one that is not pinned
// (-1: every slot holds a pinned key-frame).
int acquire_slot(tf_map* m, int32_t frame_index) {
  int s = find_slot(m, frame_index, false);
  if (s >= 0) return s;
  int best = -1;
  for (int i = 0; i < (int)m->slots.size(); i++) {
    if (m->slots[i].frame_index < 0) { best = i; break; }
    if (m->slots[i].pinned) continue;
    if (best < 0 || m->slots[i].last_use < m->slots[best].last_use) best = i;
  }
  if (best < 0) {
    m->err = "frame store full: every slot holds a key-frame pinned by tf_upload_keyframe_rgb (raise tf_config.max_frames "
             "or tf_release_frame)";
    return -1;
  }
  FrameSlot& fsl = m->slots[best];
  if (fsl.frame_index >= 0) m->frame_to_slot.erase(fsl.frame_index);
  fsl.frame_index = frame_index;
  fsl.has_rgba = fsl.has_quality = fsl.has_rgb = false;
  fsl.pinned = false;
  fsl.last_use = ++m->use_clock;
  m->frame_to_slot[frame_index] = best;
  return best;
}

int ensure_color_planes(tf_map* m, FrameSlot& s) {
  if (!s.rgba) return fail(m, TF_ERR_INVALID, "map created with use_color = 0: no colour planes in the frame store");
  return TF_OK;
}

int dev_error_to_code(tf_map* m, int e) {
  if (!e) return TF_OK;
  cudaMemsetAsync(&m->fs->error, 0, sizeof(int), m->stream);
  if (e & kErrMissing) return fail(m, TF_ERR_NOT_FOUND, "chunk id not in the map");
  if (e & kErrDepth) return fail(m, TF_ERR_INVALID, "depth image contains an infinity or a NaN (the contract is finite depths, 0 = no measurement)");
  if (e & kErrCoord) return fail(m, TF_ERR_INVALID, "chunk coordinates outside +-2^20");
  if (e & kErrPool) return fail(m, TF_ERR_CAPACITY, "chunk pool exhausted (raise tf_config.max_chunks)");
  if (e & kErrList) return fail(m, TF_ERR_CAPACITY, "frame chunk list exceeds the list capacity");
  return fail(m, TF_ERR_CAPACITY, "culling candidate grid exceeds capacity");
}

void absorb_result(tf_map* m) {
  m->n_live = m->res_h->n_live;
  m->pool_next = m->res_h->pool_next;
  m->counters.pool_used = m->n_live;
}

// ---- profiling helpers ---------------------------------------------------------------------
void prof_begin(tf_map* m, EventPair& ep) {
  if (m->ev_pool.empty()) {
    cudaEventCreate(&ep.a);
    cudaEventCreate(&ep.b);
  } else {
    ep = m->ev_pool.back();
    m->ev_pool.pop_back();
  }
  cudaEventRecord(ep.a, m->stream);
}
void prof_end(tf_map* m, EventPair& ep, int stage = 4) {
  cudaEventRecord(ep.b, m->stream);
  ep.bytes = -1;
  ep.stage = stage;
  m->ev_pending.push_back(ep);
}
// after a stream sync: fold finished event pairs; `bytes` = algorithmic bytes of the launch
void prof_collect(tf_map* m, double bytes_last) {
  for (size_t i = 0; i < m->ev_pending.size(); i++) {
    EventPair& ep = m->ev_pending[i];
    float ms = 0;
    if (cudaEventElapsedTime(&ms, ep.a, ep.b) == cudaSuccess) {
      m->stage_ms[ep.stage] += ms;
      if (ep.stage == 4) {
        m->prof_ms += ms;
        m->prof_launches++;
        m->prof_bytes += (ep.bytes >= 0 ? ep.bytes : bytes_last);
      }
    }
    m->ev_pool.push_back(ep);
  }
  m->ev_pending.clear();
}

// algorithmic bytes of one integrate launch: 16 B per visited voxel for a
// depth-only frame, 32 B with colour, plus the image planes once per frame.
double algorithmic_bytes(const tf_map* m, int64_t n_chunks, const bool* color, int n_frames) {
  double b = 0;
  for (int f = 0; f < n_frames; f++)
    b += (double)n_chunks * 512.0 * (color[f] ? 32.0 : 16.0) + (double)m->npix * (color[f] ? 12.0 : 4.0);
  return b;
}

// ---- pipeline stages -------------------------------------------------------------------------

int ensure_setup(tf_map* m, int n_frames) {
  if (n_frames <= m->setup_frames) return TF_OK;
  cudaStreamSynchronize(m->stream);
  cudaFree(m->cb.list_setup);
  m->cb.list_setup = nullptr;
  m->setup_frames = 0;
  CUDA_OK(m, dmalloc(&m->cb.list_setup, (size_t)m->list_cap * n_frames * kSetupStride));
  m->setup_frames = n_frames;
  return TF_OK;
}

// bbox + culling.  fused: the culling kernel also resolves / creates the chunks and builds the
// frame's list for the frames of `gp` (else: hit queue only, for alloc_kernel).
int launch_cull(tf_map* m, const CullParams& cp, const GroupParams& gp, const float* depth, bool fused, bool want_order) {
  if (int rc = ensure_setup(m, std::max(1, gp.n_frames))) return rc;
  EventPair ep;
  const bool st = m->prof >= 2;
  if (st) prof_begin(m, ep);
  m->parity ^= 1;
  HT(2);
  bbox_kernel<<<m->grid_bbox, kThreads, 0, m->stream>>>(cp, depth, m->fs, m->parity);
  HT(3);
  if (st) { prof_end(m, ep, 0); prof_begin(m, ep); }
  if (int rc = check_kernel(m, "bbox_kernel")) return rc;
  if (fused)
    launch_pdl(cull_kernel<true>, m->grid_cull, 0, m->stream, cp, gp, m->md, depth, m->fs, m->cb, m->cfg.n_ranks,
               m->cfg.rank, m->parity, want_order ? 1 : 0);
  else
    launch_pdl(cull_kernel<false>, m->grid_cull, 0, m->stream, cp, gp, m->md, depth, m->fs, m->cb, m->cfg.n_ranks,
               m->cfg.rank, m->parity, 1);
  if (st) prof_end(m, ep, 1);
  HT(4);
  return check_kernel(m, "cull_kernel");
}

// The four instantiations of integrate_kernel: colour x single-frame.
using IntegrateFn = void (*)(GroupParams, MapDev, const int*, const int*, const float*, const int*, int, unsigned*, float*,
                             FusedFinalize);
IntegrateFn integrate_fn(bool color, int n_frames) {
  const bool single = n_frames == 1;
  return color ? (single ? integrate_kernel<true, true> : integrate_kernel<true, false>)
               : (single ? integrate_kernel<false, true> : integrate_kernel<false, false>);
}
void launch_integrate_kernel(tf_map* m, const GroupParams& gp, bool color, const int* n_dev, int n_host, const FusedFinalize& ff) {
  launch_pdl(integrate_fn(color, gp.n_frames), color ? m->grid_integrate_c : m->grid_integrate,
             integrate_smem_bytes(gp.n_frames, color), m->stream, gp, m->md, (const int*)m->cb.list_slots,
             (const int*)m->cb.list_hpos, (const float*)m->cb.list_setup, n_dev, n_host, m->list_upd, m->list_q, ff);
}

int launch_integrate(tf_map* m, const GroupParams& gp, const int* n_dev, int n_host, double bytes,
                     const FusedFinalize* fused = nullptr) {
  FusedFinalize ff{};
  if (fused) ff = *fused;
  EventPair ep;
  if (m->prof) prof_begin(m, ep);
  bool any_color = false;
  for (int f = 0; f < gp.n_frames; f++) any_color |= gp.f[f].rgba != nullptr;
  launch_integrate_kernel(m, gp, any_color, n_dev, n_host, ff);
  if (m->prof) {
    prof_end(m, ep);
    m->ev_pending.back().bytes = bytes;
  }
  return check_kernel(m, "integrate_kernel");
}

int build_group(tf_map* m, const tf_group_frame* frames, int n_frames, const tf_camera* cam, GroupParams& gp,
                bool* color_flags) {
  if (n_frames < 1 || n_frames > kMaxGroupFrames) return fail(m, TF_ERR_INVALID, "group size must be 1..8");
  make_group_consts(m->cfg.voxel_res, m->cfg.trunc, m->cfg.dot3_order, gp);
  gp.n_frames = n_frames;
  for (int f = 0; f < n_frames; f++) {
    const int s = find_slot(m, frames[f].frame_index);
    if (s < 0) return fail(m, TF_ERR_NOT_FOUND, "frame_index not in the frame store");
    const FrameSlot& fsl = m->slots[s];
    const bool color = frames[f].use_color != 0;
    if (color && !m->cfg.use_color) return fail(m, TF_ERR_INVALID, "map created with use_color = 0");
    if (color && !fsl.has_rgba) return fail(m, TF_ERR_NOT_FOUND, "frame has no colour plane");
    make_frame_dev(frames[f].pose, *cam, m->cfg.voxel_res, frames[f].flag, fsl.depth, color ? fsl.rgba : nullptr,
                   (color && fsl.has_quality) ? fsl.quality : nullptr, gp.f[f]);
    color_flags[f] = color;
  }
  return TF_OK;
}

int upload_ids(tf_map* m, const tf_chunk_id* ids, int64_t n) {
  if (n > m->list_cap) return fail(m, TF_ERR_CAPACITY, "chunk list exceeds the list capacity");
  CUDA_OK(m, cudaMemcpyAsync(m->cb.list_ids, ids, (size_t)n * sizeof(int3), cudaMemcpyHostToDevice, m->stream));
  m->counters.h2d_bytes += n * (int64_t)sizeof(int3);
  return TF_OK;
}

// ids (host) -> list_slots (device); fails with TF_ERR_NOT_FOUND before anything is modified.
int lookup_ids(tf_map* m, const tf_chunk_id* ids, int64_t n, bool must_exist, const GroupParams* gp = nullptr) {
  static const GroupParams kNoFrames{};  // n_frames == 0
  if (int rc = ensure_setup(m, gp ? gp->n_frames : 1)) return rc;
  if (int rc = upload_ids(m, ids, n)) return rc;
  const int grid = (int)std::min<int64_t>(m->grid, (n + kThreads - 1) / kThreads);
  lookup_kernel<<<std::max(grid, 1), kThreads, 0, m->stream>>>(gp ? *gp : kNoFrames, m->md, m->fs, m->cb.list_ids, (int)n,
                                                               m->cb.list_slots, m->cb.list_hpos, m->cb.list_setup);
  if (int rc = check_kernel(m, "lookup_kernel")) return rc;
  publish_kernel<<<1, 1, 0, m->stream>>>(m->fs, m->res_d);
  if (int rc = check_kernel(m, "publish_kernel")) return rc;
  CUDA_OK(m, cudaStreamSynchronize(m->stream));
  const int e = m->res_h->error;
  if (e) {
    cudaMemsetAsync(&m->fs->error, 0, sizeof(int), m->stream);
    if (must_exist || (e & ~kErrMissing)) return dev_error_to_code(m, e);
  }
  return TF_OK;
}

}  // namespace

// ============================================================================================
extern "C" {

const char* tf_last_error(const tf_map* m) { return m ? m->err.c_str() : g_create_error.c_str(); }

void* tf_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) return nullptr;
  return p;
}
void tf_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

static void destroy_graphs(tf_map* m) {
  for (auto& plane : m->graphs)
    for (auto& row : plane)
      for (auto& fg : row) {
        if (fg.exec) cudaGraphExecDestroy(fg.exec);
        if (fg.graph) cudaGraphDestroy(fg.graph);
        fg = FrameGraph{};
      }
}

void tf_destroy(tf_map* m) {
  if (!m) return;
  cudaSetDevice(m->cfg.device);
  if (m->stream) cudaStreamSynchronize(m->stream);
  if (m->copy_stream) cudaStreamSynchronize(m->copy_stream);
  if (m->comm) nccl_api()->CommDestroy(m->comm);
  destroy_graphs(m);
  for (auto& sl : m->slots)
    if (sl.ready) cudaEventDestroy(sl.ready);
  if (m->copy_stream) cudaStreamDestroy(m->copy_stream);
  if (m->reuse_ev) cudaEventDestroy(m->reuse_ev);
  if (m->patch_copy_done) cudaEventDestroy(m->patch_copy_done);
  cudaFree(m->md.table); cudaFree(m->md.pool); cudaFree(m->md.slot_id);
  cudaFree(m->md.slot_flags); cudaFree(m->md.free_stack); cudaFree(m->fs);
  cudaFree(m->cb.mask32); cudaFree(m->cb.local_off); cudaFree(m->cb.hit_items); cudaFree(m->cb.hit_count);
  cudaFree(m->cb.word_base); cudaFree(m->cb.list_cb); cudaFree(m->cb.list_hpos);
  cudaFree(m->partial); cudaFree(m->cb.list_ids); cudaFree(m->cb.list_slots); cudaFree(m->cb.list_new);
  cudaFree(m->list_upd); cudaFree(m->list_q); cudaFree(m->cb.list_setup); cudaFree(m->dl_sdf); cudaFree(m->dl_w); cudaFree(m->dl_col);
  cudaFree(m->count_d); cudaFree(m->atlas);
  cudaFreeHost(m->res_h); cudaFreeHost(m->out_ids_h); cudaFreeHost(m->out_new_h); cudaFreeHost(m->out_upd_h);
  cudaFreeHost(m->out_q_h); cudaFreeHost(m->upd_stage_h); cudaFreeHost(m->q_stage_h);
  cudaFree(m->slab);
  for (auto& e : m->pre) { cudaFree(e.normal); cudaFree(e.weight); }
  cudaFree(m->pre_snap); cudaFree(m->pre_bil); cudaFree(m->pre_queue); cudaFree(m->pre_waiting);
  for (tf_map::Arena* a : {&m->ar_mesh_ids, &m->ar_mesh_counts, &m->ar_mesh_off, &m->ar_mesh_v, &m->ar_mesh_n, &m->ar_mesh_c,
                           &m->ar_mesh_i, &m->ar_mesh_off_h, &m->ar_tc_off, &m->ar_tc_v, &m->ar_tc_c, &m->ar_tc_tc,
                           &m->ar_tc_col, &m->ar_tc_res, &m->ar_list, &m->ar_patch_h, &m->ar_patch_d})
    if (a->p) { if (a->host) cudaFreeHost(a->p); else cudaFree(a->p); }
  cudaFree(m->st_ids); cudaFree(m->st_new); cudaFree(m->st_upd); cudaFree(m->st_q);
  cudaFreeHost(m->arena_ids); cudaFreeHost(m->arena_q); cudaFreeHost(m->arena_upd); cudaFreeHost(m->batch_rec);
  cudaFreeHost(m->ids_stage);
  for (auto& ep : m->ev_pool) { cudaEventDestroy(ep.a); cudaEventDestroy(ep.b); }
  for (auto& ep : m->ev_pending) { cudaEventDestroy(ep.a); cudaEventDestroy(ep.b); }
  if (m->stream) cudaStreamDestroy(m->stream);
  delete m;
}

static int reset_device_state(tf_map* m) {
  CUDA_OK(m, cudaMemsetAsync(m->md.table, 0xFF, (size_t)m->hash_cap * sizeof(HashEntry), m->stream));
  CUDA_OK(m, cudaMemsetAsync(m->md.slot_flags, 0, (size_t)m->md.max_chunks, m->stream));
  FrameState init;
  memset(&init, 0, sizeof(init));
  for (int p = 0; p < 2; p++)
    for (int k = 0; k < 3; k++) {  // seeds of findCubeCornerByMat: +-1e8 (ordered-int encoding of tf_device.cuh)
      const float lo = 1e8f, hi = -1e8f;
      int il, ih;
      memcpy(&il, &lo, 4);
      memcpy(&ih, &hi, 4);
      init.bbox_enc[p][k] = il >= 0 ? il : il ^ 0x7FFFFFFF;
      init.bbox_enc[p][3 + k] = ih >= 0 ? ih : ih ^ 0x7FFFFFFF;
    }
  CUDA_OK(m, cudaMemcpyAsync(m->fs, &init, sizeof(FrameState), cudaMemcpyHostToDevice, m->stream));
  CUDA_OK(m, cudaStreamSynchronize(m->stream));
  m->parity = 0;
  m->n_live = 0;
  m->pool_next = 0;
  return TF_OK;
}

int tf_create(tf_map** out, const tf_config* cfg) {
  if (!out || !cfg) { g_create_error = "null argument"; return TF_ERR_INVALID; }
  *out = nullptr;
  if (cfg->chunk_dim != 8) { g_create_error = "chunk_dim must be 8"; return TF_ERR_INVALID; }
  if (!(cfg->voxel_res > 0)) { g_create_error = "voxel_res must be positive"; return TF_ERR_INVALID; }
  if (cfg->n_ranks < 1 || cfg->rank < 0 || cfg->rank >= cfg->n_ranks) {
    g_create_error = "bad n_ranks/rank";
    return TF_ERR_INVALID;
  }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    g_create_error = std::string("no CUDA device (this library has no CPU fallback): ") + cudaGetErrorString(e);
    return TF_ERR_CUDA;
  }
  if (cfg->device < 0 || cfg->device >= ndev) { g_create_error = "bad device ordinal"; return TF_ERR_INVALID; }
  tf_map* m = new tf_map;
  m->cfg = *cfg;
  m->W = cfg->width > 0 ? cfg->width : 640;
  m->H = cfg->height > 0 ? cfg->height : 480;
  m->npix = m->W * m->H;
  if (m->W % 8 != 0) {  // the reference's bbox loop consumes 8 pixels per step (ChunkManager.h:326-328)
    g_create_error = "frame width must be a multiple of 8";
    delete m;
    return TF_ERR_INVALID;
  }
  if (m->W > 2048 || m->H > 2048) {  // (16-bit crop coordinates of the atlas descriptors, 32-bit pixel indices)
    g_create_error = "frame larger than 2048x2048";
    delete m;
    return TF_ERR_INVALID;
  }
  auto bail = [&](int rc) {
    g_create_error = m->err;
    tf_destroy(m);
    return rc;
  };
#define C_OK(call)                                                                   \
  do {                                                                               \
    cudaError_t e_ = (call);                                                         \
    if (e_ != cudaSuccess) {                                                         \
      m->err = std::string(#call) + ": " + cudaGetErrorString(e_);                   \
      return bail(TF_ERR_CUDA);                                                      \
    }                                                                                \
  } while (0)
  C_OK(cudaSetDevice(cfg->device));
  cudaDeviceProp prop;
  C_OK(cudaGetDeviceProperties(&prop, cfg->device));
  m->sm_count = prop.multiProcessorCount;
  C_OK(cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking));
  C_OK(cudaStreamCreateWithFlags(&m->copy_stream, cudaStreamNonBlocking));
  C_OK(cudaEventCreateWithFlags(&m->reuse_ev, cudaEventDisableTiming));
  int occ = 1;
  for (int color = 0; color < 2; color++)
    for (int nf : {1, kMaxGroupFrames})
      C_OK(cudaFuncSetAttribute(integrate_fn(color, nf), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)integrate_smem_bytes(nf, color)));
  int occ_c = 1;
  C_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, integrate_fn(false, 1), kThreads, integrate_smem_bytes(1)));
  C_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_c, integrate_fn(true, 1), kThreads, integrate_smem_bytes(1, true)));
  m->grid_integrate_c = m->sm_count * std::max(1, occ_c);
  m->grid = m->sm_count * 2;
  m->grid_bbox = std::max(m->grid, (m->npix / 4 + kThreads - 1) / kThreads);  // one float4 per thread
  m->grid_cull = m->sm_count * 4;
  m->grid_integrate = m->sm_count * std::max(1, occ);

  const int64_t max_chunks = cfg->max_chunks > 0 ? cfg->max_chunks : (int64_t)1 << 19;
  if (max_chunks > ((int64_t)1 << 27)) { m->err = "max_chunks too large"; return bail(TF_ERR_INVALID); }
  m->md.max_chunks = (int)max_chunks;
  int hc = 1024;
  while (hc < 2 * max_chunks) hc <<= 1;
  m->hash_cap = hc;
  m->md.hash_mask = (unsigned)hc - 1;
  m->md.n_ranks = cfg->n_ranks;
  m->md.rank = cfg->rank;
  C_OK(dmalloc(&m->md.table, (size_t)hc));
  C_OK(cudaMalloc((void**)&m->md.pool, (size_t)max_chunks * kChunkBytes));
  C_OK(dmalloc(&m->md.slot_id, (size_t)max_chunks));
  C_OK(dmalloc(&m->md.slot_flags, (size_t)max_chunks));
  C_OK(dmalloc(&m->md.free_stack, (size_t)max_chunks + 4));  // (+4: cull_kernel stages the top in 16-byte pieces)
  C_OK(dmalloc(&m->fs, 1));

  m->cand_cap = 1 << 22;  // coarse candidates per frame (4^3-chunk blocks at <= 10 mm voxels)
  m->list_cap = 1 << 19;
  C_OK(dmalloc(&m->cb.mask32, (size_t)m->cand_cap * 2));
  C_OK(dmalloc(&m->cb.local_off, (size_t)m->cand_cap));
  C_OK(dmalloc(&m->cb.hit_items, (size_t)m->cand_cap * 2));
  C_OK(dmalloc(&m->cb.hit_count, (size_t)m->cand_cap * 2 + 64));
  C_OK(dmalloc(&m->cb.word_base, (size_t)m->cand_cap / 32));
  C_OK(dmalloc(&m->cb.list_cb, (size_t)m->list_cap));
  C_OK(dmalloc(&m->cb.list_hpos, (size_t)m->list_cap));
  m->cb.list_cap = m->list_cap;
  m->cb.cand_cap = m->cand_cap;
  C_OK(dmalloc(&m->partial, (size_t)m->grid * 6));
  C_OK(dmalloc(&m->cb.list_ids, (size_t)m->list_cap));
  C_OK(dmalloc(&m->cb.list_slots, (size_t)m->list_cap));
  C_OK(dmalloc(&m->cb.list_new, (size_t)m->list_cap));
  C_OK(dmalloc(&m->list_upd, (size_t)m->list_cap));
  C_OK(dmalloc(&m->list_q, (size_t)m->list_cap));
  C_OK(dmalloc(&m->count_d, 1));

  const unsigned hflags = cudaHostAllocMapped;
  C_OK(cudaHostAlloc((void**)&m->res_h, sizeof(FrameResultHost), hflags));
  C_OK(cudaHostAlloc((void**)&m->out_ids_h, (size_t)m->list_cap * sizeof(int3), hflags));
  C_OK(cudaHostAlloc((void**)&m->out_new_h, (size_t)m->list_cap, hflags));
  C_OK(cudaHostAlloc((void**)&m->out_upd_h, (size_t)m->list_cap, hflags));
  C_OK(cudaHostAlloc((void**)&m->out_q_h, (size_t)m->list_cap * sizeof(float), hflags));
  C_OK(cudaHostAlloc((void**)&m->upd_stage_h, (size_t)m->list_cap * sizeof(unsigned), cudaHostAllocDefault));
  C_OK(cudaHostAlloc((void**)&m->q_stage_h, (size_t)m->list_cap * sizeof(float), cudaHostAllocDefault));
  memset(m->res_h, 0, sizeof(FrameResultHost));
  C_OK(cudaHostGetDevicePointer((void**)&m->res_d, m->res_h, 0));
  C_OK(dmalloc(&m->st_ids, (size_t)m->list_cap));
  C_OK(dmalloc(&m->st_new, (size_t)m->list_cap));
  C_OK(dmalloc(&m->st_upd, (size_t)m->list_cap));
  C_OK(dmalloc(&m->st_q, (size_t)m->list_cap));
  C_OK(cudaHostGetDevicePointer((void**)&m->out_ids_d, m->out_ids_h, 0));
  C_OK(cudaHostGetDevicePointer((void**)&m->out_new_d, m->out_new_h, 0));
  C_OK(cudaHostGetDevicePointer((void**)&m->out_upd_d, m->out_upd_h, 0));
  C_OK(cudaHostGetDevicePointer((void**)&m->out_q_d, m->out_q_h, 0));

  const int max_frames = cfg->max_frames > 0 ? cfg->max_frames : 32;
  // The frame store is allocated up front as one slab (4 B per pixel and slot, 16 B with colour):
  // cudaMalloc in the per-frame path would stall the stream for hundreds of microseconds.  A slot
  // is [depth | rgba | quality | rgb | valid]: the planes a broadcast moves are contiguous.
  m->slots.resize(max_frames);
  const size_t plane = (((size_t)m->npix * 4) + 15) & ~(size_t)15;  // keeps every plane 16-byte aligned (float4 loads)
  m->slot_stride = cfg->use_color ? plane * 3 + (((size_t)m->npix * 4 + 15) & ~(size_t)15) : plane;
  C_OK(cudaMalloc((void**)&m->slab, m->slot_stride * max_frames));
  for (int i = 0; i < max_frames; i++) {
    unsigned char* base = m->slab + (size_t)i * m->slot_stride;
    m->slots[i].depth = reinterpret_cast<float*>(base);
    if (cfg->use_color) {
      m->slots[i].rgba = reinterpret_cast<uchar4*>(base + plane);
      m->slots[i].quality = reinterpret_cast<float*>(base + 2 * plane);
      m->slots[i].rgb = base + 3 * plane;
      m->slots[i].valid = base + 3 * plane + (size_t)m->npix * 3;
    }
    C_OK(cudaEventCreateWithFlags(&m->slots[i].ready, cudaEventDisableTiming));
  }

  // Atlas::SetResolution (Structure/Atlas.h:62-65)
  m->patch_w = (int)std::floor(4800 * cfg->voxel_res);
  m->patch_h = (int)std::floor(3600 * cfg->voxel_res);
#undef C_OK
  if (int rc = reset_device_state(m)) return bail(rc);
  m->counters.pool_capacity = max_chunks;
  *out = m;
  return TF_OK;
}

int tf_reset(tf_map* m) {
  if (!m) return TF_ERR_INVALID;
  cudaSetDevice(m->cfg.device);
  return reset_device_state(m);
}

int tf_set_truncation(tf_map* m, const tf_truncation* t) {
  if (!m || !t) return fail(m, TF_ERR_INVALID, "tf_set_truncation: bad argument");
  m->cfg.trunc = *t;  // read when the next call builds its frame constants
  return TF_OK;
}

int tf_sync(tf_map* m) {
  if (!m) return TF_ERR_INVALID;
  use_device(m);
  CUDA_OK(m, cudaStreamSynchronize(m->copy_stream));
  CUDA_OK(m, cudaStreamSynchronize(m->stream));
  if (!m->pend.active) m->compute_done = m->compute_ticket;
  prof_collect(m, 0);
  return TF_OK;
}

int tf_wait_upload(tf_map* m, int32_t frame_index) {
  if (!m) return TF_ERR_INVALID;
  use_device(m);
  const int s = find_slot(m, frame_index, false);
  if (s < 0) return fail(m, TF_ERR_NOT_FOUND, "frame_index not in the frame store");
  // (the slot stays `pending` for the compute stream: waiting on a completed event costs nothing)
  for (;;) {  // polling the event returns earlier than cudaEventSynchronize's wake-up
    const cudaError_t e = cudaEventQuery(m->slots[s].ready);
    if (e == cudaSuccess) return TF_OK;
    if (e != cudaErrorNotReady) return fail(m, TF_ERR_CUDA, cudaGetErrorString(e));
  }
}

void* tf_copy_stream(tf_map* m) { return m ? (void*)m->copy_stream : nullptr; }

void* tf_stream(tf_map* m) { return m ? (void*)m->stream : nullptr; }

// ---- frame store ---------------------------------------------------------------------------

int tf_upload_frame(tf_map* m, int32_t frame_index, const float* depth, const uint8_t* rgba,
                    const float* quality) {
  if (!m || !depth || frame_index < 0) return fail(m, TF_ERR_INVALID, "tf_upload_frame: bad argument");
  use_device(m);
  const int s = acquire_slot(m, frame_index);
  if (s < 0) return TF_ERR_CAPACITY;
  FrameSlot& fsl = m->slots[s];
  guard_overwrite(m, fsl);
  // the planes of this upload replace the slot's contents: planes that are not passed are absent
  // afterwards (the reference reads NULL pointers as "no colour" / quality 0, Structure/Chisel.h:218-249)
  fsl.has_rgba = fsl.has_quality = fsl.has_rgb = false;
  for (auto& e : m->pre)  // a fresh frame: no normal map yet, refinement weights 0 (framePreprocess, BasicAPI.cpp:942-953)
    if (e.frame_index == frame_index) e.frame_index = -1;
  const size_t nb = (size_t)m->npix * 4;
  CUDA_OK(m, cudaMemcpyAsync(fsl.depth, depth, nb, cudaMemcpyHostToDevice, m->copy_stream));
  m->counters.h2d_bytes += nb;
  if (rgba || quality) {
    if (int rc = ensure_color_planes(m, fsl)) return rc;
  }
  if (rgba) {
    CUDA_OK(m, cudaMemcpyAsync(fsl.rgba, rgba, nb, cudaMemcpyHostToDevice, m->copy_stream));
    m->counters.h2d_bytes += nb;
    fsl.has_rgba = true;
  }
  if (quality) {
    CUDA_OK(m, cudaMemcpyAsync(fsl.quality, quality, nb, cudaMemcpyHostToDevice, m->copy_stream));
    m->counters.h2d_bytes += nb;
    fsl.has_quality = true;
  }
  CUDA_OK(m, cudaEventRecord(fsl.ready, m->copy_stream));
  fsl.pending = true;
  return TF_OK;
}

int tf_upload_keyframe_rgb(tf_map* m, int32_t frame_index, const uint8_t* rgb, const uint8_t* color_valid) {
  if (!m || !rgb) return fail(m, TF_ERR_INVALID, "tf_upload_keyframe_rgb: bad argument");
  use_device(m);
  const int s = find_slot(m, frame_index, false);
  if (s < 0) return fail(m, TF_ERR_NOT_FOUND, "frame_index not in the frame store (upload depth first)");
  FrameSlot& fsl = m->slots[s];
  if (int rc = ensure_color_planes(m, fsl)) return rc;
  guard_overwrite(m, fsl);
  fsl.pinned = true;  // Frame::rgb stays available to the atlas / texcoords until tf_release_frame
  CUDA_OK(m, cudaMemcpyAsync(fsl.rgb, rgb, (size_t)m->npix * 3, cudaMemcpyHostToDevice, m->copy_stream));
  m->counters.h2d_bytes += (int64_t)m->npix * 3;
  if (color_valid) {
    CUDA_OK(m, cudaMemcpyAsync(fsl.valid, color_valid, (size_t)m->npix, cudaMemcpyHostToDevice, m->copy_stream));
    m->counters.h2d_bytes += m->npix;
  }
  pack_rgba_kernel<<<m->grid, kThreads, 0, m->copy_stream>>>(fsl.rgb, color_valid ? fsl.valid : nullptr, fsl.rgba, m->npix);
  if (int rc = check_kernel(m, "pack_rgba_kernel")) return rc;
  fsl.has_rgb = true;
  fsl.has_rgba = true;
  CUDA_OK(m, cudaEventRecord(fsl.ready, m->copy_stream));
  fsl.pending = true;
  return TF_OK;
}

int tf_release_frame(tf_map* m, int32_t frame_index) {
  if (!m) return TF_ERR_INVALID;
  auto it = m->frame_to_slot.find(frame_index);
  if (it == m->frame_to_slot.end()) return fail(m, TF_ERR_NOT_FOUND, "frame_index not in the frame store");
  m->slots[it->second].frame_index = -1;
  m->slots[it->second].pinned = false;
  m->frame_to_slot.erase(it);
  return TF_OK;
}

int tf_frame_device_ptrs(tf_map* m, int32_t frame_index, int has_color, void** depth, void** rgba,
                         void** quality) {
  if (!m || frame_index < 0) return fail(m, TF_ERR_INVALID, "tf_frame_device_ptrs: bad argument");
  use_device(m);
  const int s = acquire_slot(m, frame_index);
  if (s < 0) return TF_ERR_CAPACITY;
  FrameSlot& fsl = m->slots[s];
  guard_overwrite(m, fsl);
  fsl.has_rgba = fsl.has_quality = fsl.has_rgb = false;
  if (has_color) {
    if (int rc = ensure_color_planes(m, fsl)) return rc;
    fsl.has_rgba = fsl.has_quality = true;
  }
  if (depth) *depth = fsl.depth;
  if (rgba) *rgba = has_color ? (void*)fsl.rgba : nullptr;
  if (quality) *quality = has_color ? (void*)fsl.quality : nullptr;
  return TF_OK;
}

// ---- multi-GPU: the frame broadcast --------------------------------------------------------------

int tf_comm_unique_id(uint8_t* id_out) {
  if (!id_out) return TF_ERR_INVALID;
  NcclApi* n = nccl_api();
  if (!n->lib) { g_create_error = n->err; return TF_ERR_CUDA; }
  nccl_unique_id id;
  const int rc = n->GetUniqueId(&id);
  if (rc != 0) { g_create_error = std::string("ncclGetUniqueId: ") + n->GetErrorString(rc); return TF_ERR_CUDA; }
  memcpy(id_out, id.internal, sizeof(id.internal));
  return TF_OK;
}

int tf_comm_init(tf_map* m, const uint8_t* id128) {
  if (!m || !id128) return fail(m, TF_ERR_INVALID, "tf_comm_init: bad argument");
  if (m->comm) return fail(m, TF_ERR_INVALID, "tf_comm_init: communicator already initialised");
  use_device(m);
  NcclApi* n = nccl_api();
  if (!n->lib) return fail(m, TF_ERR_CUDA, n->err);
  nccl_unique_id id;
  memcpy(id.internal, id128, sizeof(id.internal));
  const int rc = n->CommInitRank(&m->comm, m->cfg.n_ranks, id, m->cfg.rank);
  if (rc != 0) { m->comm = nullptr; return fail(m, TF_ERR_CUDA, std::string("ncclCommInitRank: ") + n->GetErrorString(rc)); }
  return TF_OK;
}

int tf_broadcast_frame(tf_map* m, int32_t frame_index, int has_color, int root) {
  if (!m || frame_index < 0 || root < 0 || root >= m->cfg.n_ranks) return fail(m, TF_ERR_INVALID, "tf_broadcast_frame: bad argument");
  if (!m->comm) return fail(m, TF_ERR_INVALID, "tf_broadcast_frame: call tf_comm_init first");
  use_device(m);
  int s;
  if (m->cfg.rank == root) {  // the ingest rank: the frame was uploaded (its copy is queued on the copy stream)
    s = find_slot(m, frame_index, false);
    if (s < 0) return fail(m, TF_ERR_NOT_FOUND, "tf_broadcast_frame: the root has not uploaded this frame");
    if (has_color && !(m->slots[s].has_rgba && m->slots[s].has_quality))
      return fail(m, TF_ERR_NOT_FOUND, "tf_broadcast_frame: the root's frame has no colour + quality planes");
  } else {
    s = acquire_slot(m, frame_index);
    if (s < 0) return TF_ERR_CAPACITY;
    guard_overwrite(m, m->slots[s]);
    m->slots[s].has_rgba = m->slots[s].has_quality = m->slots[s].has_rgb = false;
  }
  FrameSlot& fsl = m->slots[s];
  if (has_color) {
    if (int rc = ensure_color_planes(m, fsl)) return rc;
    fsl.has_rgba = fsl.has_quality = true;
  }
  // depth | rgba | quality are contiguous in the slot: one collective per frame, on the copy stream
  // (behind the root's H2D copy, next to the kernels of the frame being fused)
  const size_t plane = (((size_t)m->npix * 4) + 15) & ~(size_t)15;
  const size_t bytes = has_color ? 3 * plane : (size_t)m->npix * 4;
  const int rc = nccl_api()->Broadcast(fsl.depth, fsl.depth, bytes, kNcclUint8, root, m->comm, m->copy_stream);
  if (rc != 0) return fail(m, TF_ERR_CUDA, std::string("ncclBroadcast: ") + nccl_api()->GetErrorString(rc));
  CUDA_OK(m, cudaEventRecord(fsl.ready, m->copy_stream));
  fsl.pending = true;
  m->counters.kernel_launches++;  // (NCCL's broadcast kernel)
  return TF_OK;
}

// ---- prepare / integrate / finalize --------------------------------------------------------

int tf_prepare(tf_map* m, int32_t frame_index, const tf_pose* pose, const tf_camera* cam, tf_chunk_id* ids_out,
               uint8_t* is_new_out, int64_t cap, int64_t* n_out) {
  if (!m || !pose || !n_out || !cam_ok(m, cam)) return fail(m, TF_ERR_INVALID, "tf_prepare: bad argument");
  use_device(m);
  const int s = find_slot(m, frame_index);
  if (s < 0) return fail(m, TF_ERR_NOT_FOUND, "frame_index not in the frame store");
  CullParams cp;
  make_cull_params(m->cfg.voxel_res, m->cfg.trunc, m->cfg.dot3_order, *pose, *cam, cp);
  // pass 1: culling only, to learn the list length before anything is created
  static const GroupParams kNoFrames{};
  if (int rc = launch_cull(m, cp, kNoFrames, m->slots[s].depth, false, true)) return rc;
  publish_kernel<<<1, 1, 0, m->stream>>>(m->fs, m->res_d);
  if (int rc = check_kernel(m, "publish_kernel")) return rc;
  CUDA_OK(m, cudaStreamSynchronize(m->stream));
  if (int rc = dev_error_to_code(m, m->res_h->error)) return rc;
  const int64_t n = m->res_h->n_chunks;
  *n_out = n;
  if (n > cap || (n > 0 && (!ids_out || !is_new_out)))
    return fail(m, TF_ERR_CAPACITY, "tf_prepare: output capacity too small");
  // pass 2: HasChunk / CreateChunk
  alloc_kernel<<<m->grid_cull, kThreads, 0, m->stream>>>(cp, kNoFrames, m->md, m->fs, m->cb);
  if (int rc = check_kernel(m, "alloc_kernel")) return rc;
  publish_kernel<<<1, 1, 0, m->stream>>>(m->fs, m->res_d);
  if (int rc = check_kernel(m, "publish_kernel")) return rc;
  if (n > 0) {
    CUDA_OK(m, cudaMemcpyAsync(ids_out, m->cb.list_ids, (size_t)n * sizeof(int3), cudaMemcpyDeviceToHost, m->stream));
    CUDA_OK(m, cudaMemcpyAsync(is_new_out, m->cb.list_new, (size_t)n, cudaMemcpyDeviceToHost, m->stream));
    m->counters.d2h_bytes += n * 13;
  }
  CUDA_OK(m, cudaStreamSynchronize(m->stream));
  absorb_result(m);
  return dev_error_to_code(m, m->res_h->error);
}

int tf_integrate_group(tf_map* m, const tf_group_frame* frames, int32_t n_frames, const tf_camera* cam,
                       const tf_chunk_id* ids, int64_t n, uint8_t* needs_update, float* quality_out) {
  if (!m || !frames || !cam_ok(m, cam) || n < 0 || (n > 0 && (!ids || !needs_update)))
    return fail(m, TF_ERR_INVALID, "tf_integrate_group: bad argument");
  use_device(m);
  GroupParams gp;
  bool color[kMaxGroupFrames];
  if (int rc = build_group(m, frames, n_frames, cam, gp, color)) return rc;
  if (n == 0) return TF_OK;  // Structure/Chisel.h:228
  if (int rc = lookup_ids(m, ids, n, true, &gp)) return rc;
  if (int rc = launch_integrate(m, gp, nullptr, (int)n, algorithmic_bytes(m, n, color, n_frames))) return rc;
  CUDA_OK(m, cudaMemcpyAsync(m->upd_stage_h, m->list_upd, (size_t)n * sizeof(unsigned), cudaMemcpyDeviceToHost, m->stream));
  if (quality_out)
    CUDA_OK(m, cudaMemcpyAsync(m->q_stage_h, m->list_q, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, m->stream));
  CUDA_OK(m, cudaStreamSynchronize(m->stream));
  prof_collect(m, 0);
  m->counters.d2h_bytes += n * (quality_out ? 8 : 4);
  for (int64_t i = 0; i < n; i++) needs_update[i] = (needs_update[i] || m->upd_stage_h[i] != 0) ? 1 : 0;
  if (quality_out) memcpy(quality_out, m->q_stage_h, (size_t)n * sizeof(float));
  m->counters.frames_integrated += n_frames;
  m->counters.voxel_updates += n * 512 * n_frames;
  return TF_OK;
}

int tf_integrate(tf_map* m, int32_t frame_index, int use_color, const tf_pose* pose, const tf_camera* cam,
                 const tf_chunk_id* ids, int64_t n, int flag, uint8_t* needs_update, float* quality_out) {
  if (!m || !pose) return fail(m, TF_ERR_INVALID, "tf_integrate: bad argument");
  tf_group_frame g;
  g.frame_index = frame_index;
  g.use_color = use_color;
  g.flag = flag;
  g.reserved = 0;
  g.pose = *pose;
  return tf_integrate_group(m, &g, 1, cam, ids, n, needs_update, quality_out);
}

int tf_remove_chunks(tf_map* m, const tf_chunk_id* ids, int64_t n) {
  if (!m || n < 0 || (n > 0 && !ids)) return fail(m, TF_ERR_INVALID, "tf_remove_chunks: bad argument");
  use_device(m);
  if (n == 0) return TF_OK;
  if (int rc = upload_ids(m, ids, n)) return rc;
  const int grid = (int)std::min<int64_t>(m->grid, (n + kThreads - 1) / kThreads);
  remove_kernel<<<std::max(grid, 1), kThreads, 0, m->stream>>>(m->md, m->fs, m->cb.list_ids, (int)n);
  if (int rc = check_kernel(m, "remove_kernel")) return rc;
  publish_kernel<<<1, 1, 0, m->stream>>>(m->fs, m->res_d);
  if (int rc = check_kernel(m, "publish_kernel")) return rc;
  CUDA_OK(m, cudaStreamSynchronize(m->stream));
  absorb_result(m);
  return TF_OK;
}

// ---- the fused per-frame chain ---------------------------------------------------------------------

// Arguments of the three kernels of one fused frame (kept alive until the launch call returns).
struct FrameArgs {
  CullParams cp;
  GroupParams gp;
  const float* depth;
  int parity, want_order;
  bool any_color;
  FusedFinalize ff;
  const int* n_dev;
  int n_host;
  bool want_export;
  ExportArgs ex;
};

// export_kernel is PCIe-bound (eight blocks already saturate the link) and every block ends with a
// system-scope fence: a small grid finishes earlier than one block per SM.
static int export_grid(const tf_map* m) {
  static const int g = [] { const char* e = getenv("TEXFUSION_B200_EXPORT_GRID"); return e ? atoi(e) : 0; }();
  return std::min(g > 0 ? g : std::min(8, m->sm_count), kMaxExportBlocks);
}

static void launch_frame_kernels(tf_map* m, FrameArgs& a, bool profile = false) {
  bbox_kernel<<<m->grid_bbox, kThreads, 0, m->stream>>>(a.cp, a.depth, m->fs, a.parity);
  launch_pdl(cull_kernel<true>, m->grid_cull, 0, m->stream, a.cp, a.gp, m->md, a.depth, m->fs, m->cb, m->cfg.n_ranks,
             m->cfg.rank, a.parity, a.want_order);
  EventPair ep;
  if (profile) prof_begin(m, ep);
  launch_integrate_kernel(m, a.gp, a.any_color, a.n_dev, a.n_host, a.ff);
  if (profile) prof_end(m, ep);
  if (a.want_export) launch_pdl(export_kernel, export_grid(m), 0, m->stream, a.ex);
  m->counters.kernel_launches += a.want_export ? 4 : 3;
}

// Device-visible alias of a page-locked, 16-byte aligned host pointer (else nullptr).
static void* device_alias(const void* p) {
  if (!p || ((uintptr_t)p & 15u)) return nullptr;
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return at.type == cudaMemoryTypeHost ? at.devicePointer : nullptr;
}

static bool graphs_enabled() {  // on by default; TEXFUSION_B200_GRAPH=0 uses plain launches
  static const bool on = [] { const char* e = getenv("TEXFUSION_B200_GRAPH"); return !(e && e[0] == '0'); }();
  return on;
}

// Captures the chain once per (colour, group size).  Returns false when graphs cannot be used.
static bool build_frame_graph(tf_map* m, FrameGraph& fg, FrameArgs& a) {
  if (fg.failed) return false;
  fg.failed = true;  // until everything below has worked
  if (cudaStreamBeginCapture(m->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { cudaGetLastError(); return false; }
  launch_frame_kernels(m, a);
  cudaGraph_t g = nullptr;
  if (cudaStreamEndCapture(m->stream, &g) != cudaSuccess || !g) { cudaGetLastError(); return false; }
  fg.graph = g;
  if (cudaGraphInstantiate(&fg.exec, g, 0) != cudaSuccess) { cudaGetLastError(); return false; }
  cudaGraphNode_t nodes[8];
  size_t nn = 8;
  const size_t n_expected = a.want_export ? 4 : 3;
  if (cudaGraphGetNodes(g, nodes, &nn) != cudaSuccess || nn != n_expected) { cudaGetLastError(); return false; }
  const void* fn[4] = {(const void*)bbox_kernel, (const void*)cull_kernel<true>,
                       (const void*)integrate_fn(a.any_color, a.gp.n_frames),
                       (const void*)export_kernel};
  for (size_t i = 0; i < nn; i++) {
    cudaKernelNodeParams kp{};
    if (cudaGraphKernelNodeGetParams(nodes[i], &kp) != cudaSuccess) { cudaGetLastError(); return false; }
    for (size_t k = 0; k < n_expected; k++)
      if (kp.func == fn[k]) { fg.node[k] = nodes[i]; fg.kp[k] = kp; }
  }
  for (size_t k = 0; k < n_expected; k++)
    if (!fg.node[k]) return false;
  fg.failed = false;
  return true;
}

static int launch_frame_graph(tf_map* m, FrameGraph& fg, FrameArgs& a) {
  // argument lists in declaration order of the kernels (tf_kernels.cuh)
  const int* list_slots = m->cb.list_slots;
  const int* list_hpos = m->cb.list_hpos;
  const float* list_setup = m->cb.list_setup;
  void* p_bbox[] = {&a.cp, &a.depth, &m->fs, &a.parity};
  void* p_cull[] = {&a.cp, &a.gp, &m->md, &a.depth, &m->fs, &m->cb, &m->cfg.n_ranks, &m->cfg.rank, &a.parity, &a.want_order};
  void* p_int[] = {&a.gp, &m->md, &list_slots, &list_hpos, &list_setup, &a.n_dev, &a.n_host, &m->list_upd, &m->list_q,
                   &a.ff};
  void* p_exp[] = {&a.ex};
  void** params[4] = {p_bbox, p_cull, p_int, p_exp};
  const int n_nodes = a.want_export ? 4 : 3;
  HT(2);
  for (int k = 0; k < n_nodes; k++) {
    fg.kp[k].kernelParams = params[k];
    fg.kp[k].extra = nullptr;
    CUDA_OK(m, cudaGraphExecKernelNodeSetParams(fg.exec, fg.node[k], &fg.kp[k]));
  }
  HT(3);
  CUDA_OK(m, cudaGraphLaunch(fg.exec, m->stream));
  HT(4);
  m->counters.kernel_launches += n_nodes;
  return TF_OK;
}

// Waits for the completion stamp of a fused frame (written last by publish_frame into mapped host
// memory).  Polling it returns a few microseconds earlier than cudaStreamSynchronize; the stream
// is queried now and then so that a failed kernel cannot hang the caller.
static int wait_frame(tf_map* m, unsigned seq, int n_export_blocks) {
  volatile FrameResultHost* r = m->res_h;
  auto done = [&] {
    if (!(r->seq == seq && r->seq0 == seq && r->seq1 == seq)) return false;
    for (int b = 0; b < n_export_blocks; b++)  // every export block's part of the lists has arrived
      if (r->blk[b] != seq) return false;
    return true;
  };
  for (unsigned it = 1; !done(); it++) {
    if ((it & 0x3fffu) == 0) {
      const cudaError_t e = cudaStreamQuery(m->stream);
      if (e == cudaSuccess) break;  // finished: the stamps are there (or the kernels did not publish)
      if (e != cudaErrorNotReady) return fail(m, TF_ERR_CUDA, std::string("fused frame: ") + cudaGetErrorString(e));
    }
  }
  std::atomic_thread_fence(std::memory_order_acquire);
  if (!done()) {  // stream idle without a stamp: surface whatever went wrong
    CUDA_OK(m, cudaStreamSynchronize(m->stream));
    if (!done()) return fail(m, TF_ERR_CUDA, "fused frame: kernels finished without publishing a result");
  }
  return TF_OK;
}

// Shared body of the fused pipelines: prepare on frames[0], integrate the group, finalize.
static int fused_group_begin(tf_map* m, const tf_group_frame* frames, int n_frames, const tf_camera* cam,
                             tf_chunk_id* ids_out, uint8_t* new_out, uint8_t* upd_out, float* q_out, int64_t cap) {
  HT(0);
  if (m->pend.active) return fail(m, TF_ERR_INVALID, "a fused frame is already in flight (call tf_integrate_frame_end first)");
  FrameArgs a;
  bool* color = m->pend.color;
  if (int rc = build_group(m, frames, n_frames, cam, a.gp, color)) return rc;
  const int s = find_slot(m, frames[0].frame_index);
  make_cull_params(m->cfg.voxel_res, m->cfg.trunc, m->cfg.dot3_order, frames[0].pose, *cam, a.cp);
  if (int rc = ensure_setup(m, n_frames)) return rc;
  HT(1);
  const bool want_lists = ids_out || new_out || upd_out || q_out;
  const int ocap = (int)std::min<int64_t>(cap, m->list_cap);
  a.depth = m->slots[s].depth;
  a.want_order = want_lists ? 1 : 0;
  a.any_color = false;
  for (int f = 0; f < n_frames; f++) a.any_color |= color[f];
  a.n_dev = &m->fs->n_work;
  a.n_host = 0;
  FusedFinalize& ff = a.ff;
  ff = FusedFinalize{};
  ff.enabled = 1;
  ff.ordered = want_lists ? 1 : 0;
  ff.fs = m->fs;
  ff.cb = m->cb;
  ff.ids_out = ids_out ? m->st_ids : nullptr;
  ff.new_out = new_out ? m->st_new : nullptr;
  ff.upd_out = upd_out ? m->st_upd : nullptr;
  ff.q_out = q_out ? m->st_q : nullptr;
  ff.out_cap = ocap;
  ff.res = m->res_d;
  ff.seq = ++m->seq;
  ff.export_follows = want_lists ? 1 : 0;
  // Lists go straight into the caller's buffers when those are page-locked (tf_host_alloc) and
  // 16-byte aligned, else through the map's own mapped staging buffers + a host copy.
  a.want_export = want_lists;
  ExportArgs& ex = a.ex;
  ex = ExportArgs{};
  void* direct[4] = {device_alias(ids_out), device_alias(new_out), device_alias(upd_out), device_alias(q_out)};
  if (want_lists) {
    ex.ids_s = ff.ids_out, ex.new_s = ff.new_out, ex.upd_s = ff.upd_out, ex.q_s = ff.q_out;
    ex.ids_h = ids_out ? (direct[0] ? (int3*)direct[0] : m->out_ids_d) : nullptr;
    ex.new_h = new_out ? (direct[1] ? (unsigned char*)direct[1] : m->out_new_d) : nullptr;
    ex.upd_h = upd_out ? (direct[2] ? (unsigned char*)direct[2] : m->out_upd_d) : nullptr;
    ex.q_h = q_out ? (direct[3] ? (float*)direct[3] : m->out_q_d) : nullptr;
    ex.cap = ocap;
    ex.fs = m->fs;
    ex.res = m->res_d;
    ex.seq = ff.seq;
    ex.batch_item = -1;
  }

  // bbox -> cull (+ HasChunk / CreateChunk) -> integrate (+ Finalize, garbage collection, publication)
  FrameGraph& fg = m->graphs[want_lists ? 1 : 0][a.any_color ? 1 : 0][n_frames];
  bool launched = false;
  if (m->prof == 0 && graphs_enabled() && (fg.exec || build_frame_graph(m, fg, a)) && !fg.failed) {
    m->parity ^= 1;
    a.parity = m->parity;
    if (int rc = launch_frame_graph(m, fg, a)) return rc;
    HT(5);
    launched = true;
  }
  if (!launched) {  // plain launches (profiling with events between the stages, or graphs unavailable)
    if (int rc = launch_cull(m, a.cp, a.gp, a.depth, true, want_lists)) return rc;
    a.parity = m->parity;
    if (int rc = launch_integrate(m, a.gp, a.n_dev, 0, -1, &ff)) return rc;
    if (a.want_export) {
      launch_pdl(export_kernel, export_grid(m), 0, m->stream, a.ex);
      if (int rc = check_kernel(m, "export_kernel")) return rc;
    }
    HT(5);
  }
  tf_map::PendingFrame& pf = m->pend;
  pf.active = true;
  pf.ticket = m->compute_ticket;
  pf.graph = launched;
  pf.seq = ff.seq;
  pf.n_frames = n_frames;
  pf.n_export_blocks = want_lists ? export_grid(m) : 0;
  pf.ocap = ocap;
  pf.ids_out = ids_out, pf.new_out = new_out, pf.upd_out = upd_out, pf.q_out = q_out;
  for (int k = 0; k < 4; k++) pf.direct[k] = direct[k] != nullptr;
  return TF_OK;
}

static int fused_group_end(tf_map* m, tf_frame_stats* stats) {
  tf_map::PendingFrame& pf = m->pend;
  if (!pf.active) return fail(m, TF_ERR_INVALID, "tf_integrate_frame_end: no fused frame in flight");
  pf.active = false;
  if (pf.graph) {
    if (int rc = wait_frame(m, pf.seq, pf.n_export_blocks)) return rc;
  } else {
    CUDA_OK(m, cudaStreamSynchronize(m->stream));
  }
  m->compute_done = std::max(m->compute_done, pf.ticket);
  const bool* color = pf.color;
  const int n_frames = pf.n_frames, ocap = pf.ocap;
  tf_chunk_id* ids_out = pf.ids_out;
  uint8_t *new_out = pf.new_out, *upd_out = pf.upd_out;
  float* q_out = pf.q_out;
  const bool* direct = pf.direct;
  HT(6);
  absorb_result(m);
  const FrameResultHost r = *m->res_h;
  prof_collect(m, algorithmic_bytes(m, r.n_chunks, color, n_frames));
  const int64_t nout = std::min<int64_t>(r.n_chunks, ocap);
  if (ids_out && !direct[0]) memcpy(ids_out, m->out_ids_h, (size_t)nout * sizeof(int3));
  if (new_out && !direct[1]) memcpy(new_out, m->out_new_h, (size_t)nout);
  if (upd_out && !direct[2]) memcpy(upd_out, m->out_upd_h, (size_t)nout);
  if (q_out && !direct[3]) memcpy(q_out, m->out_q_h, (size_t)nout * sizeof(float));
  m->counters.d2h_bytes += sizeof(FrameResultHost) + nout * ((ids_out ? 12 : 0) + (new_out ? 1 : 0) + (upd_out ? 1 : 0) + (q_out ? 4 : 0));
  if (stats) {
    stats->n_chunks = r.n_chunks;
    stats->n_new = r.n_new;
    stats->n_updated = r.n_updated;
    stats->n_removed = r.n_removed;
    stats->voxel_updates = (int64_t)r.n_chunks * 512 * n_frames;
  }
  m->counters.frames_integrated += n_frames;
  m->counters.voxel_updates += (int64_t)r.n_chunks * 512 * n_frames;
  return dev_error_to_code(m, r.error);
}

int tf_integrate_frame(tf_map* m, int32_t frame_index, int use_color, const tf_pose* pose, const tf_camera* cam,
                       tf_frame_stats* stats, tf_chunk_id* ids_out, uint8_t* is_new_out, uint8_t* updated_out,
                       float* quality_out, int64_t cap) {
  if (!m || !pose || !cam_ok(m, cam) || cap < 0) return fail(m, TF_ERR_INVALID, "tf_integrate_frame: bad argument");
  use_device(m);
  tf_group_frame g;
  g.frame_index = frame_index;
  g.use_color = use_color;
  g.flag = 1;
  g.reserved = 0;
  g.pose = *pose;
  if (int rc = fused_group_begin(m, &g, 1, cam, ids_out, is_new_out, updated_out, quality_out, cap)) return rc;
  return fused_group_end(m, stats);
}

int tf_integrate_frame_begin(tf_map* m, int32_t frame_index, int use_color, const tf_pose* pose, const tf_camera* cam,
                             tf_chunk_id* ids_out, uint8_t* is_new_out, uint8_t* updated_out, float* quality_out, int64_t cap) {
  if (!m || !pose || !cam_ok(m, cam) || cap < 0) return fail(m, TF_ERR_INVALID, "tf_integrate_frame_begin: bad argument");
  use_device(m);
  tf_group_frame g;
  g.frame_index = frame_index;
  g.use_color = use_color;
  g.flag = 1;
  g.reserved = 0;
  g.pose = *pose;
  return fused_group_begin(m, &g, 1, cam, ids_out, is_new_out, updated_out, quality_out, cap);
}

int tf_integrate_frame_end(tf_map* m, tf_frame_stats* stats_out) {
  if (!m) return TF_ERR_INVALID;
  use_device(m);
  return fused_group_end(m, stats_out);
}

int tf_stream_step(tf_map* m, const tf_camera* cam, const tf_stream_step_args* a, tf_frame_stats* stats_out) {
  if (!m || !a) return fail(m, TF_ERR_INVALID, "tf_stream_step: bad argument");
  if (int rc = tf_integrate_frame_begin(m, a->frame_index, a->use_color, &a->pose, cam, a->ids_out, a->is_new_out, a->updated_out,
                                        a->quality_out, a->cap))
    return rc;
  int rc_in = TF_OK;  // (the fused frame is in flight: it is collected below whatever the ingest calls return)
  if (a->next_index >= 0) {
    if (a->next_depth) rc_in = tf_upload_frame(m, a->next_index, a->next_depth, a->next_rgba, a->next_quality);
    if (!rc_in && m->comm) rc_in = tf_broadcast_frame(m, a->next_index, a->next_has_color, a->broadcast_root);
  }
  const std::string err_in = rc_in ? m->err : std::string();
  if (int rc = tf_integrate_frame_end(m, stats_out)) return rc;
  if (rc_in) return fail(m, rc_in, err_in);
  if (a->wait_index >= 0) return tf_wait_upload(m, a->wait_index);
  return TF_OK;
}

// ---- loop-closure batches ---------------------------------------------------------------------------
//
// All items of a batch are queued without intermediate host synchronisation: de-integration =
// id upload + lookup + integrate; re-integration = the fused chain, whose ordered lists go to a
// slice of a result arena in mapped host memory (export_kernel, batch mode).  The host synchronises
// once per sub-batch and then derives every item's valid list (updated chunks, in list order).

constexpr int kArenaCap = 1 << 22;     // list entries per sub-batch
constexpr int kSubBatch = 32;          // re-integration items per synchronisation
constexpr int64_t kIdsStage = 1 << 20; // staged de-integration ids per synchronisation

// (page-locking tens of megabytes takes tens of milliseconds: only what a batch needs, once)
static int ensure_arena(tf_map* m, bool lists, bool ids) {
  if (!m->batch_rec) CUDA_OK(m, cudaHostAlloc((void**)&m->batch_rec, (size_t)kSubBatch * sizeof(int4), cudaHostAllocMapped));
  if (lists && !m->arena_ids) {
    CUDA_OK(m, cudaHostAlloc((void**)&m->arena_ids, (size_t)kArenaCap * sizeof(int3), cudaHostAllocMapped));
    CUDA_OK(m, cudaHostAlloc((void**)&m->arena_q, (size_t)kArenaCap * sizeof(float), cudaHostAllocMapped));
    CUDA_OK(m, cudaHostAlloc((void**)&m->arena_upd, (size_t)kArenaCap, cudaHostAllocMapped));
  }
  if (ids && !m->ids_stage)
    CUDA_OK(m, cudaHostAlloc((void**)&m->ids_stage, (size_t)kIdsStage * sizeof(int3), cudaHostAllocDefault));
  return TF_OK;
}

struct PendingItem {
  const tf_batch_item* it;
  int rec;          // index into batch_rec
  unsigned seq;     // completion stamp export_kernel writes into batch_rec[rec].w
  int n_frames;
  bool color[kMaxGroupFrames];
  size_t first_event, n_events;  // profiling events of this item (m->ev_pending)
};

// One finished re-integration item: its valid list (updated chunks, in list order) from the arena
// into the caller's arrays.
static int collect_item(tf_map* m, const PendingItem& p) {
  const int4 r = m->batch_rec[p.rec];
  if (r.x < 0) return fail(m, TF_ERR_CAPACITY, "tf_integrate_batch: result arena exhausted");
  int rc = TF_OK;
  const int3* ids = m->arena_ids + r.y;
  const float* q = m->arena_q + r.y;
  const unsigned char* upd = m->arena_upd + r.y;
  tf_chunk_id* vout = p.it->valid_out;
  float* qout = p.it->quality_out;
  const int64_t cap = p.it->cap;
  int64_t nv = 0;
  for (int i = 0; i < r.x; i++) {
    if (!upd[i]) continue;
    if (nv < cap) {
      if (vout) memcpy(&vout[nv], &ids[i], sizeof(int3));
      if (qout) qout[nv] = q[i];
    }
    nv++;
  }
  if (p.it->n_valid_out) *p.it->n_valid_out = nv;
  if (nv > cap && vout) rc = fail(m, TF_ERR_CAPACITY, "tf_integrate_batch: valid_out too small");
  m->counters.d2h_bytes += (int64_t)r.x * 17 + sizeof(int4);
  m->counters.frames_integrated += p.n_frames;
  m->counters.voxel_updates += (int64_t)r.z * 512 * p.n_frames;
  for (size_t e = p.first_event; e < p.first_event + p.n_events && e < m->ev_pending.size(); e++)
    m->ev_pending[e].bytes = algorithmic_bytes(m, r.z, p.color, p.n_frames);
  return rc;
}

// Items whose completion stamp has arrived are collected while the device works on the later ones
// (called between the enqueues of a batch; never blocks).  `done`: items of `pend` already collected.
static int collect_ready(tf_map* m, const std::vector<PendingItem>& pend, size_t& done) {
  int rc = TF_OK;
  while (done < pend.size()) {
    const volatile int* w = &m->batch_rec[pend[done].rec].w;
    if ((unsigned)*w != pend[done].seq) break;
    std::atomic_thread_fence(std::memory_order_acquire);
    if (int rc1 = collect_item(m, pend[done])) rc = rc1;
    done++;
  }
  return rc;
}

// synchronise, then hand the re-integration items that have not been collected yet their valid lists
static int flush_batch(tf_map* m, std::vector<PendingItem>& pend, size_t& done, int rc_collected) {
  publish_kernel<<<1, 1, 0, m->stream>>>(m->fs, m->res_d);  // error bits and allocator state of the whole sub-batch
  if (int rc = check_kernel(m, "publish_kernel")) return rc;
  CUDA_OK(m, cudaStreamSynchronize(m->stream));
  m->compute_done = m->compute_ticket;
  absorb_result(m);
  const int dev_err = m->res_h->error;
  int rc = rc_collected;  // (an item collected early may already have failed: capacity of its valid_out)
  for (; done < pend.size(); done++)
    if (int rc1 = collect_item(m, pend[done])) rc = rc1;
  pend.clear();
  done = 0;
  m->ids_stage_used = 0;
  prof_collect(m, 0);
  if (int rc2 = dev_error_to_code(m, dev_err)) return rc2;
  return rc;
}

int tf_integrate_batch(tf_map* m, const tf_batch_item* items, int64_t n_items, const tf_camera* cam) {
  if (!m || n_items < 0 || (n_items > 0 && !items) || !cam_ok(m, cam))
    return fail(m, TF_ERR_INVALID, "tf_integrate_batch: bad argument");
  use_device(m);
  int max_frames = 1;
  bool any_lists = false, any_ids = false;
  for (int64_t k = 0; k < n_items; k++) {
    if (items[k].flag == 0) any_ids = true;
    else if (items[k].valid_out || items[k].quality_out || items[k].n_valid_out) any_lists = true;
    if (!items[k].frames || items[k].n_frames < 1 || items[k].n_frames > kMaxGroupFrames)
      return fail(m, TF_ERR_INVALID, "tf_integrate_batch: group size must be 1..8");
    if (items[k].flag == 0 && (items[k].n_ids < 0 || (items[k].n_ids > 0 && !items[k].ids)))
      return fail(m, TF_ERR_INVALID, "tf_integrate_batch: de-integration item without a chunk list");
    max_frames = std::max(max_frames, items[k].n_frames);
  }
  if (int rc = ensure_arena(m, any_lists, any_ids)) return rc;
  if (int rc = ensure_setup(m, max_frames)) return rc;
  CUDA_OK(m, cudaMemsetAsync(&m->fs->arena_off, 0, sizeof(int), m->stream));
  std::vector<PendingItem> pend;
  size_t done = 0;   // items of `pend` collected while later ones were being queued
  int rc_early = TF_OK;
  tf_group_frame fr[kMaxGroupFrames];
  for (int64_t k = 0; k < n_items; k++) {
    const tf_batch_item& it = items[k];
    if (int rc1 = collect_ready(m, pend, done)) rc_early = rc1;
    for (int f = 0; f < it.n_frames; f++) {
      fr[f] = it.frames[f];
      fr[f].flag = it.flag ? 1 : 0;
    }
    if (it.flag == 0) {
      // de-integration over kf.validChunks (GCFusion/MobileFusion.cpp:135-143); ids that are not in
      // the map are skipped by the kernels and reported when the batch is flushed
      if (it.n_ids == 0) continue;  // Structure/Chisel.h:228
      GroupParams gp;
      bool color[kMaxGroupFrames];
      if (int rc = build_group(m, fr, it.n_frames, cam, gp, color)) return rc;
      if (it.n_ids > m->list_cap) return fail(m, TF_ERR_CAPACITY, "chunk list exceeds the list capacity");
      if (it.n_ids > kIdsStage - m->ids_stage_used) {  // staging full: drain what is queued
        if (int rc = flush_batch(m, pend, done, rc_early)) return rc;
        rc_early = TF_OK;
        CUDA_OK(m, cudaMemsetAsync(&m->fs->arena_off, 0, sizeof(int), m->stream));
      }
      const tf_chunk_id* src = it.ids;
      if (it.n_ids <= kIdsStage) {
        memcpy(m->ids_stage + m->ids_stage_used, it.ids, (size_t)it.n_ids * sizeof(int3));
        src = reinterpret_cast<const tf_chunk_id*>(m->ids_stage + m->ids_stage_used);
        m->ids_stage_used += it.n_ids;
      }
      if (int rc = upload_ids(m, src, it.n_ids)) return rc;
      const int grid = (int)std::min<int64_t>(m->grid, (it.n_ids + kThreads - 1) / kThreads);
      lookup_kernel<<<std::max(grid, 1), kThreads, 0, m->stream>>>(gp, m->md, m->fs, m->cb.list_ids, (int)it.n_ids,
                                                                   m->cb.list_slots, m->cb.list_hpos, m->cb.list_setup);
      if (int rc = check_kernel(m, "lookup_kernel")) return rc;
      if (int rc = launch_integrate(m, gp, nullptr, (int)it.n_ids, algorithmic_bytes(m, it.n_ids, color, it.n_frames)))
        return rc;
      m->counters.frames_integrated += it.n_frames;
      m->counters.voxel_updates += it.n_ids * 512 * it.n_frames;
    } else {
      FrameArgs a;
      PendingItem p{};
      p.it = &it;
      p.rec = (int)pend.size();
      p.n_frames = it.n_frames;
      if (int rc = build_group(m, fr, it.n_frames, cam, a.gp, p.color)) return rc;
      const int s = find_slot(m, fr[0].frame_index);
      make_cull_params(m->cfg.voxel_res, m->cfg.trunc, m->cfg.dot3_order, fr[0].pose, *cam, a.cp);
      // (an item that asks for nothing back — streaming fusion of a frame sequence — skips the
      //  ordering scan and the export; only its list length is recorded)
      const bool want_lists = it.valid_out || it.quality_out || it.n_valid_out;
      a.depth = m->slots[s].depth;
      a.want_order = want_lists ? 1 : 0;
      a.any_color = false;
      for (int f = 0; f < it.n_frames; f++) a.any_color |= p.color[f];
      a.n_dev = &m->fs->n_work;
      a.n_host = 0;
      a.ff = FusedFinalize{};
      a.ff.enabled = 1;
      a.ff.ordered = want_lists ? 1 : 0;
      a.ff.fs = m->fs;
      a.ff.cb = m->cb;
      a.ff.ids_out = want_lists ? m->st_ids : nullptr;
      a.ff.upd_out = want_lists ? m->st_upd : nullptr;
      a.ff.q_out = want_lists ? m->st_q : nullptr;
      a.ff.out_cap = m->list_cap;
      a.ff.res = m->res_d;
      a.ff.seq = ++m->seq;
      p.seq = a.ff.seq;
      a.ff.export_follows = 1;
      a.want_export = true;  // (without lists the export kernel only records the item's list length)
      a.ex = ExportArgs{};
      if (want_lists) {
        a.ex.ids_s = m->st_ids, a.ex.upd_s = m->st_upd, a.ex.q_s = m->st_q;
        CUDA_OK(m, cudaHostGetDevicePointer((void**)&a.ex.ids_h, m->arena_ids, 0));
        CUDA_OK(m, cudaHostGetDevicePointer((void**)&a.ex.upd_h, m->arena_upd, 0));
        CUDA_OK(m, cudaHostGetDevicePointer((void**)&a.ex.q_h, m->arena_q, 0));
      }
      CUDA_OK(m, cudaHostGetDevicePointer((void**)&a.ex.batch_rec, m->batch_rec, 0));
      a.ex.cap = want_lists ? m->list_cap : 0;
      a.ex.fs = m->fs;
      a.ex.res = m->res_d;
      a.ex.seq = a.ff.seq;
      a.ex.batch_item = p.rec;
      a.ex.arena_cap = kArenaCap;
      m->parity ^= 1;
      a.parity = m->parity;
      p.first_event = m->ev_pending.size();
      launch_frame_kernels(m, a, m->prof != 0);
      p.n_events = m->ev_pending.size() - p.first_event;
      if (int rc = check_kernel(m, "fused frame chain")) return rc;
      m->counters.kernel_launches--;  // (check_kernel counted one launch too many)
      pend.push_back(p);
      if ((int)pend.size() == kSubBatch) {
        if (int rc = flush_batch(m, pend, done, rc_early)) return rc;
        rc_early = TF_OK;
        CUDA_OK(m, cudaMemsetAsync(&m->fs->arena_off, 0, sizeof(int), m->stream));
      }
    }
  }
  return flush_batch(m, pend, done, rc_early);
}

#ifdef TF_TIMELINE
// Debug build only: per-kernel device timestamps of the last frame (see TL_MARK).
extern "C" int tf_debug_host(double* out16) { memcpy(out16, g_host_t, sizeof(g_host_t)); return 0; }
extern "C" int tf_debug_trace(tf_map* m, unsigned long long* out512) {
  if (!m || !out512) return TF_ERR_INVALID;
  cudaStreamSynchronize(m->stream);
  cudaMemcpyFromSymbol(out512, g_trace, sizeof(unsigned long long) * 512);
  cudaMemcpyFromSymbol(out512 + 512, g_trace2, sizeof(unsigned long long) * 512);
  return TF_OK;
}
extern "C" int tf_debug_timeline(tf_map* m, unsigned long long* out32, int reset) {
  if (!m) return TF_ERR_INVALID;
  cudaStreamSynchronize(m->stream);
  if (out32) {
    cudaMemcpyFromSymbol(out32, g_timeline, sizeof(unsigned long long) * 32);
    FrameState f;
    cudaMemcpy(&f, m->fs, sizeof(f), cudaMemcpyDeviceToHost);
    out32[29] = f.n_coarse;
    out32[30] = f.n_hit_cands;
    out32[31] = f.n_list;
  }
  if (reset) {
    unsigned long long z[32] = {};
    cudaMemcpyToSymbol(g_timeline, z, sizeof(z));
    static unsigned long long zt[512] = {};
    cudaMemcpyToSymbol(g_trace, zt, sizeof(zt));
    cudaMemcpyToSymbol(g_trace2, zt, sizeof(zt));
  }
  return TF_OK;
}
#endif

// ---- queries ---------------------------------------------------------------------------------

int tf_has_chunk(tf_map* m, tf_chunk_id id) {
  if (!m) return TF_ERR_INVALID;
  use_device(m);
  if (int rc = lookup_ids(m, &id, 1, false)) return rc;
  int slot = -1;
  CUDA_OK(m, cudaMemcpy(&slot, m->cb.list_slots, sizeof(int), cudaMemcpyDeviceToHost));
  return slot >= 0 ? 1 : 0;  // (a lazy bit may be set; the sign is what matters)
}

int64_t tf_chunk_count(tf_map* m) { return m ? m->n_live : TF_ERR_INVALID; }

int tf_list_chunks(tf_map* m, tf_chunk_id* out, int64_t cap, int64_t* n_out) {
  if (!m || !n_out) return fail(m, TF_ERR_INVALID, "tf_list_chunks: bad argument");
  use_device(m);
  *n_out = m->n_live;
  if (m->n_live == 0) return TF_OK;
  if (cap < m->n_live || !out) return fail(m, TF_ERR_CAPACITY, "tf_list_chunks: output capacity too small");
  if (int rc = ensure_arena_bytes(m, m->ar_list, (size_t)m->n_live * sizeof(int3))) return rc;
  int3* tmp = (int3*)m->ar_list.p;
  cudaMemsetAsync(m->count_d, 0, sizeof(int), m->stream);
  list_kernel<<<m->grid, kThreads, 0, m->stream>>>(m->md, m->pool_next, tmp, (int)m->n_live, m->count_d);
  if (int rc = check_kernel(m, "list_kernel")) return rc;
  CUDA_OK(m, cudaMemcpyAsync(out, tmp, (size_t)m->n_live * sizeof(int3), cudaMemcpyDeviceToHost, m->stream));
  CUDA_OK(m, cudaStreamSynchronize(m->stream));
  m->counters.d2h_bytes += m->n_live * 12;
  return TF_OK;
}

int tf_download_chunks(tf_map* m, const tf_chunk_id* ids, int64_t n, float* sdf, float* weight, uint16_t* color) {
  if (!m || n < 0 || (n > 0 && !ids)) return fail(m, TF_ERR_INVALID, "tf_download_chunks: bad argument");
  use_device(m);
  if (!m->dl_sdf) {
    m->dl_cap = 8192;
    CUDA_OK(m, dmalloc(&m->dl_sdf, (size_t)m->dl_cap * 512));
    CUDA_OK(m, dmalloc(&m->dl_w, (size_t)m->dl_cap * 512));
    CUDA_OK(m, dmalloc(&m->dl_col, (size_t)m->dl_cap * 512));
  }
  for (int64_t base = 0; base < n; base += m->dl_cap) {
    const int64_t cnt = std::min<int64_t>(m->dl_cap, n - base);
    if (int rc = lookup_ids(m, ids + base, cnt, true)) return rc;
    const int grid = (int)std::min<int64_t>(m->grid * 4, (cnt + kWarpsPerBlock - 1) / kWarpsPerBlock);
    download_kernel<<<std::max(grid, 1), kThreads, 0, m->stream>>>(m->md, m->cb.list_slots, (int)cnt, sdf ? m->dl_sdf : nullptr,
                                                                   weight ? m->dl_w : nullptr, color ? m->dl_col : nullptr);
    if (int rc = check_kernel(m, "download_kernel")) return rc;
    if (sdf) CUDA_OK(m, cudaMemcpyAsync(sdf + base * 512, m->dl_sdf, (size_t)cnt * 2048, cudaMemcpyDeviceToHost, m->stream));
    if (weight) CUDA_OK(m, cudaMemcpyAsync(weight + base * 512, m->dl_w, (size_t)cnt * 2048, cudaMemcpyDeviceToHost, m->stream));
    if (color) CUDA_OK(m, cudaMemcpyAsync(color + base * 2048, m->dl_col, (size_t)cnt * 4096, cudaMemcpyDeviceToHost, m->stream));
    CUDA_OK(m, cudaStreamSynchronize(m->stream));
    m->counters.d2h_bytes += cnt * ((sdf ? 2048 : 0) + (weight ? 2048 : 0) + (color ? 4096 : 0));
  }
  return TF_OK;
}

// ---- meshing -------------------------------------------------------------------------------------

int tf_mesh_chunks(tf_map* m, const tf_chunk_id* ids, int64_t n, int64_t* vert_off, int64_t* idx_off, float* vertices,
                   float* normals, float* colors, int32_t* indices, int64_t vert_cap, int64_t idx_cap) {
  if (!m || n < 0 || (n > 0 && (!ids || !vert_off || !idx_off)) || n > (1 << 24))
    return fail(m, TF_ERR_INVALID, "tf_mesh_chunks: bad argument");
  const bool want = vertices || normals || colors || indices;
  if (want && !(vertices && normals && colors && indices)) return fail(m, TF_ERR_INVALID, "tf_mesh_chunks: pass all four output arrays or none");
  use_device(m);
  if (vert_off) vert_off[0] = 0;
  if (idx_off) idx_off[0] = 0;
  if (n == 0) return TF_OK;
  if (int rc = ensure_arena_bytes(m, m->ar_mesh_ids, (size_t)n * sizeof(int3))) return rc;
  if (int rc = ensure_arena_bytes(m, m->ar_mesh_counts, (size_t)n * sizeof(int2))) return rc;
  if (int rc = ensure_arena_bytes(m, m->ar_mesh_off, (size_t)(2 * n + 2) * sizeof(long long))) return rc;
  if (int rc = ensure_arena_bytes(m, m->ar_mesh_off_h, (size_t)(2 * n + 2) * sizeof(long long), true)) return rc;
  CUDA_OK(m, cudaMemcpyAsync(m->ar_mesh_ids.p, ids, (size_t)n * sizeof(int3), cudaMemcpyHostToDevice, m->stream));
  m->counters.h2d_bytes += n * (int64_t)sizeof(int3);
  MeshArgs a{};
  a.ids = (const int3*)m->ar_mesh_ids.p;
  a.n = (int)n;
  a.res = m->cfg.voxel_res;
  a.l2r = m->cfg.dot3_order;
  a.counts = (int2*)m->ar_mesh_counts.p;
  a.off = (const long long*)m->ar_mesh_off.p;
  const int grid = (int)std::min<int64_t>(n, (int64_t)m->sm_count * 3);
  mesh_kernel<false><<<grid, kMeshThreads, 0, m->stream>>>(m->md, a);
  if (int rc = check_kernel(m, "mesh_kernel<count>")) return rc;
  mesh_scan_kernel<<<1, 1024, 0, m->stream>>>(a.counts, a.n, (long long*)m->ar_mesh_off.p);
  if (int rc = check_kernel(m, "mesh_scan_kernel")) return rc;
  long long* off_h = (long long*)m->ar_mesh_off_h.p;
  CUDA_OK(m, cudaMemcpyAsync(off_h, m->ar_mesh_off.p, (size_t)(2 * n + 2) * sizeof(long long), cudaMemcpyDeviceToHost, m->stream));
  CUDA_OK(m, cudaStreamSynchronize(m->stream));
  m->counters.d2h_bytes += (2 * n + 2) * (int64_t)sizeof(long long);
  for (int64_t i = 0; i <= n; i++) { vert_off[i] = off_h[i]; idx_off[i] = off_h[n + 1 + i]; }
  const int64_t nv = off_h[n], ni = off_h[2 * n + 1];
  if (!want) return TF_OK;  // size query
  if (nv > vert_cap || ni > idx_cap) return fail(m, TF_ERR_CAPACITY, "tf_mesh_chunks: output capacity too small (the offsets hold the required sizes)");
  if (nv == 0) return TF_OK;
  if (int rc = ensure_arena_bytes(m, m->ar_mesh_v, (size_t)nv * 12)) return rc;
  if (int rc = ensure_arena_bytes(m, m->ar_mesh_n, (size_t)nv * 12)) return rc;
  if (int rc = ensure_arena_bytes(m, m->ar_mesh_c, (size_t)nv * 12)) return rc;
  if (int rc = ensure_arena_bytes(m, m->ar_mesh_i, (size_t)std::max<int64_t>(ni, 1) * 4)) return rc;
  a.vert = (float*)m->ar_mesh_v.p, a.norm = (float*)m->ar_mesh_n.p, a.col = (float*)m->ar_mesh_c.p, a.idx = (int*)m->ar_mesh_i.p;
  mesh_kernel<true><<<grid, kMeshThreads, 0, m->stream>>>(m->md, a);
  if (int rc = check_kernel(m, "mesh_kernel<write>")) return rc;
  CUDA_OK(m, cudaMemcpyAsync(vertices, a.vert, (size_t)nv * 12, cudaMemcpyDeviceToHost, m->stream));
  CUDA_OK(m, cudaMemcpyAsync(normals, a.norm, (size_t)nv * 12, cudaMemcpyDeviceToHost, m->stream));
  CUDA_OK(m, cudaMemcpyAsync(colors, a.col, (size_t)nv * 12, cudaMemcpyDeviceToHost, m->stream));
  CUDA_OK(m, cudaMemcpyAsync(indices, a.idx, (size_t)ni * 4, cudaMemcpyDeviceToHost, m->stream));
  CUDA_OK(m, cudaStreamSynchronize(m->stream));
  m->counters.d2h_bytes += nv * 36 + ni * 4;
  return TF_OK;
}

// ---- atlas ---------------------------------------------------------------------------------------

int tf_atlas_patch_size(tf_map* m, int32_t* w, int32_t* h) {
  if (!m || !w || !h) return TF_ERR_INVALID;
  *w = m->patch_w;
  *h = m->patch_h;
  return TF_OK;
}

int tf_atlas_alloc_slot(tf_map* m, tf_chunk_id id, uint64_t* texloc_out) {
  if (!m || !texloc_out || !coord_ok(id.x, id.y, id.z)) return fail(m, TF_ERR_INVALID, "tf_atlas_alloc_slot: bad argument");
  const unsigned long long key = pack_key(id.x, id.y, id.z);
  auto it = m->patches.find(key);
  if (it != m->patches.end()) {
    *texloc_out = it->second;
    return TF_OK;
  }
  // Atlas::AddPatch (Structure/Atlas.cpp:43-64)
  const uint64_t loc = m->loc_next;
  uint64_t x = loc % kAtlasDim, y = loc / kAtlasDim;
  if (x >= (uint64_t)kAtlasDim || y >= (uint64_t)kAtlasDim)
    return fail(m, TF_ERR_ATLAS_FULL, "No enough space for texture storage.");
  if (x + m->patch_w >= (uint64_t)kAtlasDim) {
    x = 0;
    y += m->patch_h;
  } else {
    x += m->patch_w;
  }
  m->loc_next = x + y * kAtlasDim;
  m->patches.emplace(key, loc);
  *texloc_out = loc;
  return TF_OK;
}

static int ensure_atlas(tf_map* m) {
  if (m->atlas) return TF_OK;
  const size_t bytes = (size_t)kAtlasDim * kAtlasDim * 3;
  CUDA_OK(m, cudaMalloc((void**)&m->atlas, bytes));
  CUDA_OK(m, cudaMemsetAsync(m->atlas, 0, bytes, m->stream));  // Structure/Atlas.cpp:35-36
  return TF_OK;
}

int tf_atlas_update(tf_map* m, const tf_patch_desc* patches, int64_t n) {
  if (!m || n < 0 || (n > 0 && !patches)) return fail(m, TF_ERR_INVALID, "tf_atlas_update: bad argument");
  use_device(m);
  if (n == 0) return TF_OK;
  if (int rc = ensure_atlas(m)) return rc;
  // descriptors are assembled in a page-locked arena (a pageable source would be staged by the runtime
  // with an extra copy and a stall) and the device copy of the list is a grow-only arena as well
  if (m->patch_w > kAtlasMaxPW || m->patch_h > kAtlasMaxPH) return fail(m, TF_ERR_INVALID, "tf_atlas_update: slot larger than 96 x 72 (voxels > 20 mm)");
  if (int rc = ensure_arena_bytes(m, m->ar_patch_h, (size_t)n * sizeof(PatchDev), true)) return rc;
  if (int rc = ensure_arena_bytes(m, m->ar_patch_d, (size_t)n * sizeof(PatchDev))) return rc;
  if (m->patch_copy_done) cudaEventSynchronize(m->patch_copy_done);  // the previous call's H2D copy has left the host arena
  PatchDev* pd = (PatchDev*)m->ar_patch_h.p;
  int last_fi = -1, s = -1;  // patches arrive grouped by key-frame: one frame-store lookup per run
  for (int64_t i = 0; i < n; i++) {
    const tf_patch_desc& p = patches[i];
    if (p.frame_index != last_fi || s < 0) {
      s = find_slot(m, p.frame_index);
      last_fi = p.frame_index;
    }
    if (s < 0 || !m->slots[s].has_rgb) return fail(m, TF_ERR_NOT_FOUND, "tf_atlas_update: key-frame rgb not in the frame store");
    if (p.texloc >= (uint64_t)kAtlasDim * kAtlasDim) return fail(m, TF_ERR_INVALID, "tf_atlas_update: texloc outside the atlas");
    if (p.w <= 0 || p.h <= 0 || p.x < 0 || p.y < 0 || p.x + p.w > m->W || p.y + p.h > m->H)
      return fail(m, TF_ERR_INVALID, "tf_atlas_update: bbox outside the image");
    const int ox = (int)(p.texloc % kAtlasDim), oy = (int)(p.texloc / kAtlasDim);
    const bool shrink = p.w > m->patch_w || p.h > m->patch_h;
    const int ew = shrink ? m->patch_w : p.w, eh = shrink ? m->patch_h : p.h;
    if (ox + ew > kAtlasDim || oy + eh > kAtlasDim) return fail(m, TF_ERR_INVALID, "tf_atlas_update: slot outside the atlas");
    pd[i] = PatchDev{(unsigned)p.texloc, (unsigned short)s, (unsigned short)p.x, (unsigned short)p.y, (unsigned short)p.w,
                     (unsigned short)p.h, 0};
  }
  CUDA_OK(m, cudaMemcpyAsync(m->ar_patch_d.p, pd, (size_t)n * sizeof(PatchDev), cudaMemcpyHostToDevice, m->stream));
  if (!m->patch_copy_done) CUDA_OK(m, cudaEventCreateWithFlags(&m->patch_copy_done, cudaEventDisableTiming));
  CUDA_OK(m, cudaEventRecord(m->patch_copy_done, m->stream));
  const int grid = (int)std::min<int64_t>((n + kWarpsPerBlock - 1) / kWarpsPerBlock, (int64_t)m->sm_count * 8);
  atlas_update_kernel<<<grid, kThreads, 0, m->stream>>>((const PatchDev*)m->ar_patch_d.p, (int)n, m->atlas, m->slots[0].rgb,
                                                        m->slot_stride, m->W, m->H, m->patch_w, m->patch_h);
  if (int rc = check_kernel(m, "atlas_update_kernel")) return rc;
  m->counters.h2d_bytes += n * (int64_t)sizeof(PatchDev);
  return TF_OK;
}

int tf_atlas_download(tf_map* m, uint64_t hot_start, uint64_t hot_end, uint8_t* rgb_out) {
  if (!m || !rgb_out || hot_end < hot_start || hot_end > (uint64_t)kAtlasDim * kAtlasDim)
    return fail(m, TF_ERR_INVALID, "tf_atlas_download: bad range");
  use_device(m);
  if (int rc = ensure_atlas(m)) return rc;
  const size_t nb = (size_t)(hot_end - hot_start) * 3;
  CUDA_OK(m, cudaMemcpyAsync(rgb_out, m->atlas + hot_start * 3, nb, cudaMemcpyDeviceToHost, m->stream));
  CUDA_OK(m, cudaStreamSynchronize(m->stream));
  m->counters.d2h_bytes += nb;
  return TF_OK;
}

// The same rows device to device: the caller's pixel-unpack buffer, registered with CUDA
// (cudaGraphicsGLRegisterBuffer) and mapped, takes the place of glBufferDataARB's host pointer — the
// texture bytes never leave the GPU.
int tf_atlas_copy_to_device(tf_map* m, uint64_t hot_start, uint64_t hot_end, void* dst_device) {
  if (!m || !dst_device || hot_end < hot_start || hot_end > (uint64_t)kAtlasDim * kAtlasDim)
    return fail(m, TF_ERR_INVALID, "tf_atlas_copy_to_device: bad range");
  use_device(m);
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, dst_device) != cudaSuccess || at.type != cudaMemoryTypeDevice) {
    cudaGetLastError();
    return fail(m, TF_ERR_INVALID, "tf_atlas_copy_to_device: destination is not device memory");
  }
  if (int rc = ensure_atlas(m)) return rc;
  const size_t nb = (size_t)(hot_end - hot_start) * 3;
  if (at.device == m->cfg.device)
    CUDA_OK(m, cudaMemcpyAsync(dst_device, m->atlas + hot_start * 3, nb, cudaMemcpyDeviceToDevice, m->stream));
  else  // the display GPU is another device than the map's
    CUDA_OK(m, cudaMemcpyPeerAsync(dst_device, at.device, m->atlas + hot_start * 3, m->cfg.device, nb, m->stream));
  CUDA_OK(m, cudaStreamSynchronize(m->stream));  // the caller unmaps the resource right after
  return TF_OK;
}

// Patch::CalculateTexCoords for a batch of chunk meshes against one key-frame of the store.
int tf_patch_texcoords(tf_map* m, int32_t frame_index, const tf_pose* world_to_camera, const tf_camera* cam,
                       int64_t n_patches, const int64_t* vertex_offsets, const float* vertices, const float* colors,
                       float* texcoord_out, float* texcolor_out, tf_patch_result* results) {
  if (!m || !world_to_camera || !cam_ok(m, cam) || n_patches < 0 ||
      (n_patches > 0 && (!vertex_offsets || !vertices || !colors || !texcoord_out || !texcolor_out || !results)))
    return fail(m, TF_ERR_INVALID, "tf_patch_texcoords: bad argument");
  use_device(m);
  if (n_patches == 0) return TF_OK;
  const int s = find_slot(m, frame_index);
  if (s < 0 || !m->slots[s].has_rgb) return fail(m, TF_ERR_NOT_FOUND, "tf_patch_texcoords: key-frame rgb not in the frame store");
  const int64_t nv = vertex_offsets[n_patches];
  if (nv < 0 || vertex_offsets[0] != 0) return fail(m, TF_ERR_INVALID, "tf_patch_texcoords: bad offsets");
  // grow-only arenas: no cudaMalloc / cudaFree per call
  const size_t nvv = (size_t)std::max<int64_t>(nv, 1);
  if (int rc = ensure_arena_bytes(m, m->ar_tc_off, ((size_t)n_patches + 1) * 8)) return rc;
  if (int rc = ensure_arena_bytes(m, m->ar_tc_v, nvv * 12)) return rc;
  if (int rc = ensure_arena_bytes(m, m->ar_tc_c, nvv * 12)) return rc;
  if (int rc = ensure_arena_bytes(m, m->ar_tc_tc, nvv * 8)) return rc;
  if (int rc = ensure_arena_bytes(m, m->ar_tc_col, nvv * 12)) return rc;
  if (int rc = ensure_arena_bytes(m, m->ar_tc_res, (size_t)n_patches * sizeof(PatchTexResult))) return rc;
  long long* d_off = (long long*)m->ar_tc_off.p;
  float *d_v = (float*)m->ar_tc_v.p, *d_c = (float*)m->ar_tc_c.p, *d_tc = (float*)m->ar_tc_tc.p, *d_col = (float*)m->ar_tc_col.p;
  PatchTexResult* d_res = (PatchTexResult*)m->ar_tc_res.p;
  CUDA_OK(m, cudaMemcpyAsync(d_off, vertex_offsets, (n_patches + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, m->stream));
  CUDA_OK(m, cudaMemcpyAsync(d_v, vertices, nv * 12, cudaMemcpyHostToDevice, m->stream));
  CUDA_OK(m, cudaMemcpyAsync(d_c, colors, nv * 12, cudaMemcpyHostToDevice, m->stream));
  tf_pose_dev T;
  memcpy(T.m, world_to_camera->m, sizeof(T.m));
  patch_texcoords_kernel<<<(unsigned)n_patches, kPatchThreads, 0, m->stream>>>(
      m->slots[s].rgb, m->slots[s].depth, T, (float)(int)cam->fx, (float)(int)cam->fy, (float)(int)cam->cx, (float)(int)cam->cy,
      m->W, m->H, d_off, d_v, d_c, d_tc, d_col, d_res);
  if (int rc = check_kernel(m, "patch_texcoords_kernel")) return rc;
  CUDA_OK(m, cudaMemcpyAsync(texcoord_out, d_tc, nv * 8, cudaMemcpyDeviceToHost, m->stream));
  CUDA_OK(m, cudaMemcpyAsync(texcolor_out, d_col, nv * 12, cudaMemcpyDeviceToHost, m->stream));
  CUDA_OK(m, cudaMemcpyAsync(results, d_res, n_patches * sizeof(PatchTexResult), cudaMemcpyDeviceToHost, m->stream));
  CUDA_OK(m, cudaStreamSynchronize(m->stream));
  m->counters.h2d_bytes += nv * 24 + (n_patches + 1) * 8;
  m->counters.d2h_bytes += nv * 20 + n_patches * (int64_t)sizeof(PatchTexResult);
  return TF_OK;
}

// ---- frame pre-processing (kernels in tf_pre.cuh) -----------------------------------
//
// The loops of main.cpp:117-147 on the planes of the frame store.  They are part of a frame's INGEST:
// everything runs on the copy stream, behind the frame's upload, and re-arms the slot's `ready` event,
// so the fusion kernels see the refined planes exactly as they would see an upload.

namespace {

bool pre_cam_ok(const tf_map* m, const tf_camera* c) {
  return c && c->width == m->W && c->height == m->H && (m->W % 8) == 0 && c->fx != 0.0f && c->fy != 0.0f;
}
PreCam pre_cam(const tf_map* m, const tf_camera* c) { return PreCam{c->fx, c->fy, c->cx, c->cy, m->W, m->H, m->cfg.dot3_order}; }
PreXf pre_xf(const float* T) {
  PreXf x;
  for (int a = 0; a < 3; a++) {
    for (int b = 0; b < 3; b++) x.r[a][b] = T[a * 4 + b];
    x.t[a] = T[a * 4 + 3];
  }
  return x;
}

// Side planes of a frame (created on first use: weights 0, no normal map).
tf_map::PreEntry* pre_entry(tf_map* m, int frame_index, bool create) {
  tf_map::PreEntry* lru = nullptr;  // a free entry, else the least recently used one
  for (auto& e : m->pre) {
    if (e.frame_index == frame_index) { e.last_use = ++m->use_clock; return &e; }
    const bool e_free = e.frame_index < 0, l_free = lru && lru->frame_index < 0;
    if (!lru || (e_free && !l_free) || (e_free == l_free && e.last_use < lru->last_use)) lru = &e;
  }
  if (!create) return nullptr;
  if (!lru->normal) {
    if (cudaMalloc(&lru->normal, (size_t)m->npix * 12) != cudaSuccess || cudaMalloc(&lru->weight, (size_t)m->npix * 4) != cudaSuccess) {
      fail(m, TF_ERR_CUDA, "tf_pre: out of device memory");
      return nullptr;
    }
  }
  cudaMemsetAsync(lru->weight, 0, (size_t)m->npix * 4, m->copy_stream);
  lru->frame_index = frame_index, lru->has_normal = false, lru->last_use = ++m->use_clock;
  return lru;
}

int pre_scratch(tf_map* m) {
  if (m->pre_snap) return TF_OK;
  CUDA_OK(m, cudaMalloc(&m->pre_snap, (size_t)m->npix * 4));
  CUDA_OK(m, cudaMalloc(&m->pre_bil, sizeof(BilateralState)));
  CUDA_OK(m, cudaMalloc(&m->pre_queue, ((size_t)m->npix + 2) * 4));
  CUDA_OK(m, cudaMemsetAsync(m->pre_queue + m->npix, 0, 8, m->copy_stream));
  CUDA_OK(m, cudaMalloc(&m->pre_waiting, (size_t)m->npix));
  return TF_OK;
}

// The slot of a frame whose planes the copy stream is about to read (modify = false) or rewrite.
int pre_slot(tf_map* m, int32_t frame_index, bool modify) {
  const int s = find_slot(m, frame_index, false);
  if (s < 0) { fail(m, TF_ERR_NOT_FOUND, "tf_pre: frame_index not in the frame store (tf_upload_frame first)"); return -1; }
  if (modify) guard_overwrite(m, m->slots[s]);
  return s;
}
int pre_publish(tf_map* m, FrameSlot& fsl) {  // the slot's planes changed: consumers wait for the copy stream again
  CUDA_OK(m, cudaEventRecord(fsl.ready, m->copy_stream));
  fsl.pending = true;
  return TF_OK;
}

}  // namespace

int tf_pre_upload_depth_u16(tf_map* m, int32_t frame_index, const uint16_t* depth, float depth_scale, float max_depth) {
  if (!m || !depth || frame_index < 0 || !(depth_scale > 0.0f)) return fail(m, TF_ERR_INVALID, "tf_pre_upload_depth_u16: bad argument");
  use_device(m);
  if (int rc = pre_scratch(m)) return rc;
  const int s = acquire_slot(m, frame_index);
  if (s < 0) return TF_ERR_CAPACITY;
  FrameSlot& fsl = m->slots[s];
  guard_overwrite(m, fsl);
  fsl.has_rgba = fsl.has_quality = fsl.has_rgb = false;
  for (auto& e : m->pre)
    if (e.frame_index == frame_index) e.frame_index = -1;
  CUDA_OK(m, cudaMemcpyAsync(m->pre_snap, depth, (size_t)m->npix * 2, cudaMemcpyHostToDevice, m->copy_stream));
  m->counters.h2d_bytes += (int64_t)m->npix * 2;
  pre_depth_u16_kernel<<<m->grid, 256, 0, m->copy_stream>>>((const unsigned short*)m->pre_snap, fsl.depth, m->npix, depth_scale, max_depth);
  if (int rc = check_kernel(m, "pre_depth_u16_kernel")) return rc;
  return pre_publish(m, fsl);
}

int tf_pre_bilateral(tf_map* m, int32_t frame_index, int32_t d, float sigma_color, float sigma_space) {
  if (!m) return TF_ERR_INVALID;
  use_device(m);
  // cv::bilateralFilter's parameter rules (sigma <= 0 -> 1; d <= 0 -> radius from sigma_space; radius >= 1)
  if (sigma_color <= 0.0f) sigma_color = 1.0f;
  if (sigma_space <= 0.0f) sigma_space = 1.0f;
  int radius = d <= 0 ? (int)lrint(sigma_space * 1.5) : d / 2;
  radius = std::max(radius, 1);
  if (radius > kBilMaxRadius) return fail(m, TF_ERR_INVALID, "tf_pre_bilateral: diameter above 17 not supported");
  if (int rc = pre_scratch(m)) return rc;
  const int s = pre_slot(m, frame_index, true);
  if (s < 0) return TF_ERR_NOT_FOUND;
  FrameSlot& fsl = m->slots[s];
  const double gc = -0.5 / ((double)sigma_color * sigma_color), gs = -0.5 / ((double)sigma_space * sigma_space);
  static const int init[2] = {0x7fffffff, (int)0x80000000};  // (static: the copy is asynchronous)
  CUDA_OK(m, cudaMemcpyAsync(m->pre_bil, init, 8, cudaMemcpyHostToDevice, m->copy_stream));
  CUDA_OK(m, cudaMemcpyAsync(m->pre_snap, fsl.depth, (size_t)m->npix * 4, cudaMemcpyDeviceToDevice, m->copy_stream));
  pre_minmax_kernel<<<m->grid, 256, 0, m->copy_stream>>>(m->pre_snap, m->npix, m->pre_bil);
  if (int rc = check_kernel(m, "pre_minmax_kernel")) return rc;
  pre_bilateral_lut_kernel<<<1, 1024, 0, m->copy_stream>>>(m->pre_bil, gc);
  if (int rc = check_kernel(m, "pre_bilateral_lut_kernel")) return rc;
  pre_bilateral_kernel<<<dim3((m->W + 31) / 32, (m->H + 7) / 8), 256, 0, m->copy_stream>>>(m->pre_snap, fsl.depth, m->W, m->H, radius, gs, m->pre_bil);
  if (int rc = check_kernel(m, "pre_bilateral_kernel")) return rc;
  return pre_publish(m, fsl);
}

int tf_pre_normal_map(tf_map* m, int32_t frame_index, const tf_camera* cam) {
  if (!m || !pre_cam_ok(m, cam)) return fail(m, TF_ERR_INVALID, "tf_pre_normal_map: bad argument");
  use_device(m);
  const int s = pre_slot(m, frame_index, false);
  if (s < 0) return TF_ERR_NOT_FOUND;
  tf_map::PreEntry* e = pre_entry(m, frame_index, true);
  if (!e) return TF_ERR_CUDA;
  pre_normal_kernel<<<m->grid, 256, 0, m->copy_stream>>>(pre_cam(m, cam), m->slots[s].depth, e->normal);
  e->has_normal = true;
  return check_kernel(m, "pre_normal_kernel");
}

int tf_pre_refine_depth_by_normal(tf_map* m, int32_t frame_index, const tf_camera* cam) {
  if (!m || !pre_cam_ok(m, cam)) return fail(m, TF_ERR_INVALID, "tf_pre_refine_depth_by_normal: bad argument");
  use_device(m);
  tf_map::PreEntry* e = pre_entry(m, frame_index, false);
  if (!e || !e->has_normal) return fail(m, TF_ERR_NOT_FOUND, "tf_pre_refine_depth_by_normal: no normal map (tf_pre_normal_map first)");
  const int s = pre_slot(m, frame_index, true);
  if (s < 0) return TF_ERR_NOT_FOUND;
  pre_grazing_kernel<<<m->grid, 256, 0, m->copy_stream>>>(pre_cam(m, cam), e->normal, m->slots[s].depth);
  if (int rc = check_kernel(m, "pre_grazing_kernel")) return rc;
  return pre_publish(m, m->slots[s]);
}

int tf_pre_refine_keyframe(tf_map* m, int32_t keyframe_index, int32_t new_index, const float* ref_to_new, const tf_camera* cam) {
  if (!m || !ref_to_new || !pre_cam_ok(m, cam) || keyframe_index == new_index)
    return fail(m, TF_ERR_INVALID, "tf_pre_refine_keyframe: bad argument");
  use_device(m);
  const int sn = pre_slot(m, new_index, false);
  const int sk = sn < 0 ? -1 : pre_slot(m, keyframe_index, true);
  if (sk < 0) return TF_ERR_NOT_FOUND;
  tf_map::PreEntry* e = pre_entry(m, keyframe_index, true);
  if (!e) return TF_ERR_CUDA;
  if (int rc = pre_scratch(m)) return rc;
  RefineKfArgs a;
  a.c = pre_cam(m, cam);
  a.x = pre_xf(ref_to_new);
  a.kf_d = m->pre_snap, a.kf_w = e->weight, a.new_d = m->slots[sn].depth;
  a.out_d = m->slots[sk].depth, a.out_w = e->weight;
  a.queue = m->pre_queue, a.queue_n = m->pre_queue + m->npix, a.ticket = (unsigned*)(m->pre_queue + m->npix + 1);
  a.waiting = m->pre_waiting;
  CUDA_OK(m, cudaMemcpyAsync(m->pre_snap, m->slots[sk].depth, (size_t)m->npix * 4, cudaMemcpyDeviceToDevice, m->copy_stream));
  CUDA_OK(m, cudaMemsetAsync(a.queue_n, 0, 4, m->copy_stream));
  pre_refine_kf_kernel<<<m->grid, 256, 0, m->copy_stream>>>(a);
  if (int rc = check_kernel(m, "pre_refine_kf_kernel")) return rc;
  return pre_publish(m, m->slots[sk]);
}

int tf_pre_refine_newframe(tf_map* m, int32_t keyframe_index, int32_t new_index, const float* new_to_ref, const tf_camera* cam) {
  if (!m || !new_to_ref || !pre_cam_ok(m, cam) || keyframe_index == new_index)
    return fail(m, TF_ERR_INVALID, "tf_pre_refine_newframe: bad argument");
  use_device(m);
  const int sk = pre_slot(m, keyframe_index, false);
  const int sn = sk < 0 ? -1 : pre_slot(m, new_index, true);
  if (sn < 0) return TF_ERR_NOT_FOUND;
  pre_refine_new_kernel<<<m->grid, 256, 0, m->copy_stream>>>(pre_cam(m, cam), pre_xf(new_to_ref), m->slots[sk].depth, m->slots[sn].depth);
  if (int rc = check_kernel(m, "pre_refine_new_kernel")) return rc;
  return pre_publish(m, m->slots[sn]);
}

int tf_pre_color_quality(tf_map* m, int32_t frame_index, const uint8_t* rgb, const tf_camera* cam) {
  if (!m || !rgb || !pre_cam_ok(m, cam)) return fail(m, TF_ERR_INVALID, "tf_pre_color_quality: bad argument");
  use_device(m);
  tf_map::PreEntry* e = pre_entry(m, frame_index, false);
  if (!e || !e->has_normal) return fail(m, TF_ERR_NOT_FOUND, "tf_pre_color_quality: no normal map (tf_pre_normal_map first)");
  const int s = pre_slot(m, frame_index, true);
  if (s < 0) return TF_ERR_NOT_FOUND;
  FrameSlot& fsl = m->slots[s];
  if (int rc = ensure_color_planes(m, fsl)) return rc;
  fsl.pinned = true;  // as tf_upload_keyframe_rgb: the key-frame's rgb stays for the atlas until tf_release_frame
  CUDA_OK(m, cudaMemcpyAsync(fsl.rgb, rgb, (size_t)m->npix * 3, cudaMemcpyHostToDevice, m->copy_stream));
  m->counters.h2d_bytes += (int64_t)m->npix * 3;
  pre_color_kernel<<<m->grid, 256, 0, m->copy_stream>>>(pre_cam(m, cam), fsl.depth, e->normal, fsl.rgb, fsl.valid, fsl.quality, fsl.rgba);
  if (int rc = check_kernel(m, "pre_color_kernel")) return rc;
  fsl.has_rgb = fsl.has_rgba = fsl.has_quality = true;
  return pre_publish(m, fsl);
}

int tf_pre_download(tf_map* m, int32_t frame_index, float* depth, float* normal, float* weight, uint8_t* color_valid, float* quality) {
  if (!m) return TF_ERR_INVALID;
  use_device(m);
  const int s = pre_slot(m, frame_index, false);
  if (s < 0) return TF_ERR_NOT_FOUND;
  FrameSlot& fsl = m->slots[s];
  tf_map::PreEntry* e = pre_entry(m, frame_index, false);
  if ((normal && !(e && e->has_normal)) || (weight && !e)) return fail(m, TF_ERR_NOT_FOUND, "tf_pre_download: plane not computed for this frame");
  if ((color_valid || quality) && !(fsl.has_rgb && fsl.has_quality)) return fail(m, TF_ERR_NOT_FOUND, "tf_pre_download: no colour planes for this frame");
  const size_t n = (size_t)m->npix;
  if (depth) CUDA_OK(m, cudaMemcpyAsync(depth, fsl.depth, n * 4, cudaMemcpyDeviceToHost, m->copy_stream));
  if (normal) CUDA_OK(m, cudaMemcpyAsync(normal, e->normal, n * 12, cudaMemcpyDeviceToHost, m->copy_stream));
  if (weight) CUDA_OK(m, cudaMemcpyAsync(weight, e->weight, n * 4, cudaMemcpyDeviceToHost, m->copy_stream));
  if (color_valid) CUDA_OK(m, cudaMemcpyAsync(color_valid, fsl.valid, n, cudaMemcpyDeviceToHost, m->copy_stream));
  if (quality) CUDA_OK(m, cudaMemcpyAsync(quality, fsl.quality, n * 4, cudaMemcpyDeviceToHost, m->copy_stream));
  CUDA_OK(m, cudaStreamSynchronize(m->copy_stream));
  m->counters.d2h_bytes += (int64_t)n * ((depth ? 4 : 0) + (normal ? 12 : 0) + (weight ? 4 : 0) + (color_valid ? 1 : 0) + (quality ? 4 : 0));
  return TF_OK;
}

// ---- counters / profiling ------------------------------------------------------------------------

int tf_get_counters(tf_map* m, tf_counters* out) {
  if (!m || !out) return TF_ERR_INVALID;
  m->counters.pool_used = m->n_live;
  *out = m->counters;
  return TF_OK;
}

int tf_set_profiling(tf_map* m, int enable) {
  if (!m) return TF_ERR_INVALID;
  m->prof = enable;
  return TF_OK;
}

int tf_debug_project(tf_map* m, const float* c, const float* cz, int64_t n, float f, float ch, int32_t* u_fast,
                     int32_t* u_exact, uint8_t* accepted) {
  if (!m || !c || !cz || !u_fast || !u_exact || !accepted || n <= 0 || n > (1 << 26))
    return fail(m, TF_ERR_INVALID, "tf_debug_project: bad argument");
  use_device(m);
  float *dc = nullptr, *dz = nullptr;
  int *duf = nullptr, *due = nullptr;
  unsigned char* da = nullptr;
  int rc = TF_OK;
  auto ok = [&](cudaError_t e) {
    if (e != cudaSuccess && rc == TF_OK) rc = fail(m, TF_ERR_CUDA, cudaGetErrorString(e));
    return e == cudaSuccess;
  };
  if (ok(dmalloc(&dc, (size_t)n)) && ok(dmalloc(&dz, (size_t)n)) && ok(dmalloc(&duf, (size_t)n)) &&
      ok(dmalloc(&due, (size_t)n)) && ok(dmalloc(&da, (size_t)n)) &&
      ok(cudaMemcpyAsync(dc, c, n * 4, cudaMemcpyHostToDevice, m->stream)) &&
      ok(cudaMemcpyAsync(dz, cz, n * 4, cudaMemcpyHostToDevice, m->stream))) {
    debug_project_kernel<<<m->grid, kThreads, 0, m->stream>>>(dc, dz, (int)n, f, ch, duf, due, da);
    ok(cudaGetLastError());
    ok(cudaMemcpyAsync(u_fast, duf, n * 4, cudaMemcpyDeviceToHost, m->stream));
    ok(cudaMemcpyAsync(u_exact, due, n * 4, cudaMemcpyDeviceToHost, m->stream));
    ok(cudaMemcpyAsync(accepted, da, n, cudaMemcpyDeviceToHost, m->stream));
    ok(cudaStreamSynchronize(m->stream));
  }
  cudaFree(dc); cudaFree(dz); cudaFree(duf); cudaFree(due); cudaFree(da);
  return rc;
}

int tf_debug_divide(tf_map* m, const float* num, const float* den, int64_t n, float* q_kernel, float* q_ieee,
                    uint8_t* accepted) {
  if (!m || !num || !den || !q_kernel || !q_ieee || !accepted || n <= 0 || n > (1 << 26))
    return fail(m, TF_ERR_INVALID, "tf_debug_divide: bad argument");
  use_device(m);
  float *dn = nullptr, *dd = nullptr, *dk = nullptr, *di = nullptr;
  unsigned char* da = nullptr;
  int rc = TF_OK;
  auto ok = [&](cudaError_t e) {
    if (e != cudaSuccess && rc == TF_OK) rc = fail(m, TF_ERR_CUDA, cudaGetErrorString(e));
    return e == cudaSuccess;
  };
  if (ok(dmalloc(&dn, (size_t)n)) && ok(dmalloc(&dd, (size_t)n)) && ok(dmalloc(&dk, (size_t)n)) &&
      ok(dmalloc(&di, (size_t)n)) && ok(dmalloc(&da, (size_t)n)) &&
      ok(cudaMemcpyAsync(dn, num, n * 4, cudaMemcpyHostToDevice, m->stream)) &&
      ok(cudaMemcpyAsync(dd, den, n * 4, cudaMemcpyHostToDevice, m->stream))) {
    debug_divide_kernel<<<m->grid, kThreads, 0, m->stream>>>(dn, dd, (int)n, dk, di, da);
    ok(cudaGetLastError());
    ok(cudaMemcpyAsync(q_kernel, dk, n * 4, cudaMemcpyDeviceToHost, m->stream));
    ok(cudaMemcpyAsync(q_ieee, di, n * 4, cudaMemcpyDeviceToHost, m->stream));
    ok(cudaMemcpyAsync(accepted, da, n, cudaMemcpyDeviceToHost, m->stream));
    ok(cudaStreamSynchronize(m->stream));
  }
  cudaFree(dn); cudaFree(dd); cudaFree(dk); cudaFree(di); cudaFree(da);
  return rc;
}

int tf_get_stage_times(tf_map* m, int reset, double* ms6) {
  if (!m || !ms6) return TF_ERR_INVALID;
  CUDA_OK(m, cudaStreamSynchronize(m->stream));
  prof_collect(m, 0);
  for (int i = 0; i < kStages; i++) {
    ms6[i] = m->stage_ms[i];
    if (reset) m->stage_ms[i] = 0;
  }
  return TF_OK;
}

int tf_get_kernel_time(tf_map* m, int reset, double* integrate_ms, int64_t* launches, double* bytes) {
  if (!m) return TF_ERR_INVALID;
  CUDA_OK(m, cudaStreamSynchronize(m->stream));
  prof_collect(m, 0);
  if (integrate_ms) *integrate_ms = m->prof_ms;
  if (launches) *launches = m->prof_launches;
  if (bytes) *bytes = m->prof_bytes;
  if (reset) {
    m->prof_ms = 0;
    m->prof_bytes = 0;
    m->prof_launches = 0;
  }
  return TF_OK;
}

}  // extern "C"
