// tf_mesh.cuh — marching-cubes meshing of chunks in place.
//
//   mesh_kernel<kWrite>   ChunkManager::GenerateMeshEfficient           (Structure/ChunkManager.cpp:595-1002)
//                         + extractGradientFromCubic / GetNeighborSDF    (:277-455, Structure/ChunkManager.h:790-830)
//                         per chunk of RecomputeMeshes                   (:232-264)
//   mesh_scan_kernel      per-chunk vertex / index counts -> offsets
//
// One block of 512 threads per chunk, thread = voxel in the reference's loop order
// (voxel index (z*8+y)*8+x, Chunk.h:91-93).  What the serial reference does with running state is
// restated as order-independent rules:
//   * vertices live in per-chunk edge slots (3 x 9^3, :905-932); a slot written by several voxels
//     keeps the LAST writer in loop order -> atomicMax of the voxel index per slot, the winner writes;
//   * the final vertex order is the slot order of the used slots (:965-975) -> block scan of the used flags;
//   * the index list is in voxel order, then table order -> block scan of the per-voxel triangle counts.
// The kernel runs twice: kWrite = false counts vertices and indices per chunk, mesh_scan_kernel turns
// the counts into offsets, kWrite = true writes.  Voxels are read where they lie (chunk pool, through
// the hash for the 4 x 4 x 4 neighbourhood the gradients can reach); nothing is copied to the host
// except the mesh.  Arithmetic: the reference's float operations one by one (no FMA; -fmad=false).
#pragma once
#include "tf_device.cuh"
#include "tf_mc_table.h"

namespace tfb {

constexpr int kMeshThreads = 512;
constexpr int kMeshSlots = 3 * 729;  // vertByEdge (Structure/ChunkManager.cpp:646)

struct MeshArgs {
  const int3* ids;
  int n;
  float res;
  int l2r;               // association of Eigen's squaredNorm (tf_config.dot3_order)
  int2* counts;          // [n] vertices, indices per chunk (written by the counting pass)
  const long long* off;  // [2 * (n + 1)] vertex offsets, then index offsets (writing pass)
  float* vert;
  float* norm;
  float* col;
  int* idx;
};

// block-wide exclusive scan over kMeshThreads threads
__device__ __forceinline__ int mesh_block_scan(int v, int* total, int* scratch /*[17]*/) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int incl = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int n = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += n;
  }
  if (lane == 31) scratch[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    const int ws = lane < kMeshThreads / 32 ? scratch[lane] : 0;
    int wincl = ws;
#pragma unroll
    for (int d = 1; d < kMeshThreads / 32; d <<= 1) {
      const int n = __shfl_up_sync(0xffffffffu, wincl, d);
      if (lane >= d) wincl += n;
    }
    if (lane < kMeshThreads / 32) scratch[lane] = wincl - ws;
    if (lane == kMeshThreads / 32 - 1) scratch[16] = wincl;
  }
  __syncthreads();
  const int res = incl - v + scratch[wid];
  *total = scratch[16];
  __syncthreads();
  return res;
}

// voxel planes of a chunk given its hash value (slot | lazy bit): a chunk that was created but never
// written reads as its initial state (Chunk.cpp:60-68, ColorVoxel.cpp:26-31)
__device__ __forceinline__ float mesh_sdf(const MapDev& md, int entry, int vi) {
  return (entry & kLazyBit) ? 999.0f : reinterpret_cast<const float*>(md.pool + (size_t)(entry & (kLazyBit - 1)) * kChunkBytes + kSdfOff)[vi];
}
__device__ __forceinline__ float mesh_weight(const MapDev& md, int entry, int vi) {
  return (entry & kLazyBit) ? 0.0f : reinterpret_cast<const float*>(md.pool + (size_t)(entry & (kLazyBit - 1)) * kChunkBytes + kWeightOff)[vi];
}
__device__ __forceinline__ uint2 mesh_color(const MapDev& md, int entry, int vi) {
  return (entry & kLazyBit) ? make_uint2(0u, 0u)
                            : reinterpret_cast<const uint2*>(md.pool + (size_t)(entry & (kLazyBit - 1)) * kChunkBytes + kColorOff)[vi];
}

// cubeIndexOffsets (Structure/ChunkManager.cpp:64-65): corner k of a cube = voxel + (x, y, z)
__device__ __forceinline__ int mc_cx(int k) { return (0x66 >> k) & 1; }  // 0 1 1 0 0 1 1 0
__device__ __forceinline__ int mc_cy(int k) { return (0xcc >> k) & 1; }  // 0 0 1 1 0 0 1 1
__device__ __forceinline__ int mc_cz(int k) { return k >> 2; }           // 0 0 0 0 1 1 1 1
// the corner reached from corner k by flipping one axis
__device__ __forceinline__ int mc_corner(int x, int y, int z) { return z * 4 + (y ? 3 - x : x); }

// edge slot of edge s of the cube at voxel (x, y, z) (Structure/ChunkManager.cpp:905-932)
__device__ __forceinline__ int mc_slot(int x, int y, int z, int s) {
  const int b = (x + ((0x622 >> s) & 1)) + (y + ((0xc44 >> s) & 1)) * 9 + (z + ((0x0f0 >> s) & 1)) * 81;
  const int a = s >= 8 ? 2 : (s & 1);
  return a + b * 3;
}

struct McVoxel {       // what a thread keeps about its voxel between the phases
  float s[8], w[8];    // corner sdf / weight
  int x, y, z;
  bool meshed;         // all corners observed, sign change inside the cube
};

// One crossing edge of the voxel's cube: vertex (relative to the voxel centroid), normal, colour and
// normalValidFlag (Structure/ChunkManager.cpp:758-860).  nb: the 4 x 4 x 4 table of hash values.
__device__ __forceinline__ bool mc_edge(const MapDev& md, const int* nb, const McVoxel& v, int e, float res, int l2r,
                                        float3* pos, float3* nrm, float3* col) {
  const int c0 = kMcEdgeCorners[e] & 15, c1 = kMcEdgeCorners[e] >> 4;
  const float sdf0 = v.s[c0], sdf1 = v.s[c1];
  if (!(__fmul_rn(sdf0, sdf1) < 0.0f)) return false;  // only edges with a zero crossing (:768)
  const float t = __fdiv_rn(sdf0, __fsub_rn(sdf0, sdf1));
  // cubeCoordOffsets.col(c) = float(offset) * res; vertex = a + t * (b - a), per coefficient (:770-773)
  const float ax = __fmul_rn((float)mc_cx(c0), res), ay = __fmul_rn((float)mc_cy(c0), res), az = __fmul_rn((float)mc_cz(c0), res);
  const float bx = __fmul_rn((float)mc_cx(c1), res), by = __fmul_rn((float)mc_cy(c1), res), bz = __fmul_rn((float)mc_cz(c1), res);
  *pos = make_float3(__fadd_rn(ax, __fmul_rn(t, __fsub_rn(bx, ax))), __fadd_rn(ay, __fmul_rn(t, __fsub_rn(by, ay))),
                     __fadd_rn(az, __fmul_rn(t, __fsub_rn(bz, az))));
  // the corner with the smaller |sdf| supplies normal and colour (:775)
  const int cid = fabsf(sdf0) > fabsf(sdf1) ? c1 : c0;
  const int kx = mc_cx(cid), ky = mc_cy(cid), kz = mc_cz(cid);
  const int p[3] = {v.x + kx, v.y + ky, v.z + kz};                // corner position in the chunk's frame, 0..8
  const int cc[3] = {p[0] == 8, p[1] == 8, p[2] == 8};            // which of the 2 x 2 x 2 chunks holds it
  const int l[3] = {p[0] & 7, p[1] & 7, p[2] & 7};                // its voxel there
  const int centre = nb[(cc[0] + 1) + 4 * (cc[1] + 1) + 16 * (cc[2] + 1)];
  // extractGradientFromCubic (:277-455): central differences around the corner; towards the inside of
  // the cube the neighbour is a cube corner, towards the outside it is looked up (GetNeighborSDF)
  float dd[6];
  bool ok = true;
#pragma unroll
  for (int axis = 0; axis < 3; axis++) {
    const int bit = axis == 0 ? kx : axis == 1 ? ky : kz;
    const int flipped = mc_corner(axis == 0 ? 1 - kx : kx, axis == 1 ? 1 - ky : ky, axis == 2 ? 1 - kz : kz);
    dd[2 * axis + (1 - bit)] = v.s[flipped];
    const int k = 2 * axis + bit;                 // the direction that leaves the cube
    const int sign = bit ? 1 : -1;
    // edgeFlag is taken from the position in the CURRENT chunk's frame (0 or 7), so a corner at 8 —
    // voxel 0 of the next chunk — is not flagged and its "-" neighbour is read at the wrapped index
    // of its own chunk (:293-314): reproduced as is
    const bool edge = bit ? (p[axis] == 7) : (p[axis] == 0);
    int ll[3] = {l[0], l[1], l[2]};
    ll[axis] = (ll[axis] + sign + 8) & 7;         // voxelNeighborIndex: wrap-around (:108-157)
    const int vi = ll[0] + 8 * ll[1] + 64 * ll[2];
    int entry = centre;
    if (edge) {
      int c3[3] = {cc[0], cc[1], cc[2]};
      c3[axis] += sign;
      entry = nb[(c3[0] + 1) + 4 * (c3[1] + 1) + 16 * (c3[2] + 1)];
    }
    float d = 0.0f;
    if (entry < 0) ok = false;                    // chunk not in the map
    else d = mesh_sdf(md, entry, vi);
    if (!(d < 1.0f)) ok = false;
    dd[k] = d;
  }
  const float g0 = __fsub_rn(dd[1], dd[0]), g1 = __fsub_rn(dd[3], dd[2]), g2 = __fsub_rn(dd[5], dd[4]);
  const float q0 = __fmul_rn(g0, g0), q1 = __fmul_rn(g1, g1), q2 = __fmul_rn(g2, g2);
  const float z = l2r ? __fadd_rn(__fadd_rn(q0, q1), q2) : __fadd_rn(q0, __fadd_rn(q1, q2));  // squaredNorm (Eigen redux)
  const float g = __fsqrt_rn(z);
  float n0 = g0, n1 = g1, n2 = g2;
  if (l2r || z > 0.0f) {  // normalize(): Eigen >= 3.3 leaves a zero vector alone, 3.2 divides
    n0 = __fdiv_rn(g0, g), n1 = __fdiv_rn(g1, g), n2 = __fdiv_rn(g2, g);
  }
  if (g > __fmul_rn(res, 100.0f)) ok = false;
  if (!(v.w[cid] > 50.0f)) return false;          // weight_threshold (:793-794): flag stays 0
  *nrm = make_float3(n0, n1, n2);
  const int vi = l[0] + 8 * l[1] + 64 * l[2];
  const uint2 c = mesh_color(md, centre, vi);
  const float cw = (float)(c.y >> 16);
  if (cw > 0.0f)
    *col = make_float3(__fdiv_rn(__fdiv_rn((float)(c.x & 0xffffu), 255.0f), cw), __fdiv_rn(__fdiv_rn((float)(c.x >> 16), 255.0f), cw),
                       __fdiv_rn(__fdiv_rn((float)(c.y & 0xffffu), 255.0f), cw));
  else
    *col = make_float3(1.0f, 1.0f, 1.0f);
  return ok;
}

template <bool kWrite>
__global__ void __launch_bounds__(kMeshThreads) mesh_kernel(const MapDev md, const MeshArgs a) {
  __shared__ int nb[64];               // hash values of the chunks at offsets -1..2 per axis (-1: absent)
  __shared__ int owner[kMeshSlots];    // last voxel (loop order) that writes the slot
  __shared__ int rank[kMeshSlots];     // compacted vertex index of the slot
  __shared__ int scratch[17];
  const int t = threadIdx.x;
  for (int c = blockIdx.x; c < a.n; c += gridDim.x) {
    const int3 id = a.ids[c];
    if (t < 64) {
      const int dx = (t & 3) - 1, dy = ((t >> 2) & 3) - 1, dz = (t >> 4) - 1;
      const int3 q = make_int3(id.x + dx, id.y + dy, id.z + dz);
      nb[t] = coord_ok(q.x, q.y, q.z) ? hash_find(md, pack_key(q.x, q.y, q.z)) : -1;
    }
    for (int s = t; s < kMeshSlots; s += kMeshThreads) owner[s] = -1;
    __syncthreads();
    McVoxel v;
    v.x = t & 7, v.y = (t >> 3) & 7, v.z = t >> 6;
    v.meshed = false;
    int index = 0;
    if (nb[21] >= 0) {  // the chunk itself (RecomputeMeshes skips ids that are not in the map, :239-241)
      bool observed = true;
      int positive = 0;
#pragma unroll
      for (int k = 0; k < 8; k++) {  // (:681-731) all 8 corners must lie in existing chunks and have sdf <= 1
        const int px = v.x + mc_cx(k), py = v.y + mc_cy(k), pz = v.z + mc_cz(k);
        const int entry = nb[((px == 8) + 1) + 4 * ((py == 8) + 1) + 16 * ((pz == 8) + 1)];
        float s = 2.0f, w = 0.0f;
        if (entry >= 0 && observed) {
          const int vi = (px & 7) + 8 * (py & 7) + 64 * (pz & 7);
          s = mesh_sdf(md, entry, vi);
          w = mesh_weight(md, entry, vi);
        }
        if (s > 1.0f) observed = false;
        v.s[k] = s;
        v.w[k] = w;
        positive += s > 0.0f;
        index |= (0.0f > s) ? (1 << k) : 0;
      }
      v.meshed = observed && (positive % 8) > 0;
    }
    // triangles of this voxel whose three edges carry a valid normal (:864-941)
    const unsigned long long row = v.meshed ? kMcTriangles[index] : ~0ull;
    unsigned valid_edges = 0, tested = 0;
    int ntri = 0;
    unsigned tri_mask = 0;  // bit j: triangle j of the row is emitted
    if ((row & 15ull) != 15ull) {
#pragma unroll 1
      for (int j = 0; j < 5; j++) {
        const int s0 = (int)((row >> (12 * j)) & 15ull), s1 = (int)((row >> (12 * j + 4)) & 15ull), s2 = (int)((row >> (12 * j + 8)) & 15ull);
        if (s0 == 15) break;
        const int es[3] = {s0, s1, s2};
        bool all = true;
#pragma unroll
        for (int k = 0; k < 3; k++) {
          const unsigned bit = 1u << es[k];
          if (!(tested & bit)) {
            tested |= bit;
            float3 p, nn, cc;
            if (mc_edge(md, nb, v, es[k], a.res, a.l2r, &p, &nn, &cc)) valid_edges |= bit;
          }
          all = all && (valid_edges & bit);
        }
        if (all) {
          tri_mask |= 1u << j;
          ntri++;
          atomicMax(&owner[mc_slot(v.x, v.y, v.z, s2)], t);
          atomicMax(&owner[mc_slot(v.x, v.y, v.z, s1)], t);
          atomicMax(&owner[mc_slot(v.x, v.y, v.z, s0)], t);
        }
      }
    }
    __syncthreads();
    // compaction of the used slots in slot order (:965-975): 5 consecutive slots per thread
    int used = 0;
    for (int k = 0; k < 5; k++) {
      const int s = t * 5 + k;
      used += (s < kMeshSlots && owner[s] >= 0) ? 1 : 0;
    }
    int n_vert, n_tri;
    int base = mesh_block_scan(used, &n_vert, scratch);
    for (int k = 0; k < 5; k++) {
      const int s = t * 5 + k;
      if (s < kMeshSlots && owner[s] >= 0) rank[s] = base++;
    }
    const int tri_base = mesh_block_scan(ntri, &n_tri, scratch);  // (also orders the rank[] writes before the reads below)
    if (!kWrite) {
      if (t == 0) a.counts[c] = make_int2(n_vert, 3 * n_tri);
    } else if (ntri) {
      const long long v0 = a.off[c], i0 = a.off[a.n + 1 + c];
      // chunk origin (Chunk.cpp:52) + centroid of the voxel (ChunkManager.cpp:50-62)
      const float half = __fmul_rn(a.res, 0.5f);
      const float ox = __fadd_rn(__fmul_rn((float)(8 * id.x), a.res), __fadd_rn(__fmul_rn((float)v.x, a.res), half));
      const float oy = __fadd_rn(__fmul_rn((float)(8 * id.y), a.res), __fadd_rn(__fmul_rn((float)v.y, a.res), half));
      const float oz = __fadd_rn(__fmul_rn((float)(8 * id.z), a.res), __fadd_rn(__fmul_rn((float)v.z, a.res), half));
      int out = 0;
      unsigned written = 0;
      for (int j = 0; j < 5; j++) {
        if (!((tri_mask >> j) & 1u)) continue;
        const int es[3] = {(int)((row >> (12 * j + 8)) & 15ull), (int)((row >> (12 * j + 4)) & 15ull), (int)((row >> (12 * j)) & 15ull)};
        for (int k = 0; k < 3; k++) {  // indices in the order s2, s1, s0 (:934-936)
          const int slot = mc_slot(v.x, v.y, v.z, es[k]);
          const int r = rank[slot];
          a.idx[i0 + 3 * (tri_base + out) + k] = r;
          if (owner[slot] == t && !((written >> es[k]) & 1u)) {  // this voxel's value is the one that stays
            written |= 1u << es[k];
            float3 p, nn, cc;
            mc_edge(md, nb, v, es[k], a.res, a.l2r, &p, &nn, &cc);
            float* vo = a.vert + 3 * (v0 + r);
            float* no = a.norm + 3 * (v0 + r);
            float* co = a.col + 3 * (v0 + r);
            vo[0] = __fadd_rn(p.x, ox), vo[1] = __fadd_rn(p.y, oy), vo[2] = __fadd_rn(p.z, oz);
            no[0] = nn.x, no[1] = nn.y, no[2] = nn.z;
            co[0] = cc.x, co[1] = cc.y, co[2] = cc.z;
          }
        }
        out++;
      }
    }
    __syncthreads();
  }
}

// counts -> offsets: off[0..n] vertex offsets, off[n+1..2n+1] index offsets (one block; n is a few thousand)
__global__ void __launch_bounds__(1024) mesh_scan_kernel(const int2* __restrict__ counts, int n, long long* off) {
  __shared__ long long part[2][1024];
  const int t = threadIdx.x, per = (n + 1023) / 1024;
  long long sv = 0, si = 0;
  for (int k = t * per; k < min(n, (t + 1) * per); k++) sv += counts[k].x, si += counts[k].y;
  part[0][t] = sv, part[1][t] = si;
  __syncthreads();
  if (t == 0) {
    long long av = 0, ai = 0;
    for (int k = 0; k < 1024; k++) {
      const long long v = part[0][k], i = part[1][k];
      part[0][k] = av, part[1][k] = ai;
      av += v, ai += i;
    }
    off[n] = av, off[2 * n + 1] = ai;
  }
  __syncthreads();
  long long av = part[0][t], ai = part[1][t];
  for (int k = t * per; k < min(n, (t + 1) * per); k++) {
    off[k] = av, off[n + 1 + k] = ai;
    av += counts[k].x, ai += counts[k].y;
  }
}

}  // namespace tfb
