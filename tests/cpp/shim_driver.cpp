// Drives the C++ shim (texturefusion_b200/host/chisel_b200.h) the way
// MobileFusion::ReIntegrateKeyframe does (GCFusion/MobileFusion.cpp:114-221) and prints a
// summary that tests/test_shim_gpu.py compares with the CPU oracle.
// usage: shim_driver <input.bin> <voxel_res>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "chisel_b200.h"

using namespace chisel;

// position-weighted checksum over 32-bit words: h += word * (running index), mod 2^64
struct Sum {
  uint64_t h = 0, idx = 0;
  void add(const void* p, size_t nbytes) {
    const uint32_t* w = (const uint32_t*)p;
    for (size_t i = 0; i < nbytes / 4; i++) h += (uint64_t)w[i] * (++idx);
  }
};

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 2;
  const float res = (float)atof(argv[2]);
  int32_t hdr[3];  // W, H, n_frames
  if (fread(hdr, 4, 3, f) != 3) return 2;
  const int W = hdr[0], H = hdr[1], nfr = hdr[2];
  const size_t npix = (size_t)W * H;
  float camf[6];  // fx fy cx cy near far
  if (fread(camf, 4, 6, f) != 6) return 2;
  PinholeCamera cam;
  cam.SetIntrinsics(camf[0], camf[1], camf[2], camf[3]);
  cam.SetWidth(W); cam.SetHeight(H); cam.SetNearPlane(camf[4]); cam.SetFarPlane(camf[5]);

  ProjectionIntegrator integrator;
  integrator.SetTruncator(TruncatorPtr(new QuadraticTruncator(0.0019f, 0.00152f, 0.001504f, 6.0f)));
  integrator.SetWeighter(WeighterPtr(new ConstantWeighter(1.0f)));
  try {
    Chisel chiselMap(ChunkID(8, 8, 8), res, true, W, H, &integrator);
    ChunkIDList ids, valid;
    std::vector<bool> needsUpdate, isNew;
    for (int k = 0; k < nfr; k++) {
      Transform pose;
      int32_t kf_index;
      std::vector<float> depth(npix), quality(npix);
      std::vector<unsigned char> rgba(npix * 4);
#ifdef TF_WITH_EIGEN
      if (fread(pose.matrix().data(), 4, 16, f) != 16) return 2;
#else
      if (fread(pose.m, 4, 16, f) != 16) return 2;
#endif
      if (fread(&kf_index, 4, 1, f) != 1) return 2;
      if (fread(depth.data(), 4, npix, f) != npix) return 2;
      if (kf_index >= 0) {
        if (fread(rgba.data(), 1, npix * 4, f) != npix * 4) return 2;
        if (fread(quality.data(), 4, npix, f) != npix) return 2;
        chiselMap.PrepareIntersectChunks(integrator, depth.data(), pose, cam, ids, needsUpdate, isNew);
        chiselMap.IntegrateDepthScanColor(integrator, depth.data(), rgba.data(), pose, cam, ids, needsUpdate, 1, kf_index,
                                          quality.data());
      } else {
        chiselMap.IntegrateDepthScanColor(integrator, depth.data(), nullptr, pose, cam, ids, needsUpdate, 1);
      }
    }
    chiselMap.FinalizeIntegrateChunks(ids, needsUpdate, isNew, valid);
    chiselMap.chunkManager.SyncToHost(valid);
    Sum h, hobs;
    size_t nobs = 0;
    for (const ChunkID& id : valid) {
      ChunkPtr c = chiselMap.chunkManager.GetChunk(id);
      h.add(c->voxels.sdf.data(), 2048);
      h.add(c->voxels.weight.data(), 2048);
      h.add(c->colors.colorData.data(), 4096);
      for (auto& o : c->observations) { hobs.add(&o.first, 4); hobs.add(&o.second, 4); nobs++; }
    }
    size_t nmesh = 0;
    for (auto& kv : chiselMap.meshesToUpdate) nmesh += kv.second ? 1 : 0;
    printf("ids=%zu valid=%zu chunks=%lld voxel_hash=%016llx obs=%zu obs_hash=%016llx meshes=%zu\n", ids.size(), valid.size(),
           (long long)chiselMap.chunkManager.GetChunkCount(), (unsigned long long)h.h, nobs, (unsigned long long)hobs.h, nmesh);
    // error behaviour: GetChunk on an unknown id throws std::out_of_range like unordered_map::at
    try { chiselMap.chunkManager.GetChunk(ChunkID(100000, 0, 0)); printf("no throw\n"); return 1; }
    catch (const std::out_of_range&) { printf("out_of_range ok\n"); }
    // the fused convenience overload (Structure/Chisel.h:453-468)
    chiselMap.Reset();
    rewind(f);
    return 0;
  } catch (const std::exception& e) {
    fprintf(stderr, "shim_driver: %s\n", e.what());
    return 1;
  }
}
