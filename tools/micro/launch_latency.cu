// Host-side launch / completion latency probe (not part of the product): how long do a kernel
// launch, a stream synchronize and a mapped-memory completion flag take on this box?
#include <chrono>
#include <cstdio>
#include <cstring>
#include <cuda_runtime.h>
struct Big { char b[1400]; };
__global__ void k_small(int* flag, int v) { if (threadIdx.x == 0 && blockIdx.x == 0 && flag) { __threadfence_system(); *(volatile int*)flag = v; } }
__global__ void k_big(const __grid_constant__ Big p, int* flag, int v) { if (threadIdx.x == 0 && blockIdx.x == 0 && flag) { __threadfence_system(); *(volatile int*)flag = v + p.b[0]; } }
static double now() { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main() {
  cudaStream_t s; cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
  int* hflag; cudaHostAlloc(&hflag, 64, cudaHostAllocMapped); int* dflag; cudaHostGetDevicePointer(&dflag, hflag, 0);
  Big big; memset(&big, 0, sizeof(big));
  const int N = 2000;
  for (int mode = 0; mode < 6; mode++) {
    double t_l1 = 0, t_l2 = 0, t_l3 = 0, t_sync = 0, t_tot = 0;
    for (int i = 0; i < N + 100; i++) {
      *hflag = 0;
      double a = now();
      if (mode == 0 || mode == 2 || mode == 4) k_small<<<296, 256, 0, s>>>(nullptr, 0); else k_big<<<296, 256, 0, s>>>(big, nullptr, 0);
      double b = now();
      if (mode == 0 || mode == 2 || mode == 4) k_small<<<592, 256, 0, s>>>(nullptr, 0); else k_big<<<592, 256, 0, s>>>(big, nullptr, 0);
      double c = now();
      if (mode == 0 || mode == 2 || mode == 4) k_small<<<592, 256, 0, s>>>(dflag, i + 1); else k_big<<<592, 256, 0, s>>>(big, dflag, i + 1);
      double d = now();
      if (mode < 2) cudaStreamSynchronize(s);
      else if (mode < 4) { while (*(volatile int*)hflag != i + 1) {} }
      else { while (cudaStreamQuery(s) == cudaErrorNotReady) {} }
      double e = now();
      if (mode >= 2 && mode < 4) cudaStreamSynchronize(s);
      if (i >= 100) { t_l1 += b - a; t_l2 += c - b; t_l3 += d - c; t_sync += e - d; t_tot += e - a; }
    }
    const char* names[] = {"small params + streamSync", "1.4KB params + streamSync", "small + mapped flag spin", "1.4KB + mapped flag spin", "small + streamQuery spin", "1.4KB + streamQuery spin"};
    printf("%-28s launch1 %.2f launch2 %.2f launch3 %.2f wait %.2f total %.2f us\n", names[mode], t_l1 / N, t_l2 / N, t_l3 / N, t_sync / N, t_tot / N);
  }
  // graph of three kernels
  cudaGraph_t g; cudaGraphExec_t ge;
  cudaStreamBeginCapture(s, cudaStreamCaptureModeGlobal);
  k_big<<<296, 256, 0, s>>>(big, nullptr, 0); k_big<<<592, 256, 0, s>>>(big, nullptr, 0); k_big<<<592, 256, 0, s>>>(big, dflag, 7);
  cudaStreamEndCapture(s, &g); cudaGraphInstantiate(&ge, g, 0);
  double t_l = 0, t_w = 0;
  for (int i = 0; i < N + 100; i++) {
    *hflag = 0;
    double a = now(); cudaGraphLaunch(ge, s); double b = now();
    while (*(volatile int*)hflag != 7) {}
    double c = now(); cudaStreamSynchronize(s);
    if (i >= 100) { t_l += b - a; t_w += c - b; }
  }
  printf("graph(3 kernels, 1.4KB): launch %.2f wait(flag) %.2f total %.2f us\n", t_l / N, t_w / N, (t_l + t_w) / N);
  // graph with per-launch parameter updates
  {
    cudaGraphNode_t nodes[8]; size_t nn = 8; cudaGraphGetNodes(g, nodes, &nn);
    cudaKernelNodeParams kp[3];
    for (size_t k = 0; k < nn && k < 3; k++) cudaGraphKernelNodeGetParams(nodes[k], &kp[k]);
    double t_set = 0; t_l = 0; t_w = 0;
    for (int i = 0; i < N + 100; i++) {
      *hflag = 0;
      double a = now();
      for (size_t k = 0; k < nn && k < 3; k++) cudaGraphExecKernelNodeSetParams(ge, nodes[k], &kp[k]);
      double b = now(); cudaGraphLaunch(ge, s); double c = now();
      while (*(volatile int*)hflag != 7) {}
      double d = now(); cudaStreamSynchronize(s);
      if (i >= 100) { t_set += b - a; t_l += c - b; t_w += d - c; }
    }
    printf("graph + 3x SetParams: set %.2f launch %.2f wait %.2f total %.2f us\n", t_set / N, t_l / N, t_w / N, (t_set + t_l + t_w) / N);
  }
  // graph with a leading 4 KB H2D memcpy node (parameters in a pinned buffer)
  {
    char* hp; cudaHostAlloc(&hp, 4096, cudaHostAllocDefault); char* dp; cudaMalloc(&dp, 4096);
    cudaGraph_t g2; cudaGraphExec_t ge2;
    cudaStreamBeginCapture(s, cudaStreamCaptureModeGlobal);
    cudaMemcpyAsync(dp, hp, 2560, cudaMemcpyHostToDevice, s);
    k_small<<<296, 256, 0, s>>>(nullptr, 0); k_small<<<592, 256, 0, s>>>(nullptr, 0); k_small<<<592, 256, 0, s>>>(dflag, 7);
    cudaStreamEndCapture(s, &g2); cudaGraphInstantiate(&ge2, g2, 0);
    t_l = 0; t_w = 0;
    for (int i = 0; i < N + 100; i++) {
      *hflag = 0;
      double a = now(); cudaGraphLaunch(ge2, s); double b = now();
      while (*(volatile int*)hflag != 7) {}
      double c = now(); cudaStreamSynchronize(s);
      if (i >= 100) { t_l += b - a; t_w += c - b; }
    }
    printf("graph(memcpy 2.5KB + 3 kernels): launch %.2f wait %.2f total %.2f us\n", t_l / N, t_w / N, (t_l + t_w) / N);
  }
  // single kernel + flag
  {
    double t1 = 0, t2 = 0;
    for (int i = 0; i < N + 100; i++) {
      *hflag = 0;
      double a = now(); k_big<<<592, 256, 0, s>>>(big, dflag, i + 1); double b = now();
      while (*(volatile int*)hflag != i + 1) {}
      double c = now(); cudaStreamSynchronize(s);
      if (i >= 100) { t1 += b - a; t2 += c - b; }
    }
    printf("one kernel + flag: launch %.2f wait %.2f total %.2f us\n", t1 / N, t2 / N, (t1 + t2) / N);
  }
  // three PDL launches + flag
  {
    double t1 = 0, t2 = 0;
    cudaLaunchConfig_t cfg{}; cfg.blockDim = dim3(256); cfg.stream = s;
    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
    for (int i = 0; i < N + 100; i++) {
      *hflag = 0;
      double a = now();
      cfg.gridDim = dim3(296); cfg.attrs = nullptr; cfg.numAttrs = 0;
      cudaLaunchKernelEx(&cfg, k_big, big, (int*)nullptr, 0);
      cfg.gridDim = dim3(592); cfg.attrs = at; cfg.numAttrs = 1;
      cudaLaunchKernelEx(&cfg, k_big, big, (int*)nullptr, 0);
      cudaLaunchKernelEx(&cfg, k_big, big, dflag, i + 1);
      double b = now();
      while (*(volatile int*)hflag != i + 1) {}
      double c = now(); cudaStreamSynchronize(s);
      if (i >= 100) { t1 += b - a; t2 += c - b; }
    }
    printf("3 launches (2 PDL) + flag: launch %.2f wait %.2f total %.2f us\n", t1 / N, t2 / N, (t1 + t2) / N);
  }
  // graph of three kernels captured WITH programmatic dependent launch edges (as the product's frame graph)
  {
    cudaGraph_t g3; cudaGraphExec_t ge3;
    cudaLaunchConfig_t cfg{}; cfg.blockDim = dim3(256); cfg.stream = s;
    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
    cudaStreamBeginCapture(s, cudaStreamCaptureModeGlobal);
    cfg.gridDim = dim3(1200); cfg.attrs = nullptr; cfg.numAttrs = 0;
    cudaLaunchKernelEx(&cfg, k_big, big, (int*)nullptr, 0);
    cfg.gridDim = dim3(592); cfg.attrs = at; cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, k_big, big, (int*)nullptr, 0);
    cudaLaunchKernelEx(&cfg, k_big, big, dflag, 7);
    cudaStreamEndCapture(s, &g3); cudaGraphInstantiate(&ge3, g3, 0);
    cudaGraphNode_t nodes[8]; size_t nn = 8; cudaGraphGetNodes(g3, nodes, &nn);
    cudaKernelNodeParams kp[3];
    for (size_t k = 0; k < nn && k < 3; k++) cudaGraphKernelNodeGetParams(nodes[k], &kp[k]);
    for (int with_set = 0; with_set < 2; with_set++) {
      double t_set = 0, t_l = 0, t_w = 0;
      for (int i = 0; i < N + 100; i++) {
        *hflag = 0;
        double a = now();
        if (with_set) for (size_t k = 0; k < nn && k < 3; k++) cudaGraphExecKernelNodeSetParams(ge3, nodes[k], &kp[k]);
        double b = now(); cudaGraphLaunch(ge3, s); double c = now();
        while (*(volatile int*)hflag != 7) {}
        double d = now(); cudaStreamSynchronize(s);
        if (i >= 100) { t_set += b - a; t_l += c - b; t_w += d - c; }
      }
      printf("graph with PDL edges%s: set %.2f launch %.2f wait %.2f total %.2f us\n", with_set ? " + 3x SetParams" : "", t_set / N, t_l / N, t_w / N, (t_set + t_l + t_w) / N);
    }
  }
  // the same graph launched the way bench.py's timed step does: after an L2 flush and a device synchronize
  {
    cudaGraph_t g4; cudaGraphExec_t ge4;
    cudaStreamBeginCapture(s, cudaStreamCaptureModeGlobal);
    k_big<<<1200, 256, 0, s>>>(big, nullptr, 0); k_big<<<592, 256, 0, s>>>(big, nullptr, 0); k_big<<<592, 256, 0, s>>>(big, dflag, 7);
    cudaStreamEndCapture(s, &g4); cudaGraphInstantiate(&ge4, g4, 0);
    char* flush; cudaMalloc(&flush, 512u << 20);
    for (int mode = 0; mode < 3; mode++) {
      double t_l = 0, t_w = 0;
      const int M = 300;
      for (int i = 0; i < M + 20; i++) {
        *hflag = 0;
        if (mode >= 1) { cudaMemsetAsync(flush, i & 255, 512u << 20, 0); cudaDeviceSynchronize(); }
        if (mode == 2) { double t0 = now(); while (now() - t0 < 200.0) {} }  // + 200 us of host-side idle
        double a = now(); cudaGraphLaunch(ge4, s); double b = now();
        while (*(volatile int*)hflag != 7) {}
        double c = now(); cudaStreamSynchronize(s);
        if (i >= 20) { t_l += b - a; t_w += c - b; }
      }
      const char* nm[] = {"back to back", "after flush + deviceSync", "after flush + deviceSync + 200us idle"};
      printf("graph launch %-38s: launch %.2f wait %.2f total %.2f us\n", nm[mode], t_l / M, t_w / M, (t_l + t_w) / M);
    }
  }
  // SetParams with CHANGING values (the product patches a new pose etc. every frame)
  {
    cudaGraph_t g5; cudaGraphExec_t ge5;
    cudaStreamBeginCapture(s, cudaStreamCaptureModeGlobal);
    k_big<<<1200, 256, 0, s>>>(big, nullptr, 0); k_big<<<592, 256, 0, s>>>(big, nullptr, 0); k_big<<<592, 256, 0, s>>>(big, dflag, 7);
    cudaStreamEndCapture(s, &g5); cudaGraphInstantiate(&ge5, g5, 0);
    cudaGraphNode_t nodes[8]; size_t nn = 8; cudaGraphGetNodes(g5, nodes, &nn);
    cudaKernelNodeParams kp[3];
    for (size_t k = 0; k < 3; k++) cudaGraphKernelNodeGetParams(nodes[k], &kp[k]);
    int* fl[3] = {nullptr, nullptr, dflag}; int vv[3] = {0, 0, 7};
    // find which node is which by its grid / flag is not possible from here: patch only the struct argument (same for all)
    double t_set = 0, t_l = 0, t_w = 0;
    for (int i = 0; i < N + 100; i++) {
      *hflag = 0;
      big.b[1] = (char)i; big.b[700] = (char)(i >> 3);
      double a = now();
      for (size_t k = 0; k < 3; k++) {
        void** orig = kp[k].kernelParams;
        void* args[3] = {&big, orig[1], orig[2]};
        cudaKernelNodeParams p2 = kp[k]; p2.kernelParams = args;
        cudaGraphExecKernelNodeSetParams(ge5, nodes[k], &p2);
      }
      double b = now(); cudaGraphLaunch(ge5, s); double c = now();
      while (*(volatile int*)hflag != 7) {}
      double d = now(); cudaStreamSynchronize(s);
      if (i >= 100) { t_set += b - a; t_l += c - b; t_w += d - c; }
    }
    (void)fl; (void)vv;
    printf("graph + 3x SetParams with changing values: set %.2f launch %.2f wait %.2f total %.2f us\n", t_set / N, t_l / N, t_w / N, (t_set + t_l + t_w) / N);
  }
  return 0;
}
