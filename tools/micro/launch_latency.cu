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
  return 0;
}
