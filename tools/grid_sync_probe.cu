// What a cooperative launch and its grid barriers cost on this GPU (floor of the one-launch steps), and what a split-phase
// barrier built from one atomic counter costs.  Build and run on the GPU box:
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/grid_sync_probe tools/grid_sync_probe.cu && /tmp/grid_sync_probe
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

__global__ void k_plain(int* sink) { if (threadIdx.x == 9999) *sink = 1; }
__global__ void k_coop(int n_sync, int* sink) {
  cg::grid_group g = cg::this_grid();
  for (int i = 0; i < n_sync; ++i) {
    __threadfence();
    g.sync();
  }
  if (threadIdx.x == 9999) *sink = 1;
}
// one atomic per CTA on a monotonically increasing counter; everyone polls it
__global__ void k_atomic(int n_sync, unsigned* counter, unsigned base, int* sink) {
  for (int i = 0; i < n_sync; ++i) {
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      atomicAdd(counter, 1u);
      const unsigned target = base + (i + 1) * gridDim.x;
      while (*reinterpret_cast<volatile unsigned*>(counter) < target) __nanosleep(40);
      __threadfence();
    }
    __syncthreads();
  }
  if (threadIdx.x == 9999) *sink = 1;
}

template <typename F>
static float time_us(F launch, int iters) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int i = 0; i < 20; ++i) launch(i);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  for (int i = 0; i < iters; ++i) launch(20 + i);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms * 1e3f / iters;
}

int main() {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int* sink;
  unsigned* counter;
  cudaMalloc(&sink, 4);
  cudaMalloc(&counter, 4);
  cudaMemset(counter, 0, 4);
  for (int threads : {256, 512, 768}) {
    const int iters = 500;
    float t0 = time_us([&](int) { k_plain<<<sms, threads>>>(sink); }, iters);
    printf("threads %4d: plain launch %.2f us", threads, t0);
    for (int ns : {0, 1, 2}) {
      int n = ns;
      void* args[] = {&n, &sink};
      float t = time_us([&](int) { cudaLaunchCooperativeKernel((void*)k_coop, dim3(sms), dim3(threads), args, 0, 0); }, iters);
      printf(" | coop %d sync %.2f us", ns, t);
    }
    unsigned base = 0;
    cudaMemset(counter, 0, 4);
    for (int ns : {1, 2}) {
      float t = time_us([&](int) { k_atomic<<<sms, threads>>>(ns, counter, base, sink); base += ns * sms; }, iters);
      printf(" | atomic %d sync %.2f us", ns, t);
    }
    printf("\n");
  }
  printf("last error: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
