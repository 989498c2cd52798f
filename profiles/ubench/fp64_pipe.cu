// Micro-benchmarks of the sm_100a FP64 pipe as seen by one SM: DFMA issue rate per warp count / ILP,
// dependent-issue latency, shared-memory load latency.  build: nvcc -arch=sm_100a -O3 -o fp64_pipe fp64_pipe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void dfma_rate(double* out, long long* cyc, int iters) {
  double a[ILP];
#pragma unroll
  for (int u = 0; u < ILP; ++u) a[u] = 1.0 + threadIdx.x * 1e-9 + u;
  const double b = 1.0000001, c = 1e-9;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int u = 0; u < ILP; ++u) a[u] = fma(a[u], b, c);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int u = 0; u < ILP; ++u) s += a[u];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void lds_latency(double* out, long long* cyc, int iters) {
  __shared__ int idx[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) idx[i] = (i * 33 + 7) & 1023;
  __syncthreads();
  int p = threadIdx.x;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) p = idx[p];
  long long t1 = clock64();
  out[threadIdx.x] = p;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

__global__ void bar_latency(double* out, long long* cyc, int iters) {
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  out[threadIdx.x] = 0;
}

int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 1024 * 8);
  long long h[8];
  const int iters = 2000;
  int warps[] = {1, 2, 4, 8, 16, 32};
  printf("DFMA: cycles per warp-instruction per SMSP (1 CTA on one SM)\n");
  for (int w : warps) {
    printf("warps %2d:", w);
#define RUN(I) { dfma_rate<I><<<1, 32 * w>>>(out, cyc, iters); cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost); \
      double ninst = (double)iters * 8 * I * ((w + 3) / 4);  /* warp-instr per SMSP */                            \
      printf("  ILP%d %6.2f (lat-bound chain %5.1f)", I, h[0] / ninst, (double)h[0] / (iters * 8)); }
    RUN(1) RUN(2) RUN(4) RUN(8)
    printf("\n");
  }
  lds_latency<<<1, 32>>>(out, cyc, 4000); cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("LDS dependent latency: %.1f cycles\n", h[0] / 4000.0);
  for (int w : {4, 16}) { bar_latency<<<1, 32 * w>>>(out, cyc, 4000); cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("__syncthreads (%d warps): %.1f cycles\n", w, h[0] / 4000.0); }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
