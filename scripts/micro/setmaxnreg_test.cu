// Does warpgroup register reallocation (setmaxnreg) give the 16 epilogue warps of a 20-warp CTA more than the 96
// registers the launch bound allows?  640 threads: warps 0-15 "epilogue" (inc to NINC), warps 16-19 dec to 24.
#include <cstdio>
#include <cuda_runtime.h>
#ifndef NINC
#define NINC 112
#endif
template <int N> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }

__global__ void __launch_bounds__(640, 1) k(const float* __restrict__ in, float* out, int iters) {
  const int warp = threadIdx.x >> 5;
  if (warp >= 16) {
    reg_dec<24>();
    return;
  }
  reg_inc<NINC>();
  // ~100 live values per thread
  float v[100];
#pragma unroll
  for (int i = 0; i < 100; ++i) v[i] = in[threadIdx.x + 640 * i];
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 100; ++i) v[i] = fmaf(v[i], v[(i + 37) % 100], 0.5f);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 100; ++i) s += v[i];
  out[blockIdx.x * 640 + threadIdx.x] = s;
}

int main() {
  float *in, *out;
  cudaMalloc(&in, 640 * 100 * 4);
  cudaMemset(in, 0, 640 * 100 * 4);
  cudaMalloc(&out, 148 * 640 * 4);
  k<<<148, 640>>>(in, out, 10);
  cudaError_t e = cudaDeviceSynchronize();
  printf("NINC=%d result: %s\n", NINC, cudaGetErrorString(e));
  return e != cudaSuccess;
}
