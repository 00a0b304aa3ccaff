// Micro-benchmark: MUFU.EX2 issue rate per SM sub-partition (how many cycles one warp-wide ex2 occupies the pipe),
// alone and interleaved with FFMA2 work. Decides whether the attention softmax is MUFU-bound.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
template <int FMA_PER_EX>
__global__ void __launch_bounds__(1024, 1) k(int iters, float* out, long long* cyc, float seed) {
  float v[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = seed * (i + threadIdx.x);
  float w[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) w[i] = seed + i;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      v[i] = ex2(v[i]);
#pragma unroll
      for (int f = 0; f < FMA_PER_EX; ++f) w[i] = fmaf(w[i], 1.0001f, 0.5f);
    }
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += v[i] + w[i];
  if (s == 1.2345f) out[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
  float* out; long long* cyc; cudaMalloc(&out, 4); cudaMalloc(&cyc, 8);
  for (int warps : {4, 8, 16}) {
    const int iters = 4000;
    long long h;
    k<0><<<148, warps * 32>>>(iters, out, cyc, 0.001f); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("  %2d warps/SM, ex2 only     : %.2f cycles per warp-ex2 per SMSP\n", warps, (double)h / (iters * 16.0 * (warps / 4)));
    k<4><<<148, warps * 32>>>(iters, out, cyc, 0.001f); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("  %2d warps/SM, ex2 + 4 FFMA : %.2f cycles per warp-ex2 per SMSP\n", warps, (double)h / (iters * 16.0 * (warps / 4)));
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
