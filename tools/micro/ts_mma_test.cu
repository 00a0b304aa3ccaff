// Micro-test: tcgen05.mma with the A operand in TENSOR MEMORY (TS form), the planned P.V of the window-attention kernel:
//   D[128 x 64] (fp32, TMEM) = A[128 x K] (bf16, TMEM, written with tcgen05.st: thread = row, two K elements per 32-bit
//   column) . B[K x 64] (bf16, shared memory, MN-major SWIZZLE_128B: row = k, 64 contiguous n), K = 208 (13 steps of 16).
// Checks the packing convention (low half = even k) and that a D region may overlap columns of the SAME lanes that A no
// longer uses. Prints the max abs error against a host reference for both packings.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "../../multimodal-sam-adapter_b200/csrc/common.cuh"
using namespace mmsam;

constexpr int KK = 208, N = 64;

__global__ void __launch_bounds__(128, 1) k(const __nv_bfloat16* A, const __nv_bfloat16* B, float* D, int swap_halves, int a_col, int d_col) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint32_t slot;
  __shared__ uint64_t bar;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, row = threadIdx.x;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(&slot, 512);
  // B -> smem, MN-major SW128: row k (128 B = 64 n), 16-byte chunk c at ((c ^ (k & 7)) << 4)
  for (int i = threadIdx.x; i < KK * 8; i += 128) {
    const int kr = i / 8, c = i % 8;
    *reinterpret_cast<uint4*>(smem + kr * 128 + ((c ^ (kr & 7)) << 4)) = *reinterpret_cast<const uint4*>(B + kr * N + c * 8);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
  // A row -> TMEM columns [a_col, a_col + KK/2): packed pairs
  for (int c0 = 0; c0 < KK / 2; c0 += 8) {
    uint32_t r[8];
    for (int j = 0; j < 8; ++j) {
      const int kk = 2 * (c0 + j);
      uint32_t lo = *reinterpret_cast<const uint16_t*>(A + row * KK + kk), hi = *reinterpret_cast<const uint16_t*>(A + row * KK + kk + 1);
      r[j] = swap_halves ? (hi | (lo << 16)) : (lo | (hi << 16));
    }
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(lane_addr + a_col + c0), "r"(r[0]),
                 "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
  }
  tmem_st_wait();
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) {
    tc_fence_after();
    constexpr uint32_t idesc = umma_idesc_bf16(128, N, 0, 1);       // A K-major (TMEM), B MN-major
    for (int s = 0; s < KK / 16; ++s)
      umma_f16_ts(tmem + d_col, tmem + a_col + s * 8, umma_desc_sw128(smem_u32(smem) + s * 16 * 128), idesc, s != 0);
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  for (int c = 0; c < N; c += 32) {
    uint32_t r[32];
    tmem_ld_32x32b_x32(lane_addr + d_col + c, r);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) D[row * N + c + j] = __uint_as_float(r[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

int main() {
  std::vector<__nv_bfloat16> hA(128 * KK), hB(KK * N);
  std::vector<float> fA(128 * KK), fB(KK * N), ref(128 * N), hD(128 * N);
  srand(1);
  for (size_t i = 0; i < hA.size(); ++i) { hA[i] = __float2bfloat16((rand() % 2001 - 1000) / 1000.f); fA[i] = __bfloat162float(hA[i]); }
  for (size_t i = 0; i < hB.size(); ++i) { hB[i] = __float2bfloat16((rand() % 2001 - 1000) / 1000.f); fB[i] = __bfloat162float(hB[i]); }
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < N; ++n) { double a = 0; for (int q = 0; q < KK; ++q) a += (double)fA[m * KK + q] * fB[q * N + n]; ref[m * N + n] = (float)a; }
  __nv_bfloat16 *dA, *dB; float* dD;
  cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dD, hD.size() * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  const int smem = KK * 128 + 1024;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  struct { int swap, a_col, d_col; const char* what; } cases[] = {
      {0, 0, 256, "low half = even k, D in separate columns"}, {1, 0, 256, "low half = odd k (expected wrong)"},
      {0, 0, 128, "low half = even k, D at column 128 of the same region (A uses columns 0..103)"},
      {0, 208, 208 + 128, "second slot: A at 208, D at 336"}};
  for (auto& c : cases) {
    cudaMemset(dD, 0, hD.size() * 4);
    k<<<1, 128, smem>>>(dA, dB, dD, c.swap, c.a_col, c.d_col);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
    double mx = 0;
    for (size_t i = 0; i < hD.size(); ++i) mx = fmax(mx, fabs((double)hD[i] - ref[i]));
    printf("%-80s max|err| %.3e  (%s)\n", c.what, mx, cudaGetErrorString(e));
  }
}
