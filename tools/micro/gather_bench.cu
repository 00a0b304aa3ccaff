// Micro-benchmark: how fast can an SM gather random 64-byte units (one head x 32 bf16 channels of the MSDeformAttn
// value map) as a function of the lane mapping. Settles the msda kernel design (see profiles/r01_msda_notes.md).
//   mode 0: 4 lanes x 16 B per unit (8 units per LDG instruction)      mode 1: 8 lanes x 8 B (4 units)
//   mode 2: 16 lanes x 4 B (2 units)                                    mode 3: 32 lanes x 2 B (1 unit)
//   mode 4: 16 lanes x 4 B, upper half-warp predicated off (1 unit per instruction, half the lanes idle)
//   mode 5: 32 lanes x 4 B reading 2 ADJACENT units (128 B contiguous, 64 B aligned: straddles a line half the time)
//   mode 6: as 5 but 128 B aligned (always one line)
//   mode 7: 8 lanes x 16 B per 128 B-aligned unit pair (4 pairs per LDG.128: the x-duplicated value layout)
//   mode 8: shared memory, 4 lanes x 16 B per random 64 B unit (LDS.128, random bank halves)
//   mode 9: shared memory, 8 lanes x 16 B = two 64 B units in opposite bank halves (conflict-free quarter-warp phases)
// window: number of 64 B units the random indices span per CTA (small = L1 hits, large = L2 hits).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t hash32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }
template <int MODE>
__global__ void __launch_bounds__(256) gather(const uint8_t* __restrict__ base, uint32_t units_total, uint32_t window, int iters, float* sink) {
  const int lane = threadIdx.x & 31, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t cta_base = (hash32(blockIdx.x * 7919u + 1u) % (units_total - window));
  float acc = 0.f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll 4
    for (int j = 0; j < 16; ++j) {
      const uint32_t seed = hash32((uint32_t)warp * 65537u + (uint32_t)(it * 16 + j));
      if (MODE == 0) { uint32_t u = cta_base + hash32(seed + (lane >> 2)) % window; uint4 v = *reinterpret_cast<const uint4*>(base + (size_t)u * 64 + (lane & 3) * 16); acc += __uint_as_float(v.x ^ v.y ^ v.z ^ v.w); }
      if (MODE == 1) { uint32_t u = cta_base + hash32(seed + (lane >> 3)) % window; uint2 v = *reinterpret_cast<const uint2*>(base + (size_t)u * 64 + (lane & 7) * 8); acc += __uint_as_float(v.x ^ v.y); }
      if (MODE == 2) { uint32_t u = cta_base + hash32(seed + (lane >> 4)) % window; uint32_t v = *reinterpret_cast<const uint32_t*>(base + (size_t)u * 64 + (lane & 15) * 4); acc += __uint_as_float(v); }
      if (MODE == 3) { uint32_t u = cta_base + seed % window; uint16_t v = *reinterpret_cast<const uint16_t*>(base + (size_t)u * 64 + lane * 2); acc += (float)v; }
      if (MODE == 4) { uint32_t u = cta_base + seed % window; if (lane < 16) { uint32_t v = *reinterpret_cast<const uint32_t*>(base + (size_t)u * 64 + lane * 4); acc += __uint_as_float(v); } }
      if (MODE == 5) { uint32_t u = cta_base + seed % (window - 1); uint32_t v = *reinterpret_cast<const uint32_t*>(base + (size_t)u * 64 + lane * 4); acc += __uint_as_float(v); }
      if (MODE == 7) { uint32_t u = (cta_base + hash32(seed + (lane >> 3)) % (window - 1)) & ~1u; uint4 v = *reinterpret_cast<const uint4*>(base + (size_t)u * 64 + (lane & 7) * 16); acc += __uint_as_float(v.x ^ v.y ^ v.z ^ v.w); }
      if (MODE == 6) { uint32_t u = (cta_base + seed % (window - 1)) & ~1u; uint32_t v = *reinterpret_cast<const uint32_t*>(base + (size_t)u * 64 + lane * 4); acc += __uint_as_float(v); }
    }
  }
  if (acc == 123.456f) sink[0] = acc;
}
template <int MODE>
__global__ void __launch_bounds__(256) gather_smem(const uint8_t* __restrict__ base, int iters, float* sink) {
  extern __shared__ __align__(128) uint8_t sm[];
  constexpr uint32_t UNITS = 96 * 1024 / 64;
  for (uint32_t i = threadIdx.x; i < UNITS * 4; i += blockDim.x) reinterpret_cast<uint4*>(sm)[i] = reinterpret_cast<const uint4*>(base)[i + blockIdx.x * 64];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  float acc = 0.f;
  // cheap per-lane-group LCG (1 IMAD + 1 LOP per load) so the loop measures the LDS pipe, not the index arithmetic
  uint32_t st = hash32((uint32_t)warp * 65537u + (uint32_t)(MODE == 8 ? (lane >> 2) : (lane >> 3)));
  const uint32_t sub = (lane & 3) * 16;
  const uint32_t odd = (lane >> 2) & 1;
  for (int it = 0; it < iters; ++it) {
#pragma unroll 8
    for (int j = 0; j < 16; ++j) {
      st = st * 1664525u + 1013904223u;
      uint32_t u = (st >> 12) & (1024u - 1u);        // 1024 units = 64 KB of the 96 KB window
      if (MODE == 9) u = (u & ~1u) | odd;            // the two 4-lane groups of a quarter warp: opposite bank halves
      uint4 v = *reinterpret_cast<const uint4*>(sm + u * 64 + sub);
      acc += __uint_as_float(v.x ^ v.y ^ v.z ^ v.w);
    }
  }
  if (acc == 123.456f) sink[0] = acc;
}
template <int MODE> void run_smem(const uint8_t* d, float* sink, const char* name) {
  const int iters = 64, blocks = 148 * 8;
  cudaFuncSetAttribute(gather_smem<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
  gather_smem<MODE><<<blocks, 256, 96 * 1024>>>(d, 2, sink);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a);
  gather_smem<MODE><<<blocks, 256, 96 * 1024>>>(d, iters, sink);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  const double u = (double)blocks * 8 * iters * 16 * 8;
  printf("  %-44s smem 96 KB window     : %7.1f us  %6.2f G units/s  %6.2f TB/s gathered  %.2f cycles/unit/SM @1.9GHz (incl. 96 KB fill per CTA)\n", name, ms * 1e3, u / ms / 1e6, u * 64 / ms / 1e9, ms * 1e-3 * 1.9e9 * 148 / u);
}
template <int MODE> void run(const uint8_t* d, uint32_t units, uint32_t window, float* sink, const char* name) {
  const int iters = 64, blocks = 148 * 8;
  const double units_per_instr[8] = {8, 4, 2, 1, 1, 2, 2, 8};
  gather<MODE><<<blocks, 256>>>(d, units, window, 2, sink);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a);
  gather<MODE><<<blocks, 256>>>(d, units, window, iters, sink);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  const double instr = (double)blocks * 8 * iters * 16;
  const double u = instr * units_per_instr[MODE];
  printf("  %-44s window %8u units: %7.1f us  %6.2f G units/s  %6.2f TB/s gathered  %.2f cycles/unit/SM @1.9GHz\n", name, window, ms * 1e3, u / ms / 1e6, u * 64 / ms / 1e9, ms * 1e-3 * 1.9e9 * 148 / u);
}
int main() {
  const uint32_t units = 1u << 20;   // 64 MB
  uint8_t* d; cudaMalloc(&d, (size_t)units * 64); cudaMemset(d, 1, (size_t)units * 64);
  float* sink; cudaMalloc(&sink, 4);
  for (uint32_t window : {512u, 8192u, 500000u}) {
    run<0>(d, units, window, sink, "mode0 4 lanes x 16B (8 units/instr)");
    run<1>(d, units, window, sink, "mode1 8 lanes x 8B (4 units/instr)");
    run<2>(d, units, window, sink, "mode2 16 lanes x 4B (2 units/instr)");
    run<3>(d, units, window, sink, "mode3 32 lanes x 2B (1 unit/instr)");
    run<4>(d, units, window, sink, "mode4 16 lanes x 4B, half warp off (1 unit)");
    run<5>(d, units, window, sink, "mode5 32 lanes x 4B adjacent pair, 64B aligned");
    run<6>(d, units, window, sink, "mode6 32 lanes x 4B adjacent pair, 128B aligned");
    run<7>(d, units, window, sink, "mode7 8 lanes x 16B aligned pair (4 pairs/instr)");
  }
  run_smem<8>(d, sink, "mode8 smem 4 lanes x 16B random units");
  run_smem<9>(d, sink, "mode9 smem 8 lanes = 2 units, opposite halves");
  cudaError_t e = cudaDeviceSynchronize();
  printf("%s\n", cudaGetErrorString(e));
  return 0;
}
