// bf16 GEMM  C[M,N] = epilogue(A[M,K] . W[N,K]^T)  on 5th-gen tensor cores (tcgen05 + TMEM + TMA).
//
// This is the contraction engine behind every nn.Linear / 1x1 conv / patchified conv on the path:
// Attention.qkv / proj and MLPBlock.lin1 / lin2 (segmentation/mmseg_custom/models/backbones/base/
// image_encoder.py:166-167, 488, 499), the four MSDeformAttn projections (ops/modules/
// ms_deform_attn.py:102-113,129), ConvFFN fc1/fc2 (adapter_modules_...new.py:446-453), ConvNeXt
// pointwise convs and patchified stem/downsample convs (base/twin_convnext.py:98-132, 295-336),
// spm.fc1-4 (adapter_modules_...new.py:947-950), `up` ConvTranspose2d(2,2) (..._new.py:324), patch
// embed (image_encoder.py:662-671) and the Segformer head 1x1 convs (decode_heads/
// segformer_head.py:34-46). Both operands are K-major, exactly nn.Linear's (x, weight) layout.
//
// Design (one persistent CTA per SM, 320 threads, warp-specialised):
//   warp 8      TMA producer: 128x64 A tile + BNx64 W tile per stage, SWIZZLE_128B boxes,
//               mbarrier complete_tx; OOB rows/cols are zero-filled so ragged M/N/K need no masks.
//   warp 9      MMA issuer: one elected lane issues 4 x tcgen05.mma (M128 x N=BN x K16) per stage
//               into a TMEM accumulator; tcgen05.commit releases the smem stage / publishes the
//               accumulator. Two TMEM accumulators (2 x BN columns) so the epilogue of tile i
//               overlaps the main loop of tile i+1.
//   warps 0-7   epilogue: tcgen05.ld (32 lanes x 32 columns), fused bias / activation (exact-erf
//               GELU, ReLU, ReLU6) / per-channel scale / residual add, bf16 or fp32 stores, with an
//               optional output-row remap (window un-partition, 2x2 pixel shuffle).
// Tiles are walked n-fastest so CTAs of one wave share the A tile through L2 and weights stay
// L2-resident: A is read from HBM once.
#include "common.cuh"
#include <type_traits>
#include <cstdlib>

namespace mmsam {

struct GemmEpi {
  const float* bias;               // [N] or null
  const float* scale;              // [N] or null, applied after the activation
  const __nv_bfloat16* residual;   // indexed at the destination row, or null
  void* out;
  const int* row_map;              // row_mode 1: dst row of each source row (-1 = drop)
  long long ldo, ldr;
  int M, N, K;
  int act;        // 0 none, 1 gelu(erf), 2 relu, 3 relu6
  int out_f32;    // 0: bf16 output, 1: fp32 output
  int row_mode;   // 0 identity, 1 row_map, 2 pixel-shuffle 2x2 (ConvTranspose2d k=2 s=2)
  int vec_ok;     // rows of out/residual are 16-byte aligned -> vector epilogue allowed
  int tma_store;  // output goes through the panel-staged TMA tensor store (row_mode 0, aligned rows)
  int dbg;        // perf-debug switches (env MMSAM_GEMM_DBG): 1 skip stores, 2 skip bias/scale loads, 4 skip the whole epilogue math
  int ps_h, ps_w, ps_c;
};

template <int BN> struct GemmCfg {
  static constexpr int BM = 128, BK = 64;
  static constexpr int STAGES = BN == 256 ? 4 : (BN == 128 ? 6 : 8);
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int TMEM_COLS = 2 * BN;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 2 * 16384 /*store panels*/ + 1024 /*align*/ + 256 /*barriers*/;
};

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == 1) return gelu_erf(v);
  if (act == 2) return fmaxf(v, 0.f);
  if (act == 3) return fminf(fmaxf(v, 0.f), 6.f);
  return v;
}

template <int BN>
__global__ void __maxnreg__(192)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmC, const GemmEpi ep) {
  using Cfg = GemmCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  // keep the pointer derived from the __shared__ array (so loads compile to LDS, not generic LD)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* stage_out = smem + STAGES * Cfg::STAGE_BYTES;                 // 2 x [128 rows x 128 B] store panels
  uint64_t* bars = reinterpret_cast<uint64_t*>(stage_out + 2 * 16384);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* tfull = bars + 2 * STAGES;
  uint64_t* tempty = bars + 2 * STAGES + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_m = (ep.M + Cfg::BM - 1) / Cfg::BM;
  const int num_n = (ep.N + BN - 1) / BN;
  const int num_k = (ep.K + Cfg::BK - 1) / Cfg::BK;
  const int num_tiles = num_m * num_n;

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 8); }
    fence_barrier_init();
  }
  if (warp == 9) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 8) {
    // ---------------- TMA producer ----------------
    if (lane == 0) {
      int s = 0; uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = tile / num_n, n_blk = tile % num_n;
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(&empty[s], ph ^ 1);
          mbar_arrive_expect_tx(&full[s], Cfg::STAGE_BYTES);
          uint8_t* sa = smem + s * Cfg::STAGE_BYTES;
          tma_load_2d(sa, &tmA, &full[s], kb * Cfg::BK, m_blk * Cfg::BM);
          tma_load_2d(sa + Cfg::A_BYTES, &tmB, &full[s], kb * Cfg::BK, n_blk * BN);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 9) {
    // ---------------- MMA issuer ----------------
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(Cfg::BM, BN, 0, 0);
      int s = 0; uint32_t ph = 0; int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        const uint32_t aph = (it >> 1) & 1;
        mbar_wait(&tempty[acc], aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + s * Cfg::STAGE_BYTES);
          const uint32_t b_addr = a_addr + Cfg::A_BYTES;
#pragma unroll
          for (int k = 0; k < Cfg::BK / 16; ++k) {
            umma_f16_ss(d_tmem, umma_desc_sw128(a_addr + k * 32), umma_desc_sw128(b_addr + k * 32),
                        idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty[s]);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        umma_commit(&tfull[acc]);
      }
    }
  } else {
    // ---------------- epilogue (warps 0..7) ----------------
    const int quad = warp & 3;   // TMEM lane quadrant this warp may touch
    const int half = warp >> 2;  // which half of the BN columns
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int m_blk = tile / num_n, n_blk = tile % num_n;
      const int acc = it & 1;
      const uint32_t aph = (it >> 1) & 1;
      mbar_wait(&tfull[acc], aph);
      tc_fence_after();
      const int row = m_blk * Cfg::BM + quad * 32 + lane;
      const bool row_ok = row < ep.M;
      long long dst_row = row;
      int ps_y = 0, ps_x = 0, ps_b = 0;
      if (ep.row_mode == 1) {
        dst_row = row_ok ? (long long)ep.row_map[row] : -1;
      } else if (ep.row_mode == 2) {
        const int hw = ep.ps_h * ep.ps_w;
        ps_b = row / hw;
        const int r = row - ps_b * hw;
        ps_y = r / ep.ps_w;
        ps_x = r - ps_y * ep.ps_w;
      }
      // All TMEM loads of this warp's column half are issued up front (BN/2 <= 128 fp32 columns =
      // BN/2 registers) and the accumulator is handed back to the MMA warp right after they land, so the
      // global-memory part of the epilogue (bias / residual loads, stores) overlaps the next tile's MMAs
      // instead of sitting between them.
      constexpr int NCH = BN / 64;  // 32-column chunks per warp
      uint32_t r[NCH][32];
#pragma unroll
      for (int ci = 0; ci < NCH; ++ci)
        tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(quad * 32) << 16) + acc * BN + half * (BN / 2) + ci * 32, r[ci]);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
      if (ep.tma_store) {
        // ---- panel-staged TMA tensor store ----
        // A thread owns one accumulator ROW; storing rows straight to global costs one L2 request per 16 B
        // (measured: 0.95 TB/s, 35 % of the qkv GEMM). Instead the 4 warps of a column half assemble a
        // [128 rows x 128 B] panel in shared memory (SWIZZLE_128B layout: conflict-free 16 B stores) and one
        // thread issues a single TMA tensor store for it; the hardware also clips rows >= M / cols >= N.
        uint8_t* panel = stage_out + half * 16384;
        const int rloc = quad * 32 + lane;
        const bool leader = (warp & 3) == 0 && lane == 0;
        auto run_panels = [&](auto pc_tag) {
          constexpr int PC = decltype(pc_tag)::value;   // columns per 128-byte panel row (64 bf16 / 32 fp32)
          constexpr bool F32 = PC == 32;
#pragma unroll
          for (int pi = 0; pi < (BN / 2) / PC; ++pi) {
            const int col0 = n_blk * BN + half * (BN / 2) + pi * PC;
            if (col0 >= ep.N || (ep.dbg & 4)) break;      // uniform over the 4 warps of this half
            uint4 pk[8];
#pragma unroll
            for (int cc = 0; cc < PC / 32; ++cc) {
              constexpr int dummy = 0; (void)dummy;
              const int ci = pi * (PC / 32) + cc;          // compile-time after unrolling
              const int col = col0 + cc * 32;
              float v[32];
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[ci][j]);
              if (col < ep.N) {
                const bool colfull = col + 32 <= ep.N;
                if (ep.bias && !(ep.dbg & 2)) {
#pragma unroll
                  for (int j = 0; j < 32; j += 4) {
                    float4 b;
                    if (colfull) b = __ldg(reinterpret_cast<const float4*>(ep.bias + col + j));
                    else {
                      b.x = col + j < ep.N ? ep.bias[col + j] : 0.f; b.y = col + j + 1 < ep.N ? ep.bias[col + j + 1] : 0.f;
                      b.z = col + j + 2 < ep.N ? ep.bias[col + j + 2] : 0.f; b.w = col + j + 3 < ep.N ? ep.bias[col + j + 3] : 0.f;
                    }
                    v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
                  }
                }
                if (ep.act) {
#pragma unroll
                  for (int j = 0; j < 32; ++j) v[j] = apply_act(v[j], ep.act);
                }
                if (ep.scale) {
#pragma unroll
                  for (int j = 0; j < 32; ++j) v[j] *= (colfull || col + j < ep.N) ? __ldg(ep.scale + col + j) : 0.f;
                }
                if (ep.residual && row_ok) {
                  if (colfull) {
                    const uint4* rp = reinterpret_cast<const uint4*>(ep.residual + (long long)row * ep.ldr + col);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                      float f[8];
                      unpack8(__ldg(rp + j), f);
#pragma unroll
                      for (int k = 0; k < 8; ++k) v[8 * j + k] += f[k];
                    }
                  } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                      if (col + j < ep.N) v[j] += __bfloat162float(ep.residual[(long long)row * ep.ldr + col + j]);
                  }
                }
              }
              if constexpr (F32) {
#pragma unroll
                for (int j = 0; j < 8; ++j)
                  pk[j] = make_uint4(__float_as_uint(v[4 * j]), __float_as_uint(v[4 * j + 1]), __float_as_uint(v[4 * j + 2]),
                                     __float_as_uint(v[4 * j + 3]));
              } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) pk[cc * 4 + j] = pack8(v + 8 * j);
              }
            }
            if (leader) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // previous panel drained
            asm volatile("bar.sync %0, 128;" ::"r"(1 + half) : "memory");
#pragma unroll
            for (int j = 0; j < 8; ++j)
              *reinterpret_cast<uint4*>(panel + rloc * 128 + ((j ^ (rloc & 7)) << 4)) = pk[j];
            fence_proxy_async();
            asm volatile("bar.sync %0, 128;" ::"r"(1 + half) : "memory");
            if (leader && !(ep.dbg & 1)) {
              asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                               reinterpret_cast<uint64_t>(&tmC)),
                           "r"(smem_u32(panel)), "r"(col0), "r"(m_blk * Cfg::BM)
                           : "memory");
              asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
          }
        };
        if (ep.out_f32) run_panels(std::integral_constant<int, 32>{});
        else run_panels(std::integral_constant<int, 64>{});
        continue;   // next tile
      }
#pragma unroll
      for (int ci = 0; ci < NCH; ++ci) {
        const int col_l = half * (BN / 2) + ci * 32;
        const int col = n_blk * BN + col_l;
        if (!row_ok || col >= ep.N || (ep.dbg & 4)) continue;
        long long drow = dst_row;
        int dcol = col;
        if (ep.row_mode == 2) {
          const int sub = col / ep.ps_c;
          dcol = col - sub * ep.ps_c;
          drow = ((long long)ps_b * 2 * ep.ps_h + 2 * ps_y + (sub >> 1)) * (2 * ep.ps_w) + 2 * ps_x + (sub & 1);
        }
        if (drow < 0) continue;
        float v[32];
        const bool full32 = ep.vec_ok && col + 32 <= ep.N;
        if (full32) {
          uint4 rres[4];
          if (ep.residual) {
            const uint4* rp = reinterpret_cast<const uint4*>(ep.residual + drow * ep.ldr + dcol);
#pragma unroll
            for (int j = 0; j < 4; ++j) rres[j] = __ldg(rp + j);
          }
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ep.bias && !(ep.dbg & 2)) b = __ldg(reinterpret_cast<const float4*>(ep.bias + col + j));
            v[j] = __uint_as_float(r[ci][j]) + b.x;
            v[j + 1] = __uint_as_float(r[ci][j + 1]) + b.y;
            v[j + 2] = __uint_as_float(r[ci][j + 2]) + b.z;
            v[j + 3] = __uint_as_float(r[ci][j + 3]) + b.w;
          }
          if (ep.act) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = apply_act(v[j], ep.act);
          }
          if (ep.scale) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 sc = __ldg(reinterpret_cast<const float4*>(ep.scale + col + j));
              v[j] *= sc.x; v[j + 1] *= sc.y; v[j + 2] *= sc.z; v[j + 3] *= sc.w;
            }
          }
          if (ep.residual) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float f[8];
              unpack8(rres[j], f);
#pragma unroll
              for (int k = 0; k < 8; ++k) v[8 * j + k] += f[k];
            }
          }
          if (ep.dbg & 1) {
            // perf-debug: keep all the math (the stores below stay reachable), skip only the stores
          } else if (ep.out_f32) {
            float4* op = reinterpret_cast<float4*>(reinterpret_cast<float*>(ep.out) + drow * ep.ldo + dcol);
#pragma unroll
            for (int j = 0; j < 8; ++j) op[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          } else {
            uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(ep.out) + drow * ep.ldo + dcol);
#pragma unroll
            for (int j = 0; j < 4; ++j) op[j] = pack8(v + 8 * j);
          }
        } else {
          // ragged last column chunk / unaligned rows: scalar, fully guarded
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (col + j < ep.N) {
              float x = __uint_as_float(r[ci][j]);
              if (ep.bias) x += ep.bias[col + j];
              x = apply_act(x, ep.act);
              if (ep.scale) x *= ep.scale[col + j];
              if (ep.residual) x += __bfloat162float(ep.residual[drow * ep.ldr + dcol + j]);
              if (ep.out_f32) reinterpret_cast<float*>(ep.out)[drow * ep.ldo + dcol + j] = x;
              else reinterpret_cast<__nv_bfloat16*>(ep.out)[drow * ep.ldo + dcol + j] = __float2bfloat16_rn(x);
            }
          }
        }
      }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // this thread's TMA stores (if any) are complete
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

template <int BN>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const GemmEpi& ep,
                       int max_ctas, cudaStream_t st) {
  using Cfg = GemmCfg<BN>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_bf16_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  const int num_tiles = ((ep.M + 127) / 128) * ((ep.N + BN - 1) / BN);
  int grid = num_tiles < max_ctas ? num_tiles : max_ctas;
  gemm_bf16_kernel<BN><<<grid, 320, Cfg::SMEM_BYTES, st>>>(tmA, tmB, tmC, ep);
  MMSAM_LAUNCH_CHECK();
  return MMSAM_OK;
}

}  // namespace mmsam

namespace mmsam_host {
EncodeTiledFn get_encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = (EncodeTiledFn)p;
  }
  return fn;
}
int make_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                      uint32_t box_rows, uint32_t box_cols) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) return MMSAM_ERR_DRIVER;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? MMSAM_OK : MMSAM_ERR_DRIVER;
}
}  // namespace mmsam_host

// See include/mmsam_b200.h for the contract.
MMSAM_API int mmsam_gemm_bf16(const void* A, long long lda, const void* W, long long ldw, const float* bias,
                              const float* scale, const void* residual, long long ldr, void* out,
                              long long ldo, int M, int N, int K, int act, int out_f32, int row_mode,
                              const int* row_map_dev, int ps_h, int ps_w, int ps_c, int block_n,
                              int max_ctas, void* stream) {
  using namespace mmsam;
  if (M < 0 || N < 0 || K <= 0) return MMSAM_ERR_BAD_ARG;
  if (M == 0 || N == 0) return MMSAM_OK;
  if (!A || !W || !out) return MMSAM_ERR_BAD_ARG;
  if ((K & 7) || (lda & 7) || (ldw & 7) || lda < K || ldw < K) return MMSAM_ERR_BAD_ARG;
  if ((((uintptr_t)A | (uintptr_t)W) & 15)) return MMSAM_ERR_BAD_ARG;
  if (act < 0 || act > 3 || row_mode < 0 || row_mode > 2) return MMSAM_ERR_BAD_ARG;
  if (row_mode == 1 && !row_map_dev) return MMSAM_ERR_BAD_ARG;
  if (row_mode == 2 && (ps_h <= 0 || ps_w <= 0 || ps_c <= 0 || (ps_c & 31) || N != 4 * ps_c || M % (ps_h * ps_w)))
    return MMSAM_ERR_BAD_ARG;
  // the vector epilogue needs 16-byte aligned rows; ragged N tails fall back to scalar stores
  const int oelt = out_f32 ? 4 : 2;
  int vec_ok = 1;
  if ((((uintptr_t)out) & 15) || ((ldo * oelt) & 15)) vec_ok = 0;
  if (residual && ((((uintptr_t)residual) & 15) || (ldr & 7))) vec_ok = 0;
  if (bias && (((uintptr_t)bias) & 15)) return MMSAM_ERR_BAD_ARG;
  if (scale && (((uintptr_t)scale) & 15)) return MMSAM_ERR_BAD_ARG;
  if (max_ctas <= 0 || max_ctas > kNumSMs) max_ctas = kNumSMs;

  int bn = block_n;
  if (bn != 64 && bn != 128 && bn != 256) {
    // auto: widest tile that still gives every SM a tile, then prefer the least padded N
    const long long num_m = (M + 127) / 128;
    bn = 256;
    if (N <= 64) bn = 64;
    else if (N <= 128) bn = 128;
    else if (num_m * ((N + 255) / 256) < kNumSMs && num_m * ((N + 127) / 128) >= num_m * ((N + 255) / 256) * 2 - 1) bn = 128;
    if (bn == 256 && (N % 256) != 0 && (N % 256) <= 128 && N < 1024) bn = 128;
  }
  CUtensorMap tmA, tmB;
  int rc = mmsam_host::make_tmap_2d_bf16(&tmA, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, 128, 64);
  if (rc) return rc;
  rc = mmsam_host::make_tmap_2d_bf16(&tmB, W, (uint64_t)N, (uint64_t)K, (uint64_t)ldw, (uint32_t)bn, 64);
  if (rc) return rc;
  GemmEpi ep;
  ep.bias = bias; ep.scale = scale; ep.residual = (const __nv_bfloat16*)residual; ep.out = out;
  ep.row_map = row_map_dev; ep.ldo = ldo; ep.ldr = ldr; ep.M = M; ep.N = N; ep.K = K; ep.act = act;
  ep.out_f32 = out_f32; ep.row_mode = row_mode; ep.vec_ok = vec_ok; { const char* d = getenv("MMSAM_GEMM_DBG"); ep.dbg = d ? atoi(d) : 0; } ep.ps_h = ps_h; ep.ps_w = ps_w; ep.ps_c = ps_c;
  cudaStream_t st = (cudaStream_t)stream;
  // output tensor map for the TMA-store epilogue (plain row mapping, 16-byte aligned rows)
  CUtensorMap tmC = tmA;
  ep.tma_store = 0;
  {
    const char* e = getenv("MMSAM_GEMM_TMA_STORE");
    const bool want = !(e && e[0] == '0');
    if (want && row_mode == 0 && vec_ok) {
      mmsam_host::EncodeTiledFn enc = mmsam_host::get_encode_tiled();
      cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)M};
      cuuint64_t strides[1] = {(cuuint64_t)ldo * oelt};
      cuuint32_t box[2] = {(cuuint32_t)(out_f32 ? 32 : 64), 128};
      cuuint32_t estr[2] = {1, 1};
      if (enc && enc(&tmC, out_f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, out, dims, strides,
                     box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS)
        ep.tma_store = 1;
    }
  }
  if (bn == 256) return launch_gemm<256>(tmA, tmB, tmC, ep, max_ctas, st);
  if (bn == 128) return launch_gemm<128>(tmA, tmB, tmC, ep, max_ctas, st);
  return launch_gemm<64>(tmA, tmB, tmC, ep, max_ctas, st);
}
