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
// Design: persistent CTA PAIRS (cluster 2x1x1, one pair per TPC, 74 pairs), 320 threads per CTA.
//   A pair owns a 256 x BN output tile and drives it with tcgen05.mma.cta_group::2 (M=256): each CTA
//   stages its own 128 rows of A and only HALF of the BN weight rows, the tensor core reads the
//   other half out of the peer's shared memory. Per SM that is 32 KB of operands per 64-deep k-step
//   instead of 48 KB, which is what takes the kernel off the L2 -> SM bandwidth limit the 1-CTA
//   128x256 version sat on (profiles/r01_gemm_notes.md).
//   warp 8      TMA producer (one lane): 128x64 A box + (BN/2)x64 W box per stage, SWIZZLE_128B,
//               complete_tx on the LEADER CTA's full barrier; OOB rows/cols are zero-filled so ragged
//               M/N/K need no masks.
//   warp 9      MMA issuer (leader CTA, one lane): 4 x tcgen05.mma (M256 x N=BN x K16) per stage;
//               tcgen05.commit (multicast to both CTAs) frees the smem stage / publishes the
//               accumulator. Two TMEM accumulators (2 x BN columns): the epilogue of tile i overlaps
//               the main loop of tile i+1.
//   warps 0-7   epilogue, two warps per TMEM lane quadrant (one per half of the BN columns). Per
//               128-byte-wide panel: tcgen05.ld -> bias / activation (exact-erf GELU, ReLU, ReLU6) /
//               per-channel scale / residual -> bf16 or fp32 -> the warp's private swizzled 32-row
//               slab in shared memory -> ONE TMA tensor store per slab (plain outputs), or coalesced
//               128-byte row segments through a row map (window un-partition, c2|c3|c4 packing) or a
//               2x2 pixel shuffle (ConvTranspose2d k2 s2). A thread never stores its accumulator row
//               straight to global: that costs one L1 wavefront per 16 bytes and was the limiter.
//               Residual rows are fetched the same way (coalesced, through the slab).
// Tiles are walked n-fastest so the pairs of one wave share A tiles through L2 and the weights stay
// L2-resident: A is read from HBM once.
#include "common.cuh"
#include "cg2.cuh"
#include <type_traits>
#include <cstdlib>

namespace mmsam {

struct GemmEpi {
  const float* bias;               // [N] or null
  const float* scale;              // [N] or null, applied after the activation
  const float2* rowstat;           // [M] (mean, rstd) or null: LayerNorm folded in, acc -> rstd * (acc - mean * scale[n]) + bias[n]
  const void* residual;            // indexed at the destination row, or null; bf16, or fp32 in the V_F32_RES variant
  void* out;
  const int* row_map;              // row_mode 1: dst row of each source row (-1 = drop)
  long long ldo, ldr;
  int M, N, K;
  int act;        // 0 none, 1 gelu(erf), 2 relu, 3 relu6
  int out_f32;    // 0: bf16 output, 1: fp32 output
  int row_mode;   // 0 identity, 1 row_map, 2 pixel-shuffle 2x2 (ConvTranspose2d k=2 s=2)
  int dbg;        // perf-debug switches (env MMSAM_GEMM_DBG): 1 skip stores, 2 skip the residual, 4 skip the epilogue math,
                  // 8 skip the MMA issue, 16 skip the TMA loads
  int ps_h, ps_w, ps_c;
  int res_f32;        // the residual is fp32 (generic epilogue / V_F32_RES)
  int w_group_rows;   // > 0: grouped GEMM, rows [g * w_group_rows, (g + 1) * w_group_rows) of A use weight rows [g * N, (g + 1) * N)
};

// V_F32_RES: fp32 output + fp32 residual (in place on the fp32 residual streams: ViT tokens, ConvNeXt feature maps)
enum { V_BF16 = 0, V_BF16_GELU = 1, V_BF16_MAP = 2, V_F32 = 3, V_GENERIC = 4, V_BF16_RES = 5, V_F32_RES = 6 };
// Variants whose residual block is streamed into a third slab with cp.async (see epilogue_fast).
__host__ __device__ constexpr bool var_async_res(int var) { return var == V_BF16_MAP || var == V_BF16_RES || var == V_F32_RES; }

template <int BN, int VAR = V_BF16> struct GemmCfg {
  static constexpr int BM = 128;            // rows per CTA (the pair computes 256)
  static constexpr int BK = 64;
  static constexpr int BNH = BN / 2;        // weight rows staged per CTA
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BNH * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr bool ASYNC_RES = var_async_res(VAR);
  // 227 KB: 5 x 32 KB operand stages + 2 output slabs per warp, or (residual ring) 4 stages + 3 slabs per warp
  static constexpr int STAGES = ASYNC_RES ? (BN == 256 ? 4 : (BN == 128 ? 5 : 6)) : (BN == 256 ? 5 : (BN == 128 ? 6 : 8));
  static constexpr int NUM_EPI_WARPS = 8;
  static constexpr int SLAB_BYTES = 32 * 128;   // 32 rows x 128 B
  static constexpr int SLABS_PER_WARP = ASYNC_RES ? 3 : 2;
  static constexpr int STAGING_BYTES = NUM_EPI_WARPS * SLABS_PER_WARP * SLAB_BYTES;
  static constexpr int TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STAGING_BYTES + 1024 /*align*/ + 256 /*barriers*/;
  static constexpr int THREADS = 32 * (2 + NUM_EPI_WARPS);
};

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == 1) return gelu_erf(v);
  if (act == 2) return fmaxf(v, 0.f);
  if (act == 3) return fminf(fmaxf(v, 0.f), 6.f);
  return v;
}

__device__ __forceinline__ float4 ldg4_guard(const float* p, int col, int N) {   // N % 4 == 0 on the fast paths
  return col < N ? __ldg(reinterpret_cast<const float4*>(p + col)) : make_float4(0.f, 0.f, 0.f, 0.f);
}

// Slow, fully general epilogue (unaligned rows, N not a whole number of 16-byte chunks, fp32 + residual,
// fp32 + row map ...): thread = accumulator row, guarded scalar stores. Small outputs only.
template <int BN>
__device__ __forceinline__ void epilogue_generic(const GemmEpi& ep, uint32_t tmem_base, uint64_t* tfull, uint64_t* tempty,
                                                 int warp, int lane, uint32_t rank, int pair, int num_pairs, int num_tiles,
                                                 int num_n) {
  const int quad = warp & 3, half = warp >> 2;
  int it = 0;
  for (int tile = pair; tile < num_tiles; tile += num_pairs, ++it) {
    const int mp = tile / num_n, n_blk = tile % num_n;
    const int acc = it & 1;
    mbar_wait(&tfull[acc], (it >> 1) & 1);
    tc_fence_after();
    const int row = mp * 256 + (int)rank * 128 + quad * 32 + lane;
    const bool row_ok = row < ep.M;
    long long drow0 = row_ok ? row : -1;
    if (ep.row_mode == 1) drow0 = row_ok ? ep.row_map[row] : -1;
    else if (ep.row_mode == 2 && row_ok) {
      const int hw = ep.ps_h * ep.ps_w;
      const int b = row / hw, r = row - b * hw;
      const int y = r / ep.ps_w, x = r - y * ep.ps_w;
      drow0 = ((long long)b * 2 * ep.ps_h + 2 * y) * (2 * ep.ps_w) + 2 * x;
    }
    const uint32_t tmem_acc = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * BN;
#pragma unroll 1
    for (int ci = 0; ci < BN / 64; ++ci) {
      const int col_l = half * (BN / 2) + ci * 32;
      const int col = n_blk * BN + col_l;
      uint32_t r[32];
      tmem_ld_32x32b_x32(tmem_acc + col_l, r);
      tmem_ld_wait();
      if (ci == BN / 64 - 1) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(mapa_shared(smem_u32(&tempty[acc]), 0));
      }
      if (drow0 < 0 || col >= ep.N || (ep.dbg & 5)) continue;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int cj = col + j;
        if (cj < ep.N) {
          long long drow = drow0;
          int dcol = cj;
          if (ep.row_mode == 2) {
            const int sub = cj / ep.ps_c;
            dcol = cj - sub * ep.ps_c;
            drow += (sub >> 1) * (2 * ep.ps_w) + (sub & 1);
          }
          float x = __uint_as_float(r[j]);
          if (ep.rowstat) {
            const float2 rs = ep.rowstat[row];
            x = fmaf(rs.y, fmaf(-rs.x, ep.scale[cj], x), ep.bias[cj]);
          } else if (ep.bias) x += ep.bias[cj];
          x = apply_act(x, ep.act);
          if (ep.scale && !ep.rowstat) x *= ep.scale[cj];
          if (ep.residual) x += ep.res_f32 ? reinterpret_cast<const float*>(ep.residual)[drow * ep.ldr + dcol]
                                           : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(ep.residual)[drow * ep.ldr + dcol]);
          if (ep.out_f32) reinterpret_cast<float*>(ep.out)[drow * ep.ldo + dcol] = x;
          else reinterpret_cast<__nv_bfloat16*>(ep.out)[drow * ep.ldo + dcol] = __float2bfloat16_rn(x);
        }
      }
    }
  }
}

__device__ __forceinline__ void cp_async16(uint32_t smem_dst, const void* src, int src_bytes) {   // L2 -> smem, no L1 line
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_dst), "l"(src), "r"(src_bytes) : "memory");
}

// Fast epilogues. One warp = one TMEM lane quadrant (32 accumulator rows) x one half of the BN columns, walked
// in 128-byte-wide panels (64 bf16 / 32 fp32 columns); the panel sequence of a warp is static, so the residual
// block of the NEXT panel (possibly of a later tile) is always in flight while the current one is computed:
//   * V_BF16_RES / V_BF16_MAP: cp.async.cg straight into a two-slab ring (no L1 line, no registers: with ~225 KB
//     of the SM's 228 KB carved out as shared memory the L1 can track only a handful of outstanding LDG misses,
//     measured 2 TB/s) + a third slab for the output;
//   * V_BF16 (large K: the epilogue has time): LDG into registers, two in-place slabs, one more operand stage.
// A whole tile ahead every lane also pulls the 128-byte residual line of its own row into L2.
template <int BN, int VAR>
__device__ __forceinline__ void epilogue_fast(const GemmEpi& ep, const CUtensorMap* tmC, uint32_t tmem_base, uint32_t slab0,
                                              uint64_t* tfull, uint64_t* tempty, int warp, int lane, uint32_t rank, int pair,
                                              int num_pairs, int num_tiles, int num_n) {
  constexpr bool F32 = VAR == V_F32 || VAR == V_F32_RES;
  constexpr bool MAPPED = VAR == V_BF16_MAP;
  constexpr bool ASYNC = var_async_res(VAR);
  constexpr uint32_t SLAB = GemmCfg<BN, VAR>::SLAB_BYTES;
  constexpr int PC = F32 ? 32 : 64;                      // columns per 128-byte panel row
  constexpr int CPH = (BN / 2 >= PC) ? BN / 2 : PC;      // columns per warp-half
  constexpr int NPAN = CPH / PC;
  constexpr int EPC = F32 ? 4 : 8;                       // elements per 16-byte chunk
  const int quad = warp & 3, half = warp >> 2;
  const bool has_res = VAR != V_F32 && ep.residual != nullptr && !(ep.dbg & 2);
  const int c = lane & 7, rsub = lane >> 3;
  if (half * CPH >= BN) {
    // BN == 64 with bf16 output: the second column half has no panel; only hand the accumulators back
    int it = 0;
    for (int tile = pair; tile < num_tiles; tile += num_pairs, ++it) {
      mbar_wait(&tfull[it & 1], (it >> 1) & 1);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_shared(smem_u32(&tempty[it & 1]), 0));
    }
    return;
  }
  auto first_col = [&](int t) { return (t % num_n) * BN + half * CPH; };
  // first tile at or after t in which this warp has a panel
  auto next_valid = [&](int t) {
    while (t < num_tiles && first_col(t) >= ep.N) t += num_pairs;
    return t;
  };
  // destination row of this thread's accumulator row in tile t (mode 2: row of the (dy,dx)=(0,0) sub-pixel)
  auto dst_of = [&](int t) -> int {
    if (t >= num_tiles) return -1;
    const int row = (t / num_n) * 256 + (int)rank * 128 + quad * 32 + lane;
    if (row >= ep.M) return -1;
    if (!MAPPED) return row;
    if (ep.row_mode == 1) return __ldg(ep.row_map + row);
    if (ep.row_mode == 2) {
      const int hw = ep.ps_h * ep.ps_w;
      const int b = row / hw, r = row - b * hw;
      const int y = r / ep.ps_w, x = r - y * ep.ps_w;
      return (b * 2 * ep.ps_h + 2 * y) * (2 * ep.ps_w) + 2 * x;
    }
    return row;
  };
  uint4 rv[8];
  // coalesced fetch of a 32 x 128 B residual block: 8 lanes per row, 4 rows per instruction
  auto issue_res = [&](int dst, int col0, uint32_t ring_slab) {
    const int colc = col0 + c * EPC;
    int dcolc = colc, sub = 0;
    if (MAPPED && ep.row_mode == 2) { sub = colc / ep.ps_c; dcolc = colc - sub * ep.ps_c; }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int rl = i * 4 + rsub;
      int d = __shfl_sync(0xffffffffu, dst, rl);
      if (MAPPED && ep.row_mode == 2 && d >= 0) d += (sub >> 1) * (2 * ep.ps_w) + (sub & 1);
      const bool ok = d >= 0 && colc < ep.N;
      const char* src = reinterpret_cast<const char*>(ep.residual) + (ok ? ((long long)d * ep.ldr + dcolc) * (F32 ? 4 : 2) : 0);
      if constexpr (ASYNC) {
        cp_async16(ring_slab + swz128(rl, c), src, ok ? 16 : 0);
      } else {
        rv[i] = make_uint4(0u, 0u, 0u, 0u);
        if (ok) rv[i] = __ldg(reinterpret_cast<const uint4*>(src));
      }
    }
    if constexpr (ASYNC) asm volatile("cp.async.commit_group;" ::: "memory");
  };
  auto prefetch_res_l2 = [&](int dst, int t) {
    if (dst < 0) return;
#pragma unroll
    for (int pi = 0; pi < NPAN; ++pi) {
      const int col0 = first_col(t) + pi * PC;
      if (col0 >= ep.N) break;
      long long d = dst;
      int dcol = col0;
      if (MAPPED && ep.row_mode == 2) {
        const int sub = col0 / ep.ps_c;
        dcol = col0 - sub * ep.ps_c;
        d += (sub >> 1) * (2 * ep.ps_w) + (sub & 1);
      }
      const char* p = reinterpret_cast<const char*>(ep.residual) + (d * ep.ldr + dcol) * (F32 ? 4 : 2);
      asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
      if (reinterpret_cast<uintptr_t>(p) & 127) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + 126));
    }
  };

  int sel = 0, it = 0;     // sel: non-ASYNC: which in-place slab; ASYNC: which ring slab holds the current residual block
  if (has_res) {
    const int t0 = next_valid(pair);
    if (t0 < num_tiles) issue_res(dst_of(t0), first_col(t0), slab0);
  }
  for (int tile = pair; tile < num_tiles; tile += num_pairs, ++it) {
    const int mp = tile / num_n, n_blk = tile % num_n;
    const int acc = it & 1;
    const int my_dst = (MAPPED || has_res) ? dst_of(tile) : -1;
    // the next tile in which this warp has work: its residual rows go to L2 now, its first block is requested
    // by the last panel of this tile
    const int nt = next_valid(tile + num_pairs);
    const int my_dst_nt = has_res ? dst_of(nt) : -1;
    if (has_res && nt < num_tiles && !(ep.dbg & 32)) prefetch_res_l2(my_dst_nt, nt);
    mbar_wait(&tfull[acc], (it >> 1) & 1);
    tc_fence_after();
    const uint32_t tempty_addr = mapa_shared(smem_u32(&tempty[acc]), 0);
    bool released = false;
    auto release = [&]() {
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(tempty_addr);
      released = true;
    };
    const int row0 = mp * 256 + (int)rank * 128 + quad * 32;
    const uint32_t tmem_acc = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * BN;
    float2 rs = make_float2(0.f, 1.f);          // this thread's accumulator row: (mean, rstd) of the folded LayerNorm
    if (ep.rowstat && row0 + lane < ep.M) rs = __ldg(ep.rowstat + row0 + lane);
#pragma unroll
    for (int pi = 0; pi < NPAN; ++pi) {
      const int col_l = half * CPH + pi * PC;
      const int col0 = n_blk * BN + col_l;
      if (col0 >= ep.N) break;                            // warp-uniform
      const bool more_here = pi + 1 < NPAN && col0 + PC < ep.N;
      const bool have_next = has_res && (more_here || nt < num_tiles);
      // rslab: where this panel's residual block is (or, register path, will be put); oslab: where the output goes
      const uint32_t rslab = slab0 + (uint32_t)sel * SLAB;
      const uint32_t oslab = ASYNC ? slab0 + 2 * SLAB : rslab;
      const uint32_t nslab = slab0 + (uint32_t)(sel ^ 1) * SLAB;
      sel ^= 1;
      if constexpr (!ASYNC) {
        // the TMA store that last read this slab (two panels ago) must have drained it
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        __syncwarp();
        if (has_res) {
#pragma unroll
          for (int i = 0; i < 8; ++i) sts128(rslab + swz128(i * 4 + rsub, c), rv[i]);
        }
      } else {
        __syncwarp();                                     // every lane is done with the ring slab being refilled
      }
      if (have_next) {
        if (more_here) issue_res(my_dst, col0 + PC, nslab);
        else issue_res(my_dst_nt, first_col(nt), nslab);
      }
      uint32_t r[PC / 32][32];
#pragma unroll
      for (int cc = 0; cc < PC / 32; ++cc) tmem_ld_32x32b_x32(tmem_acc + col_l + cc * 32, r[cc]);
      tmem_ld_wait();
      if (!more_here) release();
      if (has_res) {
        if constexpr (ASYNC) {
          if (have_next) asm volatile("cp.async.wait_group 1;" ::: "memory");
          else asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncwarp();                                     // residual block visible to the row owners
      }
      if (ep.dbg & 4) continue;

      uint4 pk[8];
#pragma unroll
      for (int cc = 0; cc < PC / 32; ++cc) {
        const int col = col0 + cc * 32;
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[cc][j]);
        if (ep.rowstat) {
          const float nm = -rs.x;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 sc = ldg4_guard(ep.scale, col + j, ep.N);
            const float4 b = ldg4_guard(ep.bias, col + j, ep.N);
            v[j] = fmaf(rs.y, fmaf(nm, sc.x, v[j]), b.x);
            v[j + 1] = fmaf(rs.y, fmaf(nm, sc.y, v[j + 1]), b.y);
            v[j + 2] = fmaf(rs.y, fmaf(nm, sc.z, v[j + 2]), b.z);
            v[j + 3] = fmaf(rs.y, fmaf(nm, sc.w, v[j + 3]), b.w);
          }
        } else if (ep.bias) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 b = ldg4_guard(ep.bias, col + j, ep.N);
            v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
          }
        }
        if (VAR == V_BF16_GELU) {
#pragma unroll
          for (int j = 0; j < 32; j += 2) gelu_erf2(v[j], v[j + 1]);
        } else if (ep.act == 2) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
        } else if (ep.act == 3) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fminf(fmaxf(v[j], 0.f), 6.f);
        }
        if (ep.scale && !ep.rowstat) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 s = ldg4_guard(ep.scale, col + j, ep.N);
            v[j] *= s.x; v[j + 1] *= s.y; v[j + 2] *= s.z; v[j + 3] *= s.w;
          }
        }
        if (has_res) {
          if constexpr (F32) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const uint4 q = lds128(rslab + swz128(lane, j));
              v[4 * j] += __uint_as_float(q.x); v[4 * j + 1] += __uint_as_float(q.y);
              v[4 * j + 2] += __uint_as_float(q.z); v[4 * j + 3] += __uint_as_float(q.w);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float f[8];
              unpack8(lds128(rslab + swz128(lane, cc * 4 + j)), f);
#pragma unroll
              for (int k = 0; k < 8; ++k) v[8 * j + k] += f[k];
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
      if constexpr (ASYNC && !MAPPED) {
        // single output slab: the previous panel's TMA store (issued a whole panel ago) must have read it
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
      }
      // own accumulator row -> swizzled slab (conflict-free 16-byte stores)
#pragma unroll
      for (int j = 0; j < 8; ++j) sts128(oslab + swz128(lane, j), pk[j]);
      if constexpr (!MAPPED) {
        fence_proxy_async();
        __syncwarp();
        if (lane == 0 && !(ep.dbg & 1)) {
          asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                           reinterpret_cast<uint64_t>(tmC)),
                       "r"(oslab), "r"(col0), "r"(row0)
                       : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      } else {
        __syncwarp();
        // coalesced row segments: 8 lanes write one 128-byte row piece, 4 rows per instruction
        const int colc = col0 + c * EPC;
        int dcolc = colc, sub = 0;
        if (ep.row_mode == 2) { sub = colc / ep.ps_c; dcolc = colc - sub * ep.ps_c; }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rl = i * 4 + rsub;
          int d = __shfl_sync(0xffffffffu, my_dst, rl);
          if (ep.row_mode == 2 && d >= 0) d += (sub >> 1) * (2 * ep.ps_w) + (sub & 1);
          const uint4 val = lds128(oslab + swz128(rl, c));
          if (d >= 0 && colc < ep.N && !(ep.dbg & 1))
            *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(ep.out) + (long long)d * ep.ldo + dcolc) = val;
        }
        __syncwarp();                                     // the output slab is rewritten by the next panel
      }
    }
    if (!released) release();
  }
  if (!MAPPED && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // this warp's TMA stores are complete
  __syncwarp();
}

template <int BN, int VAR>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GemmCfg<BN, VAR>::THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmC, const GemmEpi ep) {
  using Cfg = GemmCfg<BN, VAR>;
  constexpr int STAGES = Cfg::STAGES;
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  // keep the pointer derived from the __shared__ array (so loads compile to LDS, not generic LD)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* staging = smem + STAGES * Cfg::STAGE_BYTES;                    // 8 warps x 2 slabs x 4 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging + Cfg::STAGING_BYTES);
  uint64_t* full = bars;                     // used in the leader CTA only (both CTAs' TMA bytes land here)
  uint64_t* empty = bars + STAGES;           // per CTA, armed by the multicast commit
  uint64_t* tfull = bars + 2 * STAGES;       // per CTA, armed by the multicast commit
  uint64_t* tempty = bars + 2 * STAGES + 2;  // used in the leader CTA only (16 epilogue warps arrive)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int num_mp = (ep.M + 2 * Cfg::BM - 1) / (2 * Cfg::BM);
  const int num_n = (ep.N + BN - 1) / BN;
  const int num_k = (ep.K + Cfg::BK - 1) / Cfg::BK;
  const int num_tiles = num_mp * num_n;
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmC);
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 2); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 2 * Cfg::NUM_EPI_WARPS); }
    fence_barrier_init();
  }
  cluster_sync_all();                         // both CTAs are resident before the paired TMEM allocation
  if (warp == 9) tmem_alloc_cg2(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before();
  cluster_sync_all();                         // barrier inits + TMEM address visible cluster-wide
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();          // everything above overlapped the previous kernel's tail; global memory is touched only from here on

  // The two single-lane roles take the HIGHEST warp ids: the SMSP arbiter prefers high warp ids, so the TMA and
  // MMA issue slots are never starved by epilogue math running on the same scheduler.
  if (warp == 8) {
    // ---------------- TMA producer (both CTAs): the whole warp runs the loop, one elected lane issues (see elect_one) ----------------
    {
      int s = 0; uint32_t ph = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        const int mp = tile / num_n, n_blk = tile % num_n;
        const int m0 = mp * (2 * Cfg::BM) + (int)rank * Cfg::BM;
        // grouped GEMM (per-image weights): the 256-row block lies inside one group (w_group_rows % 256 == 0)
        const int nb0 = n_blk * BN + (int)rank * Cfg::BNH + (ep.w_group_rows > 0 ? (mp * (2 * Cfg::BM) / ep.w_group_rows) * ep.N : 0);
        for (int kb = 0; kb < num_k; ++kb) {
          if (ep.dbg & 16) continue;
          mbar_wait(&empty[s], ph ^ 1);
          const uint32_t lead_full = mapa_shared(smem_u32(&full[s]), 0);
          const uint32_t sa = smem_u32(smem + s * Cfg::STAGE_BYTES);
          if (elect_one()) {
            if (rank == 0) mbar_arrive_expect_tx(&full[s], 2 * Cfg::STAGE_BYTES);
            else mbar_arrive_cluster(lead_full);
            tma_load_2d_cg2(sa, &tmA, lead_full, kb * Cfg::BK, m0);
            tma_load_2d_cg2(sa + Cfg::A_BYTES, &tmB, lead_full, kb * Cfg::BK, nb0);
          }
          __syncwarp();
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 9) {
    // ---------------- MMA issuer (leader CTA only): the whole warp runs the loop, one elected lane issues (see elect_one) ----------------
    if (rank == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(2 * Cfg::BM, BN, 0, 0);
      int s = 0; uint32_t ph = 0; int it = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs, ++it) {
        const int acc = it & 1;
        const uint32_t aph = (it >> 1) & 1;
        mbar_wait(&tempty[acc], aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < num_k; ++kb) {
          if (!(ep.dbg & 16)) mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + s * Cfg::STAGE_BYTES);
          const uint32_t b_addr = a_addr + Cfg::A_BYTES;
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < Cfg::BK / 16; ++k) {
              if (ep.dbg & 8) break;
              umma_f16_ss_cg2(d_tmem, umma_desc_sw128(a_addr + k * 32), umma_desc_sw128(b_addr + k * 32), idesc,
                              (kb | k) != 0 ? 1u : 0u);
            }
            umma_commit_cg2(&empty[s]);
            if (kb == num_k - 1) umma_commit_cg2(&tfull[acc]);
          }
          __syncwarp();
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        if (num_k == 0) {
          if (elect_one()) umma_commit_cg2(&tfull[acc]);
          __syncwarp();
        }
      }
      // the peer's last remote arrivals must have landed before this CTA's barriers can go away
      if (it > 0) {
        const int last = it - 1;
        mbar_wait(&tempty[last & 1], (last >> 1) & 1);
      }
    }
    __syncwarp();
  } else {
    // ---------------- epilogue (warps 0..7, both CTAs) ----------------
    if constexpr (VAR == V_GENERIC) {
      epilogue_generic<BN>(ep, tmem_base, tfull, tempty, warp, lane, rank, pair, num_pairs, num_tiles, num_n);
    } else {
      const uint32_t slab0 = smem_u32(staging) + (uint32_t)warp * Cfg::SLABS_PER_WARP * Cfg::SLAB_BYTES;
      epilogue_fast<BN, VAR>(ep, &tmC, tmem_base, slab0, tfull, tempty, warp, lane, rank, pair, num_pairs, num_tiles, num_n);
    }
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc_cg2(tmem_base, Cfg::TMEM_COLS);
  }
}

template <int BN, int VAR>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const GemmEpi& ep,
                       int max_ctas, cudaStream_t st) {
  using Cfg = GemmCfg<BN, VAR>;
  MMSAM_SET_SMEM_ONCE((gemm_bf16_kernel<BN, VAR>), Cfg::SMEM_BYTES);
  const int num_tiles = ((ep.M + 255) / 256) * ((ep.N + BN - 1) / BN);
  int pairs = max_ctas / 2;
  if (num_tiles < pairs) pairs = num_tiles;
  cudaError_t le = mmsam_host::launch_pdl(gemm_bf16_kernel<BN, VAR>, dim3(2 * pairs), dim3(Cfg::THREADS), Cfg::SMEM_BYTES, st, tmA, tmB, tmC, ep);
  if (le != cudaSuccess) return (int)le;
  MMSAM_LAUNCH_CHECK();
  return MMSAM_OK;
}

template <int VAR>
static int launch_gemm_bn(int bn, const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const GemmEpi& ep,
                          int max_ctas, cudaStream_t st) {
  if (bn == 256) return launch_gemm<256, VAR>(tmA, tmB, tmC, ep, max_ctas, st);
  if (bn == 128) return launch_gemm<128, VAR>(tmA, tmB, tmC, ep, max_ctas, st);
  return launch_gemm<64, VAR>(tmA, tmB, tmC, ep, max_ctas, st);
}

}  // namespace mmsam

namespace mmsam_host {
bool pdl_enabled() {
  // opt-in (MMSAM_PDL=1). Measured on the bench step (tools/launch_gap.py, bench.py A/B in one gpurun call): chains of tcgen05
  // kernels gain 1.4 us per node (5.3 -> 3.9 us), but a full-chip LayerNorm followed by a full-chip GEMM loses ~7 us per pair
  // (the early-launched 227 KB CTAs take SMs the LayerNorm's last wave still wants), and the step as a whole is 0.3 - 0.6 ms
  // SLOWER (59.2 - 59.5 vs 58.9 ms) - so the default stays plain stream order.
  static const bool on = [] { const char* e = getenv("MMSAM_PDL"); return e && e[0] == '1'; }();
  return on;
}
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
static int gemm_dbg_flags() {
  static const int v = [] { const char* d = getenv("MMSAM_GEMM_DBG"); return d ? atoi(d) : 0; }();   // read once
  return v;
}

static int gemm_impl(const void* A, long long lda, const void* W, long long ldw, const float* bias,
                     const float* scale, const float* rowstat, const void* residual, long long ldr, void* out,
                     long long ldo, int M, int N, int K, int act, int out_f32, int row_mode,
                     const int* row_map_dev, int ps_h, int ps_w, int ps_c, int block_n,
                     int max_ctas, void* stream, int group_rows = 0, int res_f32 = 0) {
  using namespace mmsam;
  if (group_rows < 0 || (group_rows > 0 && ((group_rows & 255) || M % group_rows))) return MMSAM_ERR_BAD_ARG;
  const int groups = group_rows > 0 ? M / group_rows : 1;
  if (rowstat && (!bias || !scale || (((uintptr_t)rowstat) & 7))) return MMSAM_ERR_BAD_ARG;
  if (M < 0 || N < 0 || K <= 0) return MMSAM_ERR_BAD_ARG;
  if (M == 0 || N == 0) return MMSAM_OK;
  if (!A || !W || !out) return MMSAM_ERR_BAD_ARG;
  if ((K & 7) || (lda & 7) || (ldw & 7) || lda < K || ldw < K) return MMSAM_ERR_BAD_ARG;
  if ((((uintptr_t)A | (uintptr_t)W) & 15)) return MMSAM_ERR_BAD_ARG;
  if (act < 0 || act > 3 || row_mode < 0 || row_mode > 2) return MMSAM_ERR_BAD_ARG;
  if (row_mode == 1 && !row_map_dev) return MMSAM_ERR_BAD_ARG;
  if (row_mode == 2 && (ps_h <= 0 || ps_w <= 0 || ps_c <= 0 || (ps_c & 31) || N != 4 * ps_c || M % (ps_h * ps_w)))
    return MMSAM_ERR_BAD_ARG;
  // the vector epilogues need 16-byte aligned rows and whole 16-byte chunks; anything else goes scalar
  int vec_ok = 1;
  if ((((uintptr_t)out) & 15) || ((ldo * (out_f32 ? 4 : 2)) & 15)) vec_ok = 0;
  if (residual && ((((uintptr_t)residual) & 15) || (ldr & 7))) vec_ok = 0;
  if (res_f32 && !out_f32) return MMSAM_ERR_UNSUPPORTED;            // an fp32 residual goes with an fp32 output
  if (bias && (((uintptr_t)bias) & 15)) return MMSAM_ERR_BAD_ARG;
  if (scale && (((uintptr_t)scale) & 15)) return MMSAM_ERR_BAD_ARG;
  if (max_ctas <= 1 || max_ctas > kNumSMs) max_ctas = kNumSMs;

  int bn = block_n;
  if (bn != 64 && bn != 128 && bn != 256) {
    // auto: per pair-tile cost = max(main loop, epilogue) summed over the n-tiles of a row block, times the wave
    // quantisation. Units: one 64-deep k-step of a BN=256 tile (MMA-bound, 512 cycles). Narrower tiles are bound by
    // the L2 -> SM operand traffic (~0.8 / ~0.7 of that per k-step); an epilogue panel (32 rows x 128 B per warp)
    // costs ~3 units with GELU, ~2 with a residual, ~1.4 plain, and a ragged last n-tile leaves warps idle.
    const long long num_mp = (M + 255) / 256;
    const long long pairs = kNumSMs / 2;
    const double kb = (double)((K + 63) / 64);
    const double e = act == 1 ? 3.0 : (residual ? 2.0 : 1.4);
    const int pc = out_f32 ? 32 : 64;
    double best = 1e30;
    const int cand[3] = {256, 128, 64};
    const double tcost[3] = {1.0, 0.8, 0.7};
    bn = 256;
    for (int i = 0; i < 3; ++i) {
      const int nn = (N + cand[i] - 1) / cand[i];
      const int cph = cand[i] / 2 >= pc ? cand[i] / 2 : pc;
      double per_mp = 0;
      for (int nb = 0; nb < nn; ++nb) {
        int cols = N - nb * cand[i];
        if (cols > cph) cols = cph;                  // columns of the busiest warp (half 0)
        const double pan = (double)((cols + pc - 1) / pc);
        const double ml = kb * tcost[i];
        per_mp += ml > pan * e ? ml : pan * e;
      }
      const long long tiles = num_mp * nn;
      const double quant = (double)((tiles + pairs - 1) / pairs) * (double)pairs / (double)tiles;
      const double cost = per_mp * quant;
      if (cost < best - 1e-9) { best = cost; bn = cand[i]; }
    }
  }
  // epilogue variant: fast paths need 16-byte aligned rows and whole 16-byte chunks
  const int oelt = out_f32 ? 4 : 2;
  const int epc = out_f32 ? 4 : 8;
  int var;
  if (!vec_ok || (N % epc) != 0 || (out_f32 && ((residual && !res_f32) || row_mode != 0 || act == 1)) || (row_mode != 0 && act == 1))
    var = V_GENERIC;
  else if (out_f32) var = residual ? V_F32_RES : V_F32;
  else if (row_mode != 0) var = V_BF16_MAP;
  else if (act == 1 && !residual) var = V_BF16_GELU;
  else if (act == 1) var = V_GENERIC;
  else var = (residual && K <= 2048) ? V_BF16_RES : V_BF16;   // long K: the epilogue has time, keep the 5th stage
  if (var == V_GENERIC) bn = 128;   // the general epilogue is instantiated for one tile width only
  CUtensorMap tmA, tmB;
  int rc = mmsam_host::make_tmap_2d_bf16(&tmA, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, 128, 64);
  if (rc) return rc;
  rc = mmsam_host::make_tmap_2d_bf16(&tmB, W, (uint64_t)N * groups, (uint64_t)K, (uint64_t)ldw, (uint32_t)(bn / 2), 64);
  if (rc) return rc;
  GemmEpi ep;
  ep.bias = bias; ep.scale = scale; ep.rowstat = (const float2*)rowstat; ep.residual = residual; ep.out = out;
  ep.res_f32 = res_f32;
  ep.row_map = row_map_dev; ep.ldo = ldo; ep.ldr = ldr; ep.M = M; ep.N = N; ep.K = K; ep.act = act;
  ep.out_f32 = out_f32; ep.row_mode = row_mode; ep.ps_h = ps_h; ep.ps_w = ps_w; ep.ps_c = ps_c;
  ep.dbg = gemm_dbg_flags();
  ep.w_group_rows = group_rows;
  cudaStream_t st = (cudaStream_t)stream;
  CUtensorMap tmC = tmA;
  if (var == V_BF16 || var == V_BF16_GELU || var == V_F32 || var == V_BF16_RES || var == V_F32_RES) {
    // output tensor map for the per-warp TMA tensor stores: box = one 32-row x 128-byte slab
    mmsam_host::EncodeTiledFn enc = mmsam_host::get_encode_tiled();
    cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)M};
    cuuint64_t strides[1] = {(cuuint64_t)ldo * oelt};
    cuuint32_t box[2] = {(cuuint32_t)(out_f32 ? 32 : 64), 32};
    cuuint32_t estr[2] = {1, 1};
    if (!enc || enc(&tmC, out_f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, out, dims, strides,
                    box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return MMSAM_ERR_DRIVER;
  }
  switch (var) {
    case V_BF16: return launch_gemm_bn<V_BF16>(bn, tmA, tmB, tmC, ep, max_ctas, st);
    case V_BF16_GELU: return launch_gemm_bn<V_BF16_GELU>(bn, tmA, tmB, tmC, ep, max_ctas, st);
    case V_BF16_MAP: return launch_gemm_bn<V_BF16_MAP>(bn, tmA, tmB, tmC, ep, max_ctas, st);
    case V_F32: return launch_gemm_bn<V_F32>(bn, tmA, tmB, tmC, ep, max_ctas, st);
    case V_BF16_RES: return launch_gemm_bn<V_BF16_RES>(bn, tmA, tmB, tmC, ep, max_ctas, st);
    case V_F32_RES: return launch_gemm_bn<V_F32_RES>(bn, tmA, tmB, tmC, ep, max_ctas, st);
    default: return launch_gemm<128, V_GENERIC>(tmA, tmB, tmC, ep, max_ctas, st);
  }
}

// See include/mmsam_b200.h for the contract.
MMSAM_API int mmsam_gemm_bf16(const void* A, long long lda, const void* W, long long ldw, const float* bias,
                              const float* scale, const void* residual, int res_f32, long long ldr, void* out,
                              long long ldo, int M, int N, int K, int act, int out_f32, int row_mode,
                              const int* row_map_dev, int ps_h, int ps_w, int ps_c, int block_n,
                              int max_ctas, void* stream) {
  return gemm_impl(A, lda, W, ldw, bias, scale, nullptr, residual, ldr, out, ldo, M, N, K, act, out_f32, row_mode,
                   row_map_dev, ps_h, ps_w, ps_c, block_n, max_ctas, stream, 0, res_f32);
}

// LayerNorm folded into the GEMM: out = epilogue(rstd[m] * (A W^T - mean[m] * colsum[n]) + bias[n]); see the header.
MMSAM_API int mmsam_gemm_ln_bf16(const void* A, long long lda, const void* W, long long ldw, const float* bias,
                                 const float* colsum, const float* rowstat, const void* residual, long long ldr,
                                 void* out, long long ldo, int M, int N, int K, int act, int out_f32, int block_n,
                                 int max_ctas, void* stream) {
  if (!rowstat || !bias || !colsum) return MMSAM_ERR_BAD_ARG;
  return gemm_impl(A, lda, W, ldw, bias, colsum, rowstat, residual, ldr, out, ldo, M, N, K, act, out_f32, 0, nullptr,
                   0, 0, 0, block_n, max_ctas, stream);
}

// Grouped GEMM: rows [g * rows_per_group, (g + 1) * rows_per_group) of A are multiplied by their own weight matrix
// W[g] (W is [G * N, K], G = M / rows_per_group), one launch for all groups; see the header.
MMSAM_API int mmsam_gemm_grouped_bf16(const void* A, long long lda, const void* W, long long ldw, const float* bias,
                                      const float* scale, const void* residual, long long ldr, void* out, long long ldo,
                                      int M, int N, int K, int rows_per_group, int act, int block_n, int max_ctas,
                                      void* stream) {
  if (rows_per_group <= 0) return MMSAM_ERR_BAD_ARG;
  return gemm_impl(A, lda, W, ldw, bias, scale, nullptr, residual, ldr, out, ldo, M, N, K, act, 0, 0, nullptr, 0, 0, 0,
                   block_n, max_ctas, stream, rows_per_group);
}
