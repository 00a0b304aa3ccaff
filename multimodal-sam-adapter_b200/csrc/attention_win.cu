// SAM window attention (14 x 14 = 196 tokens per window, head_dim 64) with the decomposed relative-position bias, as a
// SINGLE-PASS softmax on tcgen05 tensor cores. 20 of ViT-L's 24 blocks run this shape: B * 25 windows * 16 heads = 3200
// independent 196 x 196 problems per launch (base/image_encoder.py:399-416, 483-501, 587-623; padded window tokens are
// ordinary keys, they are NOT masked).
//
// The general flash kernel (attention.cu) runs this as 2 x 128-query tiles x 2 x 128-key blocks with an online softmax:
// (196 / 256)^2 = 59 % useful MMA / exp work and two mbarrier hand-shake chains per key block; it measured 231 us per
// launch, 12 % tensor-pipe, against a 50 us HBM floor (qkv in, out written once). Here a whole score row fits one MMA:
//
//   item = (window, head): Q (196 rows, two row tiles A = 128 / B = 68 valid of 72 staged), K, V (196 keys padded to
//   N = 208 by TMA zero fill) in ONE shared-memory stage (77 KB), two stages.
//   warp 8  TMA producer, one lane.
//   warp 9  MMA issuer, one lane. Per row tile: S = Q K^T  (M128 x N208 x K64, accumulator = TMEM slot of the tile) and
//           G = Q table^T (N64: the bias pre-products of both axes), later O = P V with P read from TENSOR MEMORY
//           (tcgen05.mma TS form: the softmax warps overwrite the first 104 columns of S with bf16 P, O accumulates into
//           columns 128..191 of the same slot), so P never touches shared memory.
//   warps 0-3 / 4-7  two softmax groups, thread = query row, group g owns row tile g and TMEM slot g: the tensor core
//           works on one tile (S / G / P V) while the other group's exponentials run. Per row: G -> 28 bias values through
//           a row-private shared-memory gather (the Toeplitz index q - k + 13 is per thread), two passes over the 208
//           score columns straight out of TMEM (max, then exp2 / sum / bf16 P written back with tcgen05.st), one pass
//           over O. No online rescaling, no key blocks, no P panel in shared memory.
// TMEM: slot 0 [0, 208), slot 1 [208, 416), G [416, 480).
#include "common.cuh"
#include "cg2.cuh"
#include <cstdlib>

namespace mmsam {

namespace win {
constexpr int T = 196, KS = 14, NK = 208, D = 64;
constexpr int QA_BYTES = 128 * 128, QB_ROWS = 72, QB_BYTES = QB_ROWS * 128, KV_BYTES = NK * 128;
constexpr int STAGE_BYTES = QA_BYTES + QB_BYTES + 2 * KV_BYTES;          // 78848 = 77 KB, every tile 1024-byte aligned
constexpr int TAB_BYTES = 64 * 128;
constexpr int G_IDX = 56;                                                // bias pre-products kept per row: 28 (h) + 28 (w)
constexpr int G_WARP_FLOATS = G_IDX * 32;                                // per softmax warp: [idx][lane], conflict-free
constexpr int G_BYTES = 8 * G_WARP_FLOATS * 4;
constexpr int SMEM_BYTES = 2 * STAGE_BYTES + TAB_BYTES + G_BYTES + 256 /*barriers*/ + 1024 /*align*/;
constexpr int TM_SLOT = 208, TM_O = 128, TM_G = 416;
constexpr float kLog2e = 1.4426950408889634f;
}  // namespace win

struct WinParams {
  __nv_bfloat16* out;
  const int* out_map;
  int Bp, nh, num_items, has_bias;
  // Row tiles: A = rows [0, qa_rows), B = rows [qb_row0, 196). Plain / row-mapped output: 128 | 68 rows. Un-partitioned output
  // (H > 0: out is the [B, H, W, nh * 64] token map, windows of 14 x 14 tokens, nwh x nww per image): 126 | 70 rows = 9 | 5
  // whole window rows, so that each tile leaves through ONE 4-D TMA store, clipped at the image edge by the copy engine.
  int qa_rows, qb_row0;
  int H, W, nwh, nww;
  float scale_log2;
  long long* trace;     // perf debug (MMSAM_ATT_TRACE): clock64 stamps of CTA 0, [role 0..2][item < 16][event < 8]
};
#define WIN_TRACE(role, it, ev)                                                                              \
  do {                                                                                                       \
    if (p.trace && blockIdx.x == 0 && (it) < 16 && lane == 0) p.trace[((role) * 16 + (it)) * 8 + (ev)] = clock64(); \
  } while (0)

__device__ __forceinline__ float win_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void tmem_st_32x32b_x8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}

// One 28-column chunk of the score row = key rows kh = 2j, 2j + 1 (14 keys each): the (kh, kw) pattern is the same in
// every chunk, so the 7 chunks of a row run as a LOOP over j with the key-column biases bw2 in registers and the two
// key-row biases read from the warp's gather area. (A fully unrolled row — 7 chunks x 2 passes with compile-time key
// indices — was ~2000 instructions = 32 KB: with two softmax groups, the MMA and the TMA warp in different code regions
// the instruction cache thrashed and every phase of the tile ran 3-10x slower than its instruction count.)
// r: raw scores of columns [28j, 28j + 32) (the last 4 belong to the next chunk). EXP = false: running maximum of the
// logits u = s * scale_log2 + bw + bh (base 2); EXP = true: p = exp2(u - max), row sum, bf16 P written over the score
// columns already consumed: P column c = keys 2c, 2c + 1; every chunk also zeroes the two P columns after its own (the
// next chunk overwrites them; after the last chunk they are keys 196..199, which must not contribute).
template <bool EXP>
__device__ __forceinline__ void win_chunk(const uint32_t (&r)[32], int j, float bh_a, float bh_b, const u64 (&bw2)[7], u64 sc2,
                                          float& mx, u64& sum2, uint32_t s_addr) {
  const u64 ba = pack2(bh_a, bh_a), bb = pack2(bh_b, bh_b);
  u64 u[14];
#pragma unroll
  for (int i = 0; i < 7; ++i) {
    u[i] = fma2(pack2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1])), sc2, add2(bw2[i], ba));
    u[7 + i] = fma2(pack2(__uint_as_float(r[14 + 2 * i]), __uint_as_float(r[15 + 2 * i])), sc2, add2(bw2[i], bb));
  }
  if constexpr (!EXP) {
#pragma unroll
    for (int i = 0; i < 14; ++i) {
      float a, b;
      unpack2(u[i], a, b);
      mx = max3(mx, a, b);
    }
  } else {
    uint32_t pk[16];
#pragma unroll
    for (int i = 0; i < 14; ++i) {
      float a, b;
      unpack2(u[i], a, b);
      const float p0 = win_ex2(a), p1 = win_ex2(b);
      sum2 = add2(sum2, pack2(p0, p1));
      pk[i] = pack_bf16(p0, p1);
    }
    pk[14] = pk[15] = 0u;
    tmem_st_32x32b_x16(s_addr + 14 * j, pk);
  }
}

// One pass over the 196 score columns of the row, TMEM loads software-pipelined (the load of chunk j + 1 is in flight while
// chunk j is processed: a softmax warp has at most one other warp on its scheduler to hide latency). gh = the gather
// column holding G_h[13 + qh] of this row: bh[kh] = gh[-32 kh] (see the kernel). neg_m = -max for the exp pass.
template <bool EXP>
__device__ __forceinline__ void win_row_pass(uint32_t s_addr, const float* gh, bool has_bias, float neg_m, const u64 (&bw2)[7], u64 sc2,
                                             float& mx, u64& sum2, uint64_t* p_half, int lane) {
  uint32_t ra[32], rb[32];
  auto bias_of = [&](int kh) { return has_bias ? fmaf(gh[-32 * kh], win::kLog2e, neg_m) : neg_m; };
  tmem_ld_32x32b_x32(s_addr, ra);
#pragma unroll 1
  for (int jj = 0; jj < 3; ++jj) {
    const int j = 2 * jj;
    const float b0 = bias_of(2 * j), b1 = bias_of(2 * j + 1), b2 = bias_of(2 * j + 2), b3 = bias_of(2 * j + 3);
    tmem_ld_wait();
    tmem_ld_32x32b_x32(s_addr + 28 * (j + 1), rb);
    win_chunk<EXP>(ra, j, b0, b1, bw2, sc2, mx, sum2, s_addr);
    tmem_ld_wait();
    tmem_ld_32x32b_x32(s_addr + 28 * (j + 2), ra);
    win_chunk<EXP>(rb, j + 1, b2, b3, bw2, sc2, mx, sum2, s_addr);
    if (EXP && jj == 2) {
      // keys 0..167 (P columns 0..83) are complete: the tensor core starts O = P V on the first 128 keys while the last
      // chunk is still being exponentiated (the whole P.V latency was exposed when it was issued after the row)
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_half);
    }
  }
  const float b0 = bias_of(12), b1 = bias_of(13);
  tmem_ld_wait();
  win_chunk<EXP>(ra, 6, b0, b1, bw2, sc2, mx, sum2, s_addr);
  if constexpr (EXP) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %1, %1, %1};" ::"r"(s_addr + 100), "r"(0u) : "memory");   // keys 200..207
  }
}

__global__ void __launch_bounds__(320, 1)
attention_win_kernel(const __grid_constant__ CUtensorMap tmQA, const __grid_constant__ CUtensorMap tmQB,
                     const __grid_constant__ CUtensorMap tmKV, const __grid_constant__ CUtensorMap tmTabH,
                     const __grid_constant__ CUtensorMap tmTabW, const __grid_constant__ CUtensorMap tmOA,
                     const __grid_constant__ CUtensorMap tmOB, const WinParams p) {
  using namespace win;
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sStage = smem;
  uint8_t* sTab = sStage + 2 * STAGE_BYTES;
  float* sG = reinterpret_cast<float*>(sTab + TAB_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sG) + G_BYTES);
  uint64_t* item_full = bars + 0;    // [2]
  uint64_t* item_empty = bars + 2;   // [2]
  uint64_t* s_full = bars + 4;       // [2] S (and G) of the group's tile are in TMEM
  uint64_t* p_full = bars + 6;       // [2] P written
  uint64_t* o_full = bars + 8;       // [2] O complete
  uint64_t* slot_free = bars + 10;   // [2] O read: the slot may be overwritten
  uint64_t* g_empty = bars + 12;     // G consumed
  uint64_t* tab_full = bars + 13;
  uint64_t* p_half = bars + 14;      // [2] first 128 keys of P written
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_my = ((int)blockIdx.x < p.num_items) ? (p.num_items - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmQA); tma_prefetch_desc(&tmQB); tma_prefetch_desc(&tmKV);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&item_full[i], 1); mbar_init(&item_empty[i], 1);
      mbar_init(&s_full[i], 1); mbar_init(&p_full[i], 4); mbar_init(&o_full[i], 1); mbar_init(&slot_free[i], 4);
      mbar_init(&p_half[i], 4);
    }
    mbar_init(g_empty, 4);
    mbar_init(tab_full, 1);
    fence_barrier_init();
  }
  if (warp == 9) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_wait();

  if (warp == 8) {
    // =========================== TMA producer: the whole warp runs the loop, one elected lane issues ===========================
    {
      if (p.has_bias) {
        if (elect_one()) {
          mbar_arrive_expect_tx(tab_full, TAB_BYTES);
          tma_load_2d(sTab, &tmTabH, tab_full, 0, 0);
          tma_load_2d(sTab + 32 * 128, &tmTabW, tab_full, 0, 0);
        }
        __syncwarp();
      }
      for (int it = 0; it < n_my; ++it) {
        const int item = blockIdx.x + it * gridDim.x;
        const int head = item % p.nh, bp = item / p.nh;
        const int st = it & 1;
        mbar_wait(&item_empty[st], ((it >> 1) & 1) ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&item_full[st], STAGE_BYTES);
          uint8_t* s = sStage + st * STAGE_BYTES;
          tma_load_4d(s, &tmQA, &item_full[st], 0, head, 0, bp);
          tma_load_4d(s + QA_BYTES, &tmQB, &item_full[st], 0, head, p.qb_row0, bp);
          tma_load_4d(s + QA_BYTES + QB_BYTES, &tmKV, &item_full[st], 0, p.nh + head, 0, bp);
          tma_load_4d(s + QA_BYTES + QB_BYTES + KV_BYTES, &tmKV, &item_full[st], 0, 2 * p.nh + head, 0, bp);
        }
        __syncwarp();
      }
    }
  } else if (warp == 9) {
    // =========================== MMA issuer: the whole warp runs the loop, one elected lane issues (see elect_one) ===========================
    {
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, NK, 0, 0);
      constexpr uint32_t idesc_g = umma_idesc_bf16(128, 64, 0, 0);
      constexpr uint32_t idesc_pv = umma_idesc_bf16(128, D, 0, 1);      // A = P in TMEM (K-major), B = V MN-major
      if (p.has_bias) mbar_wait(tab_full, 0);
      const uint32_t t_addr = smem_u32(sTab);
      uint32_t gn = 0;                                                   // G uses so far (g_empty phase bookkeeping)
      auto issue_s = [&](int g, uint32_t q_addr, uint32_t k_addr, int it) {
        WIN_TRACE(2, it, g * 4 + 0);
        mbar_wait(&slot_free[g], (it & 1) ^ 1);
        if (p.has_bias) mbar_wait(g_empty, (gn & 1) ^ 1);
        ++gn;
        tc_fence_after();
        WIN_TRACE(2, it, g * 4 + 1);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16_ss(tmem + g * TM_SLOT, umma_desc_sw128(q_addr + k * 32), umma_desc_sw128(k_addr + k * 32), idesc_s, k != 0 ? 1u : 0u);
          if (p.has_bias) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16_ss(tmem + TM_G, umma_desc_sw128(q_addr + k * 32), umma_desc_sw128(t_addr + k * 32), idesc_g, k != 0 ? 1u : 0u);
          }
          umma_commit(&s_full[g]);
        }
        __syncwarp();
      };
      auto issue_pv = [&](int g, uint32_t v_addr, int it) {
        WIN_TRACE(2, it, g * 4 + 2);
        mbar_wait(&p_half[g], it & 1);          // keys 0..159 (P columns 0..83 are written)
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < 10; ++kk)
            umma_f16_ts(tmem + g * TM_SLOT + TM_O, tmem + g * TM_SLOT + kk * 8, umma_desc_sw128(v_addr + kk * 16 * 128), idesc_pv,
                        kk != 0 ? 1u : 0u);
        }
        __syncwarp();
        mbar_wait(&p_full[g], it & 1);          // keys 160..207
        tc_fence_after();
        WIN_TRACE(2, it, g * 4 + 3);
        if (elect_one()) {
#pragma unroll
          for (int kk = 10; kk < NK / 16; ++kk)
            umma_f16_ts(tmem + g * TM_SLOT + TM_O, tmem + g * TM_SLOT + kk * 8, umma_desc_sw128(v_addr + kk * 16 * 128), idesc_pv, 1u);
          umma_commit(&o_full[g]);
        }
        __syncwarp();
      };
      // Issue order: the two groups run HALF A PERIOD apart. While group 0's exponentials of tile A(it) run, the tensor core
      // finishes tile B(it - 1) (P.V) and starts tile B(it) (S, G); while group 1's exponentials of B(it) run, it finishes
      // A(it) and starts A(it + 1). Issuing S_A, S_B together would put both groups into the MUFU-bound phase at the same
      // time and leave the MMA / output phases of an item exposed (measured: 151 us per launch).
      auto stage_of = [&](int it) { return smem_u32(sStage + (it & 1) * STAGE_BYTES); };
      if (n_my > 0) {
        mbar_wait(&item_full[0], 0);
        tc_fence_after();
        issue_s(0, stage_of(0), stage_of(0) + QA_BYTES + QB_BYTES, 0);
      }
      for (int it = 0; it < n_my; ++it) {
        const uint32_t base = stage_of(it);
        const uint32_t k_addr = base + QA_BYTES + QB_BYTES, v_addr = k_addr + KV_BYTES;
        if (it > 0) {
          issue_pv(1, stage_of(it - 1) + QA_BYTES + QB_BYTES + KV_BYTES, it - 1);
          if (elect_one()) umma_commit(&item_empty[(it - 1) & 1]);   // every MMA reading that stage has been issued
          __syncwarp();
        }
        issue_s(1, base + QA_BYTES, k_addr, it);
        issue_pv(0, v_addr, it);
        if (it + 1 < n_my) {
          mbar_wait(&item_full[(it + 1) & 1], ((it + 1) >> 1) & 1);
          tc_fence_after();
          issue_s(0, stage_of(it + 1), stage_of(it + 1) + QA_BYTES + QB_BYTES, it + 1);
        }
      }
      if (n_my > 0) {
        issue_pv(1, stage_of(n_my - 1) + QA_BYTES + QB_BYTES + KV_BYTES, n_my - 1);
        if (elect_one()) umma_commit(&item_empty[(n_my - 1) & 1]);
        __syncwarp();
      }
    }
  } else {
    // =========================== softmax groups (warps 0-3: row tile A, warps 4-7: row tile B) ===========================
    const int g = warp >> 2, wq = warp & 3;
    const int row = wq * 32 + lane;                       // TMEM lane == query row inside the tile
    const int q = g ? p.qb_row0 + row : row;
    const bool warp_valid = g == 0 || p.qb_row0 + wq * 32 < T;   // tile B: warp 2 has 4 (6) rows, warp 3 none
    const bool row_valid = q < T && (g == 1 || row < p.qa_rows);
    const uint32_t grp_stg = smem_u32(sG + g * 4 * G_WARP_FLOATS);   // un-partitioned output: the group's rows, 128 B each
    const int qc = row_valid ? q : T - 1;
    const int qh = qc / KS, qw = qc - qh * KS;
    const uint32_t lane_addr = tmem + ((uint32_t)(wq * 32) << 16);
    const uint32_t s_addr = lane_addr + g * TM_SLOT;
    float* gw = sG + (g * 4 + wq) * G_WARP_FLOATS + lane;      // this warp's [idx][lane] gather area, this lane's column
    const u64 sc2 = pack2(p.scale_log2, p.scale_log2);
    bool store_pending = false;
    for (int it = 0; it < n_my; ++it) {
      const int item = blockIdx.x + it * gridDim.x;
      const int head = item % p.nh, bp = item / p.nh;
      // window coordinates of the un-partition store, off the critical path (the divisions cost the issuing warp ~300
      // cycles when they sat in front of the TMA store)
      const int o_wx = p.H > 0 ? (bp % p.nww) * KS : 0, o_wy = p.H > 0 ? ((bp / p.nww) % p.nwh) * KS + (g ? 9 : 0) : 0;
      const int o_b = p.H > 0 ? bp / (p.nww * p.nwh) : 0;
      WIN_TRACE(g, it, 0);
      mbar_wait(&s_full[g], it & 1);
      tc_fence_after();
      WIN_TRACE(g, it, 1);
      u64 bw2[7];
      if (store_pending) {      // the previous tile's rows have left the staging area (= this tile's gather area)
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        store_pending = false;
      }
      if (p.H > 0) asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");   // ... for every warp of the group
      if (p.has_bias) {
        if (warp_valid) {
          uint32_t r[64];
          tmem_ld_32x32b_x32(lane_addr + TM_G, r);
          tmem_ld_32x32b_x32(lane_addr + TM_G + 32, r + 32);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 27; ++j) {      // table rows 0..26 (h) and 32..58 (w)
            gw[j * 32] = __uint_as_float(r[j]);
            gw[(28 + j) * 32] = __uint_as_float(r[32 + j]);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(g_empty);
        if (warp_valid) {
          // rel_h[q, kh] = G_h[qh - kh + 13], rel_w[q, kw] = G_w[qw - kw + 13]  (image_encoder.py:580-584, 609-621), in base 2;
          // the key-column values live in registers, the key-row values are read per chunk from gh[-32 kh]
#pragma unroll
          for (int j = 0; j < 7; ++j) bw2[j] = pack2(gw[(28 + 13 + qw - 2 * j) * 32] * kLog2e, gw[(28 + 12 + qw - 2 * j) * 32] * kLog2e);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 7; ++j) bw2[j] = 0ull;
      }
      float l = 1.f;
      WIN_TRACE(g, it, 2);
      if (warp_valid) {
        // ---- pass 1: row maximum; pass 2: exp2(u - max), sum, P -> TMEM ----
        const float* gh = gw + (13 + qh) * 32;
        u64 sum2 = 0ull;
        float mx = -INFINITY, unused = 0.f;
        win_row_pass<false>(s_addr, gh, p.has_bias != 0, 0.f, bw2, sc2, mx, sum2, nullptr, lane);
        WIN_TRACE(g, it, 3);
        win_row_pass<true>(s_addr, gh, p.has_bias != 0, -mx, bw2, sc2, unused, sum2, &p_half[g], lane);
        float l0, l1;
        unpack2(sum2, l0, l1);
        l = l0 + l1;
        tmem_st_wait();
      }
      if (!warp_valid && lane == 0) mbar_arrive(&p_half[g]);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[g]);
      WIN_TRACE(g, it, 4);
      // ---- O = P V / l ----
      mbar_wait(&o_full[g], it & 1);
      tc_fence_after();
      WIN_TRACE(g, it, 5);
      uint32_t o[64];
      if (warp_valid) {
        tmem_ld_32x32b_x32(s_addr + TM_O, o);
        tmem_ld_32x32b_x32(s_addr + TM_O + 32, o + 32);
        tmem_ld_wait();
        WIN_TRACE(g, it, 7);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&slot_free[g]);
      const float inv = 1.f / l;
      if (p.H > 0) {
        // ---- un-partitioned output: the group's 126 / 70 rows are whole window rows -> staged densely in token order and
        //      written by ONE 4-D TMA store (box 64 ch x 14 x 9 | 5 tokens), out-of-image tokens clipped by the copy engine.
        //      (32 per-row bulk copies per warp, each issued by its own lane, cost ~2600 cycles per tile.) ----
        WIN_TRACE(3 + g, it, 0);
        asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");      // every warp of the group is done with its gather area
        WIN_TRACE(3 + g, it, 1);
        if (warp_valid && row_valid) {
          const uint32_t stg = grp_stg + row * 128;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float f[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(o[8 * i + j]) * inv;
            // SWIZZLE_128B staging (the store's tensor map says so): a thread owns a 128-byte row, and with the chunks
            // in place the 8 rows of a quarter warp would all hit the same four banks (8-way conflict, ~930 cycles per tile)
            sts128(stg + ((i ^ (row & 7)) << 4), pack8(f));
          }
        }
        WIN_TRACE(3 + g, it, 2);
        fence_proxy_async();
        asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");
        WIN_TRACE(3 + g, it, 3);
        if (wq == 0) {
          if (elect_one()) {
            asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                             reinterpret_cast<uint64_t>(g ? &tmOB : &tmOA)),
                         "r"(grp_stg), "r"(head * D), "r"(o_wx), "r"(o_wy), "r"(o_b)
                         : "memory");
          }
          __syncwarp();
        }
        WIN_TRACE(3 + g, it, 4);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        store_pending = true;
        WIN_TRACE(3 + g, it, 5);
        WIN_TRACE(g, it, 6);
      } else if (warp_valid) {
        // Output rows are nh * 128 B apart in global memory. Plain stores (a thread's own row as 8 x 16 B, or transposed so
        // that 8 lanes cover a row) took ~3000 cycles per tile: with 227 of the SM's 228 KB carved out as shared memory the
        // LSU path tracks only a few outstanding lines and every store instruction waits for a slot. Each lane instead
        // stages its row (128 contiguous bytes; the warp's bias-gather area is free once bh / bw are in registers) and
        // hands it to the bulk-copy engine: one cp.async.bulk per row, asynchronous, no LSU slot held.
        WIN_TRACE(3 + g, it, 0);
        long long orow = (long long)bp * T + q;
        if (row_valid && p.out_map) orow = p.out_map[orow];
        if (!row_valid) orow = -1;
        WIN_TRACE(3 + g, it, 1);
        // 32 rows of the warp's 7 KB at a 144-byte pitch: with 128 the 8 rows of a quarter warp share four banks (8-way conflict)
        uint8_t* stg = reinterpret_cast<uint8_t*>(sG + (g * 4 + wq) * G_WARP_FLOATS) + lane * 144;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float f[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(o[8 * i + j]) * inv;
          *reinterpret_cast<uint4*>(stg + i * 16) = pack8(f);
        }
        WIN_TRACE(3 + g, it, 2);
        fence_proxy_async();
        WIN_TRACE(3 + g, it, 3);
        if (orow >= 0) {
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 128;" ::"l"(p.out + orow * (p.nh * D) + head * D),
                       "r"(smem_u32(stg))
                       : "memory");
        }
        WIN_TRACE(3 + g, it, 4);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        store_pending = true;
        WIN_TRACE(3 + g, it, 5);
        WIN_TRACE(g, it, 6);
      }
    }
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");      // this thread's output rows are written
  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

}  // namespace mmsam

extern long long* g_win_trace_buf;

// SAM-window specialisation of mmsam_attention_bf16 (same contract; T == 196, Kh == Kw == 14). Called by attention.cu.
static int attention_win_launch(const void* qkv, void* out, const int* out_row_map_dev, const void* tab_h, const void* tab_w, int Bp,
                                int nh, float scale, int max_ctas, cudaStream_t stream, int B, int H, int W);

int mmsam_attention_win(const void* qkv, void* out, const int* out_row_map_dev, const void* tab_h, const void* tab_w, int Bp,
                        int nh, float scale, int max_ctas, cudaStream_t stream) {
  return attention_win_launch(qkv, out, out_row_map_dev, tab_h, tab_w, Bp, nh, scale, max_ctas, stream, 0, 0, 0);
}

// See include/mmsam_b200.h for the contract.
MMSAM_API int mmsam_attention_window_bf16(const void* qkv, void* out, const void* tab_h, const void* tab_w, int B, int H, int W, int nh,
                                          float scale, int max_ctas, void* stream) {
  if (B < 0 || H <= 0 || W <= 0 || nh <= 0) return MMSAM_ERR_BAD_ARG;
  if (B == 0) return MMSAM_OK;
  if (!qkv || !out || (tab_h != nullptr) != (tab_w != nullptr)) return MMSAM_ERR_BAD_ARG;
  if ((((uintptr_t)qkv | (uintptr_t)out | (uintptr_t)tab_h | (uintptr_t)tab_w) & 15)) return MMSAM_ERR_BAD_ARG;
  const int nwh = (H + 13) / 14, nww = (W + 13) / 14;
  return attention_win_launch(qkv, out, nullptr, tab_h, tab_w, B * nwh * nww, nh, scale, max_ctas, (cudaStream_t)stream, B, H, W);
}

static int attention_win_launch(const void* qkv, void* out, const int* out_row_map_dev, const void* tab_h, const void* tab_w, int Bp,
                                int nh, float scale, int max_ctas, cudaStream_t stream, int B, int H, int W) {
  using namespace mmsam;
  using namespace mmsam::win;
  mmsam_host::EncodeTiledFn enc = mmsam_host::get_encode_tiled();
  if (!enc) return MMSAM_ERR_DRIVER;
  const bool has_bias = tab_h != nullptr;
  CUtensorMap tmQA, tmQB, tmKV, tmH, tmW;
  const uint64_t C3 = (uint64_t)3 * nh * D;
  cuuint64_t dims[4] = {(cuuint64_t)D, (cuuint64_t)(3 * nh), (cuuint64_t)T, (cuuint64_t)Bp};
  cuuint64_t strides[3] = {(cuuint64_t)D * 2, C3 * 2, (cuuint64_t)T * C3 * 2};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  const bool unpart = H > 0;
  const cuuint32_t rows[3] = {128, QB_ROWS, NK};
  CUtensorMap* maps[3] = {&tmQA, &tmQB, &tmKV};
  for (int i = 0; i < 3; ++i) {
    cuuint32_t box[4] = {D, 1, rows[i], 1};
    if (enc(maps[i], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(qkv), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return MMSAM_ERR_DRIVER;
  }
  if (has_bias) {
    int rc = mmsam_host::make_tmap_2d_bf16(&tmH, tab_h, 32, D, D, 32, D);
    if (rc) return rc;
    rc = mmsam_host::make_tmap_2d_bf16(&tmW, tab_w, 32, D, D, 32, D);
    if (rc) return rc;
  } else {
    tmH = tmQA; tmW = tmQA;
  }
  CUtensorMap tmOA = tmQA, tmOB = tmQA;
  if (unpart) {
    const uint64_t C = (uint64_t)nh * D;
    cuuint64_t odims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t ostr[3] = {C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t boxa[4] = {D, KS, 9, 1}, boxb[4] = {D, KS, 5, 1};
    if (enc(&tmOA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, out, odims, ostr, boxa, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS ||
        enc(&tmOB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, out, odims, ostr, boxb, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return MMSAM_ERR_DRIVER;
  }
  WinParams p;
  p.out = (__nv_bfloat16*)out; p.out_map = out_row_map_dev; p.Bp = Bp; p.nh = nh; p.num_items = Bp * nh; p.has_bias = has_bias ? 1 : 0;
  p.qa_rows = unpart ? 9 * KS : 128; p.qb_row0 = unpart ? 9 * KS : 128;
  p.H = H; p.W = W; p.nwh = unpart ? (H + KS - 1) / KS : 0; p.nww = unpart ? (W + KS - 1) / KS : 0;
  p.scale_log2 = scale * kLog2e;
  static long long* trace_buf = nullptr;
  static const int want_trace = getenv("MMSAM_ATT_TRACE") != nullptr;
  if (want_trace && !trace_buf) {
    cudaMalloc(&trace_buf, 5 * 16 * 8 * sizeof(long long));
    cudaMemset(trace_buf, 0, 5 * 16 * 8 * sizeof(long long));
    g_win_trace_buf = trace_buf;
  }
  p.trace = want_trace ? trace_buf : nullptr;
  MMSAM_SET_SMEM_ONCE(attention_win_kernel, SMEM_BYTES);
  if (max_ctas <= 0 || max_ctas > kNumSMs) max_ctas = kNumSMs;
  const int grid = p.num_items < max_ctas ? p.num_items : max_ctas;
  cudaError_t le = mmsam_host::launch_pdl(attention_win_kernel, dim3(grid), dim3(320), SMEM_BYTES, stream, tmQA, tmQB, tmKV, tmH, tmW, tmOA, tmOB, p);
  if (le != cudaSuccess) return (int)le;
  MMSAM_LAUNCH_CHECK();
  return MMSAM_OK;
}

// perf debug: copy the clock64 trace of the last traced launch to the host (3 x 16 x 8 values); returns 0 when tracing is off
extern "C" __attribute__((visibility("default"))) int mmsam_dbg_attention_win_trace(long long* host_out) {
  long long* buf = nullptr;
  // the buffer pointer lives in mmsam_attention_win's static; re-derive it through a tiny launch-free trick: keep a copy
  extern long long* g_win_trace_buf;
  buf = g_win_trace_buf;
  if (!buf) return 0;
  cudaDeviceSynchronize();
  cudaMemcpy(host_out, buf, 5 * 16 * 8 * sizeof(long long), cudaMemcpyDeviceToHost);
  return 1;
}
long long* g_win_trace_buf = nullptr;
