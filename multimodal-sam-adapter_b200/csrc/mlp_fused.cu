// ConvNeXt block tail in ONE kernel:   t += gamma * ( GELU( LN(y) W1^T + b1 ) W2^T + b2 )
// (twin_convnext.py:98-132: norm -> pwconv1 -> GELU -> pwconv2 -> gamma -> residual), y = the 7x7 depthwise conv output
// (bf16 [M, C]), t = the tower's fp32 residual stream [M, C], hidden width 4C. The 4C-wide intermediate never leaves the
// SM: per 256-row tile the two GEMMs are chained 128 hidden units at a time.
//
// Before: LayerNorm kernel + GEMM(GELU) + GEMM(residual) = 3 launches, the [M, 4C] bf16 intermediate written and re-read
// (54 x 100 MB per step at stage 2), and two short-K GEMMs (K = C <= 384) that ran at ~660 TFLOP/s because their
// epilogues (GELU; fp32 residual) had nothing to hide behind.
//
// Design (CTA pair = cluster of 2, tcgen05 cta_group::2, M = 256 rows per pair, 128 per CTA):
//   * the pair's A tile (128 rows x C per CTA, bf16, SWIZZLE_128B k-blocks) stays resident in shared memory for the tile;
//     LayerNorm is folded: the MMA runs on the raw y, the epilogue applies rstd * (acc - mean * colsum[n]) + bias[n]
//     (W1 carries the LN weight, bias the LN bias). mean / rstd of a row are computed by the epilogue warps from the
//     resident tile while the first MMAs run - no LayerNorm or row-statistics pass over HBM.
//   * per chunk j of 128 hidden units:  G1(j): Hacc = A . W1[128j..128j+127, :]^T     (N = 128, K = C, fp32 in TMEM)
//                                       epilogue-1: Hacc -> registers (the accumulator is handed back at once) -> LN fold,
//                                                    bias, exact GELU -> bf16 -> a swizzled K-major shared-memory tile H
//                                       G2(j): O += H(j) . W2[:, 128j..128j+127]^T      (O: 128 lanes x C fp32 in TMEM)
//     issue order G1(j+1), G2(j): the tensor core has the next chunk's first GEMM while the GELU of chunk j runs.
//     The chunk is 128 wide because an M = 256 MMA costs the same ~69 cycles at N = 64 and N = 128 (measured: its A-operand
//     read from shared memory bounds it), and because every chunk costs the issuing thread ~800 cycles of barrier polls.
//     C = 384: O (384) + one Hacc (128) fill the 512 TMEM columns -> single accumulator / single H tile; C <= 192: two of each.
//   * the drain adds b2, scales by gamma and ADDS the result into t with a TMA reduction (cp.reduce.async.bulk.tensor
//     .add, fp32): the residual is never fetched by the SM, every element of t receives exactly one add (deterministic).
//   * weights stream through two 2-CTA rings of 64-wide k-block units (each CTA stages half of the N rows of either
//     operand, the tensor core reads the other half from the peer): L2 -> SM traffic per tile = |W1| + |W2| per PAIR.
// TMEM: O [0, C), Hacc[a] [C + 128 a, C + 128 a + 128).
#include "common.cuh"
#include "cg2.cuh"
#include <cstdlib>

namespace mmsam {

template <int C> struct MlpCfg {
  static constexpr int HID = 4 * C;
  static constexpr int HN = 128;                      // hidden units per chunk
  static constexpr int NCH = HID / HN;
  static constexpr int KB1 = (C + 63) / 64;           // 64-wide k-blocks of GEMM1 (the last may be zero-padded by TMA)
  static constexpr int KS1 = C / 16;                  // k-steps of GEMM1
  static constexpr int NSPLIT = C > 256 ? 2 : 1;      // GEMM2 instruction N = C / NSPLIT (<= 256)
  static constexpr int N2 = C / NSPLIT;
  static constexpr int NACC = C > 256 ? 1 : 2;        // Hacc buffers (TMEM) = H tiles (shared memory)
  static constexpr int A_BYTES = KB1 * 128 * 128;
  static constexpr int W1_UNIT = (HN / 2) * 128;            // per CTA: 64 weight rows x one 64-wide k-block
  static constexpr int W2_UNIT = (C / 2) * 128;             // per CTA: C / 2 weight rows x 64 hidden columns
  static constexpr int NU1 = C > 256 ? KB1 : 2 * KB1;       // W1 ring: one / two chunks deep
  static constexpr int NU2 = C > 256 ? 2 : 4;               // W2 ring: one / two chunks deep
  static constexpr int H_BYTES = 2 * 128 * 128;             // 128 rows x 128 hidden units, two k-blocks
  static constexpr int SMEM_BYTES = A_BYTES + NU1 * W1_UNIT + NU2 * W2_UNIT + NACC * H_BYTES + 512 + 1024;
  static constexpr int TMEM_COLS = 512;
  static constexpr int THREADS = 352;                       // 8 epilogue warps + A/W1 producer + MMA + W2 producer
  static_assert(C % 32 == 0 && C <= 384, "C");
  static_assert(W1_UNIT % 1024 == 0 && (W2_UNIT / NSPLIT) % 1024 == 0, "swizzle atoms");
  static_assert(C + NACC * HN <= 512, "TMEM");
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory");
};

struct MlpParams {
  const float* colsum1;   // unused: the LayerNorm is applied to the resident A tile, not folded into the epilogue
  const float* bias1;     // [4C]  b1 + W1 . ln_bias
  const float* bias2;     // [C]
  const float* gamma;     // [C] or null
  int M;
  float eps;
  long long* trace;   // perf debug (MMSAM_MLP_TRACE): clock64 stamps of pair 0, [role 0..2][chunk < 64][event < 12]
  int dbg;   // perf debug (env MMSAM_MLP_DBG): 1 no drain reduce, 2 no epilogue-1 math, 4 no G1 MMAs, 8 no G2 MMAs
};

#define MLP_TRACE(role, gi, ev)                                                                                   \
  do {                                                                                                            \
    if (p.trace && (blockIdx.x >> 1) == 0 && (gi) < 64) p.trace[((role) * 64 + (gi)) * 12 + (ev)] = clock64();   \
  } while (0)

template <int C>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(MlpCfg<C>::THREADS, 1)
convnext_mlp_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW1,
                    const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmT, const MlpParams p) {
  using Cfg = MlpCfg<C>;
  constexpr int NU1 = Cfg::NU1, NU2 = Cfg::NU2, NCH = Cfg::NCH, NACC = Cfg::NACC, KB1 = Cfg::KB1;
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sW1 = sA + Cfg::A_BYTES;
  uint8_t* sW2 = sW1 + NU1 * Cfg::W1_UNIT;
  uint8_t* sH = sW2 + NU2 * Cfg::W2_UNIT;             // NACC tiles; warp w's 4 KB slice of tile 0 doubles as its drain slab
  uint64_t* bars = reinterpret_cast<uint64_t*>(sH + NACC * Cfg::H_BYTES);
  // A poll of an mbarrier costs the MMA-issuing thread ~250 cycles even when the phase is already complete, so everything
  // one GEMM waits for lands on ONE barrier: g1_rdy[a] = 16 epilogue warps (accumulator a is back in registers) + both
  // CTAs' producers + the TMA bytes of the chunk's W1 units; g2_rdy[a] = 16 epilogue warps (H tile a written) + producers
  // + the bytes of the chunk's W2 units. Two polls per 128-wide chunk.
  uint64_t* w1_empty = bars;                 // [NU1] per CTA (multicast commit): unit consumed
  uint64_t* w2_empty = w1_empty + NU1;       // [NU2] per CTA
  uint64_t* a_full = w2_empty + NU2;         // per CTA: this CTA's A tile has landed
  uint64_t* a_ready = a_full + 1;            // leader: 16 epilogue warps normalised their rows of A
  uint64_t* a_empty = a_full + 2;            // per CTA: every G1 of the tile has read A
  uint64_t* hacc_full = a_full + 3;          // [2] per CTA: G1 accumulator complete
  uint64_t* g1_rdy = a_full + 5;             // [2] leader
  uint64_t* g2_rdy = a_full + 7;             // [2] leader
  uint64_t* h_free = a_full + 9;             // [2] per CTA: G2 has read H
  uint64_t* o_full = a_full + 11;            // per CTA
  uint64_t* o_empty = a_full + 12;           // leader: 16 epilogue warps drained O
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_full + 13);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int num_tiles = (p.M + 255) / 256;
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmW1); tma_prefetch_desc(&tmW2); tma_prefetch_desc(&tmT);
    for (int i = 0; i < NU1; ++i) mbar_init(&w1_empty[i], 1);
    for (int i = 0; i < NU2; ++i) mbar_init(&w2_empty[i], 1);
    mbar_init(a_full, 1); mbar_init(a_ready, 16); mbar_init(a_empty, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&hacc_full[i], 1); mbar_init(&g1_rdy[i], 18);
      mbar_init(&g2_rdy[i], 18); mbar_init(&h_free[i], 1);
    }
    mbar_init(o_full, 1); mbar_init(o_empty, 16);
    fence_barrier_init();
  }
  cluster_sync_all();
  if (warp == 9) tmem_alloc_cg2(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr uint32_t TM_H = C;               // Hacc[a] at TM_H + 128 a
  pdl_wait();          // the prologue above overlapped the previous kernel's tail; global memory is touched only from here on

  if (warp == 8) {
    // ---------------- A tile + W1 ring producer (both CTAs): the whole warp runs the loop, one elected lane issues ----------------
    {
      int s1 = 0; uint32_t ph1 = 0; int ti = 0, g = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs, ++ti) {
        const int m0 = tile * 256 + (int)rank * 128;
        mbar_wait(a_empty, (ti & 1) ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(a_full, Cfg::A_BYTES);
#pragma unroll
          for (int kb = 0; kb < KB1; ++kb) tma_load_2d(sA + kb * 16384, &tmA, a_full, kb * 64, m0);
        }
        __syncwarp();
        for (int j = 0; j < NCH; ++j, ++g) {
          const uint32_t lead = mapa_shared(smem_u32(&g1_rdy[g % NACC]), 0);
#pragma unroll 1
          for (int kb = 0; kb < KB1; ++kb) {
            mbar_wait(&w1_empty[s1], ph1 ^ 1);
            if (elect_one()) {
              if (kb == 0) {      // after the wait: the G1 that used this stage last has been issued, i.e. its phase is over
                if (rank == 0) mbar_arrive_expect_tx(&g1_rdy[g % NACC], 2 * KB1 * Cfg::W1_UNIT);
                else mbar_arrive_cluster(lead);
              }
              tma_load_2d_cg2(smem_u32(sW1 + s1 * Cfg::W1_UNIT), &tmW1, lead, kb * 64, j * Cfg::HN + (int)rank * (Cfg::HN / 2));
            }
            __syncwarp();
            if (++s1 == NU1) { s1 = 0; ph1 ^= 1; }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 10) {
    // ---------------- W2 ring producer (both CTAs): the whole warp runs the loop, one elected lane issues ----------------
    {
      int s2 = 0, g = 0; uint32_t ph2 = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        for (int j = 0; j < NCH; ++j, ++g) {
          const uint32_t lead = mapa_shared(smem_u32(&g2_rdy[g % NACC]), 0);
#pragma unroll 1
          for (int kb = 0; kb < 2; ++kb) {
            mbar_wait(&w2_empty[s2], ph2 ^ 1);
            const uint32_t dst = smem_u32(sW2 + s2 * Cfg::W2_UNIT);
            if (elect_one()) {
              if (kb == 0) {
                if (rank == 0) mbar_arrive_expect_tx(&g2_rdy[g % NACC], 2 * 2 * Cfg::W2_UNIT);
                else mbar_arrive_cluster(lead);
              }
#pragma unroll
              for (int h = 0; h < Cfg::NSPLIT; ++h)
                tma_load_2d_cg2(dst + h * (Cfg::N2 / 2) * 128, &tmW2, lead, j * Cfg::HN + kb * 64, h * Cfg::N2 + (int)rank * (Cfg::N2 / 2));
            }
            __syncwarp();
            if (++s2 == NU2) { s2 = 0; ph2 ^= 1; }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 9) {
    // ---------------- MMA issuer (leader CTA only): the whole warp runs the loop, one elected lane issues (see elect_one) ----------------
    if (rank == 0) {
      constexpr uint32_t idesc1 = umma_idesc_bf16(256, Cfg::HN, 0, 0);
      constexpr uint32_t idesc2 = umma_idesc_bf16(256, Cfg::N2, 0, 0);
      int s1 = 0, s2 = 0; uint32_t ph1 = 0, ph2 = 0;
      const uint32_t a_addr = smem_u32(sA);
      auto issue_g2 = [&](int gp) {
        const int jp = gp % NCH, tp = gp / NCH, hb = gp % NACC;
        if (lane == 0) MLP_TRACE(0, gp, 3);
        mbar_wait(&g2_rdy[hb], (gp / NACC) & 1);
        if (jp == 0) mbar_wait(o_empty, (tp & 1) ^ 1);
        tc_fence_after();
        if (lane == 0) MLP_TRACE(0, gp, 4);
        const uint32_t h_addr = smem_u32(sH + hb * Cfg::H_BYTES);
        if (lane == 0) MLP_TRACE(0, gp, 5);
        const bool leader = elect_one();
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) {
          const uint32_t w_addr = smem_u32(sW2 + s2 * Cfg::W2_UNIT);
          if (leader) {
#pragma unroll
            for (int k = 0; k < 4 && !(p.dbg & 8); ++k) {
#pragma unroll
              for (int h = 0; h < Cfg::NSPLIT; ++h)
                umma_f16_ss_cg2(tmem_base + h * Cfg::N2, umma_desc_sw128(h_addr + kb * 16384 + k * 32),
                                umma_desc_sw128(w_addr + h * (Cfg::N2 / 2) * 128 + k * 32), idesc2, (jp | kb | k) != 0 ? 1u : 0u);
            }
            umma_commit_cg2(&w2_empty[s2]);
          }
          if (++s2 == NU2) { s2 = 0; ph2 ^= 1; }
        }
        if (leader) {
          umma_commit_cg2(&h_free[hb]);
          if (jp == NCH - 1) umma_commit_cg2(o_full);
        }
        __syncwarp();
        if (lane == 0) MLP_TRACE(0, gp, 6);
      };
      int g = 0, ti = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs, ++ti) {
        // the previous tile's last G2 first: its o_full lets the epilogue warps drain and move on to this tile's A
        // (they announce a_ready), so it must not wait behind this tile's first G1
        if (g > 0) issue_g2(g - 1);
        if (lane == 0) MLP_TRACE(0, g, 7);
        mbar_wait(a_ready, ti & 1);
        if (lane == 0) MLP_TRACE(0, g, 8);
        for (int j = 0; j < NCH; ++j, ++g) {
          const int ab = g % NACC;
          if (lane == 0) MLP_TRACE(0, g, 0);
          mbar_wait(&g1_rdy[ab], (g / NACC) & 1);
          tc_fence_after();
          if (lane == 0) MLP_TRACE(0, g, 1);
          const uint32_t d_tmem = tmem_base + TM_H + ab * Cfg::HN;
          const bool leader = elect_one();
#pragma unroll
          for (int kb = 0; kb < KB1; ++kb) {
            const uint32_t w_addr = smem_u32(sW1 + s1 * Cfg::W1_UNIT);
            if (leader) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                if (kb * 4 + k < Cfg::KS1 && !(p.dbg & 4))
                  umma_f16_ss_cg2(d_tmem, umma_desc_sw128(a_addr + kb * 16384 + k * 32), umma_desc_sw128(w_addr + k * 32), idesc1,
                                  (kb | k) != 0 ? 1u : 0u);
              }
              umma_commit_cg2(&w1_empty[s1]);
            }
            if (++s1 == NU1) { s1 = 0; ph1 ^= 1; }
          }
          if (leader) {
            umma_commit_cg2(&hacc_full[ab]);
            if (j == NCH - 1) umma_commit_cg2(a_empty);
          }
          __syncwarp();
          if (lane == 0) MLP_TRACE(0, g, 2);
          if (j > 0) issue_g2(g - 1);
        }
      }
      if (g > 0) {
        issue_g2(g - 1);
        // the peer's last remote arrivals must have landed before this CTA's barriers can go away
        mbar_wait(o_empty, (ti - 1) & 1);
      }
    }
    __syncwarp();
  } else {
    // ---------------- epilogue warps 0..7 (both CTAs): LN statistics, GELU, drain ----------------
    const int quad = warp & 3, half = warp >> 2;
    const int row_l = quad * 32 + lane;                               // row inside this CTA's 128
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16);
    // this warp's rows of hidden k-block `half` of an H tile == 4 KB that only this warp writes: its drain staging buffers
    const uint32_t acc_free_lead[2] = {mapa_shared(smem_u32(&g1_rdy[0]), 0), mapa_shared(smem_u32(&g1_rdy[1]), 0)};
    const uint32_t h_ready_lead[2] = {mapa_shared(smem_u32(&g2_rdy[0]), 0), mapa_shared(smem_u32(&g2_rdy[1]), 0)};
    const uint32_t o_empty_lead = mapa_shared(smem_u32(o_empty), 0);
    const uint32_t a_ready_lead = mapa_shared(smem_u32(a_ready), 0);
    constexpr int NP = C / 32;                                        // 32-column fp32 panels of O
    constexpr int NP0 = (NP + 1) / 2;
    const int pan_lo = half == 0 ? 0 : NP0, pan_hi = half == 0 ? NP0 : NP;
    if (lane == 0) {                                                  // every accumulator starts out free
#pragma unroll
      for (int a = 0; a < NACC; ++a) mbar_arrive_cluster(acc_free_lead[a]);
    }
    // LayerNorm of the resident A tile, in place: (y - mean) * rstd rounded to bf16 - what the LayerNorm kernel would have
    // written - so the chunk epilogues have no per-element LayerNorm arithmetic. Mapping for this step only: warp w owns
    // rows [16 w, 16 w + 16), two lanes per row (lanes l and l + 16 take alternate halves of the k-blocks and meet through
    // one shuffle); a quarter-warp reads 8 different 16-byte columns of the swizzled tile (conflict-free).
    auto normalise_a = [&](int tile_idx) {
      mbar_wait(a_full, tile_idx & 1);
      constexpr int KBH = (KB1 + 1) / 2;
      const int nrow = warp * 16 + (lane & 15), kh = lane >> 4;
      const int kb_lo = kh == 0 ? 0 : KBH, kb_hi = kh == 0 ? KBH : KB1;
      const uint32_t a_row = smem_u32(sA) + nrow * 128;
      // this lane's part of the row stays in registers between the passes where it fits (C <= 192); C = 384 re-reads it:
      // beside 225 KB of shared memory there is no L1 left, a spilled register costs an L2 round trip
      constexpr bool KEEP = KBH <= 2;
      uint4 raw[KEEP ? KBH * 8 : 1];
      u64 s2 = pack2(0.f, 0.f), ss2 = pack2(0.f, 0.f);
#pragma unroll
      for (int i = 0; i < KBH; ++i) {
        if (kb_lo + i < kb_hi) {
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const uint4 v = lds128(a_row + (kb_lo + i) * 16384 + ((c ^ (nrow & 7)) << 4));
            if (KEEP) raw[i * 8 + c] = v;
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const u64 pr = pack2(bf16lo(w[k]), bf16hi(w[k]));
              s2 = add2(s2, pr);
              ss2 = fma2(pr, pr, ss2);
            }
          }
        }
      }
      float s, s_hi, ss, ss_hi;
      unpack2(s2, s, s_hi);
      unpack2(ss2, ss, ss_hi);
      s += s_hi; ss += ss_hi;
      s += __shfl_xor_sync(0xffffffffu, s, 16);
      ss += __shfl_xor_sync(0xffffffffu, ss, 16);
      const float mean = s * (1.f / C);
      const float rstd = rsqrtf(fmaxf(ss * (1.f / C) - mean * mean, 0.f) + p.eps);
      const u64 rs2 = pack2(rstd, rstd), nm2 = pack2(-mean * rstd, -mean * rstd);
#pragma unroll
      for (int i = 0; i < KBH; ++i) {
        if (kb_lo + i < kb_hi) {
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const uint4 v = KEEP ? raw[i * 8 + c] : lds128(a_row + (kb_lo + i) * 16384 + ((c ^ (nrow & 7)) << 4));
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
            uint32_t o[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              float lo, hi;
              unpack2(fma2(pack2(bf16lo(w[k]), bf16hi(w[k])), rs2, nm2), lo, hi);
              o[k] = pack_bf16(lo, hi);
            }
            sts128(a_row + (kb_lo + i) * 16384 + ((c ^ (nrow & 7)) << 4), make_uint4(o[0], o[1], o[2], o[3]));
          }
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(a_ready_lead);
    };
    constexpr int NHS = 2 * NACC;                                     // drain half-slabs per warp
    int g = 0, ti = 0, hs = 0;
    if (pair < num_tiles) normalise_a(0);
    for (int tile = pair; tile < num_tiles; tile += num_pairs, ++ti) {
      if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the slab's last reduce has read it
      __syncwarp();
      // ---- per hidden chunk: Hacc -> registers -> LN fold + bias + GELU -> bf16 -> the H tile in shared memory ----
      for (int j = 0; j < NCH; ++j, ++g) {
        const int ab = g % NACC;
        const int n0 = j * Cfg::HN + half * 64;
        const float4* bip = reinterpret_cast<const float4*>(p.bias1 + n0);
        float4 bi[8];                                    // bias of the first 32 columns, in flight across the wait
#pragma unroll
        for (int i = 0; i < 8; ++i) bi[i] = __ldg(bip + i);
        if (warp == 0 && lane == 0) MLP_TRACE(1 + rank, g, 0);
        mbar_wait(&hacc_full[ab], (g / NACC) & 1);
        tc_fence_after();
        if (warp == 0 && lane == 0) MLP_TRACE(1 + rank, g, 1);
        const uint32_t t_addr = lane_addr + TM_H + ab * Cfg::HN + half * 64;
        uint32_t r[64];
        tmem_ld_32x32b_x32(t_addr, r);
        tmem_ld_32x32b_x32(t_addr + 32, r + 32);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(acc_free_lead[ab]);     // the tensor core may start the G1 that reuses this accumulator
        if (warp == 0 && lane == 0) MLP_TRACE(1 + rank, g, 2);
        uint32_t pk[32];
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          float4 bn[8];
          if (hf == 0) {                                 // the second half's bias: issued before the first half's arithmetic
#pragma unroll
            for (int i = 0; i < 8; ++i) bn[i] = __ldg(bip + 8 + i);
          }
#pragma unroll
          for (int i = 0; i < 8 && !(p.dbg & 2); ++i) {
            const int q = hf * 8 + i;
            float v0 = __uint_as_float(r[4 * q]) + bi[i].x, v1 = __uint_as_float(r[4 * q + 1]) + bi[i].y;
            float v2 = __uint_as_float(r[4 * q + 2]) + bi[i].z, v3 = __uint_as_float(r[4 * q + 3]) + bi[i].w;
            gelu_erf2(v0, v1);
            gelu_erf2(v2, v3);
            pk[2 * q] = pack_bf16(v0, v1);
            pk[2 * q + 1] = pack_bf16(v2, v3);
          }
          if (hf == 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i) bi[i] = bn[i];
          }
        }
        if (warp == 0 && lane == 0) MLP_TRACE(1 + rank, g, 3);
        mbar_wait(&h_free[ab], ((g / NACC) & 1) ^ 1);                // G2 of the chunk that used this H tile has read it
        const uint32_t h_row = smem_u32(sH) + ab * Cfg::H_BYTES + half * 16384;
#pragma unroll
        for (int c = 0; c < 8; ++c)
          sts128(h_row + swz128(row_l, c), make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]));
        fence_proxy_async();
        if (warp == 0 && lane == 0) MLP_TRACE(1 + rank, g, 4);
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(h_ready_lead[ab]);
        if (warp == 0 && lane == 0) MLP_TRACE(1 + rank, g, 5);
      }

      // ---- the next tile's A (it landed while the last chunks ran): normalised before this tile's drain so that the tensor
      //      core has the next G1 to run while the epilogue warps drain ----
      if (tile + num_pairs < num_tiles) {
        if (warp == 0 && lane == 0) MLP_TRACE(1 + rank, g, 6);
        normalise_a(ti + 1);
        if (warp == 0 && lane == 0) MLP_TRACE(1 + rank, g, 8);
      }
      // ---- drain: O -> (+ b2) * gamma -> 16-column half-slabs (32 rows x 64 B, SWIZZLE_64B) -> TMA reduce-add into t;
      //      NHS half-slabs per warp rotate so that the copy engine reads one while the next is being written ----
      float4 b2v[4], gmv[4];
      auto load_bg = [&](int col0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          b2v[i] = __ldg(reinterpret_cast<const float4*>(p.bias2 + col0) + i);
          gmv[i] = p.gamma ? __ldg(reinterpret_cast<const float4*>(p.gamma + col0) + i) : make_float4(1.f, 1.f, 1.f, 1.f);
        }
      };
      if (pan_lo < pan_hi) load_bg(pan_lo * 32);
      if (warp == 0 && lane == 0) MLP_TRACE(1 + rank, g - 1, 9);
      mbar_wait(o_full, ti & 1);
      tc_fence_after();
      if (warp == 0 && lane == 0) MLP_TRACE(1 + rank, g - 1, 10);
      const int row0 = tile * 256 + (int)rank * 128 + quad * 32;
      for (int hp = 2 * pan_lo; hp < 2 * pan_hi; ++hp) {
        const int col0 = hp * 16;
        uint32_t r[16];
        tmem_ld_32x32b_x16(lane_addr + col0, r);
        tmem_ld_wait();
        if (hp == 2 * pan_hi - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(o_empty_lead);
        }
        uint4 out[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          out[i] = make_uint4(__float_as_uint((__uint_as_float(r[4 * i]) + b2v[i].x) * gmv[i].x),
                              __float_as_uint((__uint_as_float(r[4 * i + 1]) + b2v[i].y) * gmv[i].y),
                              __float_as_uint((__uint_as_float(r[4 * i + 2]) + b2v[i].z) * gmv[i].z),
                              __float_as_uint((__uint_as_float(r[4 * i + 3]) + b2v[i].w) * gmv[i].w));
        }
        if (hp + 1 < 2 * pan_hi) load_bg(col0 + 16);          // in flight across the slab hand-over below
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(NHS - 1) : "memory");
        __syncwarp();
        const uint32_t hslab = smem_u32(sH) + (uint32_t)(hs >> 1) * Cfg::H_BYTES + (uint32_t)warp * 4096 + (uint32_t)(hs & 1) * 2048;
#pragma unroll
        for (int i = 0; i < 4; ++i) sts128(hslab + lane * 64 + ((i ^ ((lane >> 1) & 3)) << 4), out[i]);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0 && !(p.dbg & 1)) {
          asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                           reinterpret_cast<uint64_t>(&tmT)),
                       "r"(hslab), "r"(col0), "r"(row0)
                       : "memory");
        }
        if (lane == 0) asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        hs = hs + 1 == NHS ? 0 : hs + 1;
      }
      if (warp == 0 && lane == 0) MLP_TRACE(1 + rank, g - 1, 11);
      if (pan_lo == pan_hi) {                 // (C == 32 only) no panel: still hand O back
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(o_empty_lead);
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    __syncwarp();
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc_cg2(tmem_base, Cfg::TMEM_COLS);
  }
}

template <int C>
static int launch_mlp(const CUtensorMap& tmA, const CUtensorMap& tmW1, const CUtensorMap& tmW2, const CUtensorMap& tmT,
                      const MlpParams& p, int max_ctas, cudaStream_t st) {
  using Cfg = MlpCfg<C>;
  MMSAM_SET_SMEM_ONCE((convnext_mlp_kernel<C>), Cfg::SMEM_BYTES);
  const int num_tiles = (p.M + 255) / 256;
  int pairs = max_ctas / 2;
  if (num_tiles < pairs) pairs = num_tiles;
  cudaError_t le = mmsam_host::launch_pdl(convnext_mlp_kernel<C>, dim3(2 * pairs), dim3(Cfg::THREADS), Cfg::SMEM_BYTES, st, tmA, tmW1, tmW2, tmT, p);
  if (le != cudaSuccess) return (int)le;
  MMSAM_LAUNCH_CHECK();
  return MMSAM_OK;
}

}  // namespace mmsam

static long long* g_mlp_trace_buf = nullptr;
// perf debug: copy the clock64 trace of the last traced launch to the host (3 x 64 x 12 values); returns 0 when tracing is off
extern "C" __attribute__((visibility("default"))) int mmsam_dbg_mlp_trace(long long* host_out) {
  if (!g_mlp_trace_buf) return 0;
  cudaDeviceSynchronize();
  cudaMemcpy(host_out, g_mlp_trace_buf, 3 * 64 * 12 * sizeof(long long), cudaMemcpyDeviceToHost);
  return 1;
}

// See include/mmsam_b200.h for the contract.
MMSAM_API int mmsam_convnext_mlp_bf16(const void* y, long long ldy, const void* W1, const float* colsum1, const float* bias1,
                                      const void* W2, const float* bias2, const float* gamma, float* t, long long ldt, int M,
                                      int C, float eps, int max_ctas, void* stream) {
  using namespace mmsam;
  if (M < 0) return MMSAM_ERR_BAD_ARG;
  if (M == 0) return MMSAM_OK;
  if (!y || !W1 || !bias1 || !W2 || !bias2 || !t) return MMSAM_ERR_BAD_ARG;      // colsum1: unused (see the header), may be null
  if (C != 96 && C != 192 && C != 384) return MMSAM_ERR_UNSUPPORTED;
  if ((ldy & 7) || ldy < C || (ldt & 3) || ldt < C) return MMSAM_ERR_BAD_ARG;
  if ((((uintptr_t)y | (uintptr_t)W1 | (uintptr_t)W2 | (uintptr_t)t | (uintptr_t)colsum1 | (uintptr_t)bias1 | (uintptr_t)bias2 |
        (uintptr_t)gamma) & 15))
    return MMSAM_ERR_BAD_ARG;
  if (max_ctas <= 1 || max_ctas > kNumSMs) max_ctas = kNumSMs;
  const int HID = 4 * C;
  const int n2 = C > 256 ? C / 2 : C;
  CUtensorMap tmA, tmW1, tmW2, tmT;
  int rc = mmsam_host::make_tmap_2d_bf16(&tmA, y, (uint64_t)M, (uint64_t)C, (uint64_t)ldy, 128, 64);
  if (rc) return rc;
  rc = mmsam_host::make_tmap_2d_bf16(&tmW1, W1, (uint64_t)HID, (uint64_t)C, (uint64_t)C, 64, 64);
  if (rc) return rc;
  rc = mmsam_host::make_tmap_2d_bf16(&tmW2, W2, (uint64_t)C, (uint64_t)HID, (uint64_t)HID, (uint32_t)(n2 / 2), 64);
  if (rc) return rc;
  {
    mmsam_host::EncodeTiledFn enc = mmsam_host::get_encode_tiled();
    cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)M};
    cuuint64_t strides[1] = {(cuuint64_t)ldt * 4};
    cuuint32_t box[2] = {16, 32};
    cuuint32_t estr[2] = {1, 1};
    if (!enc || enc(&tmT, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, t, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return MMSAM_ERR_DRIVER;
  }
  MlpParams p;
  p.colsum1 = colsum1; p.bias1 = bias1; p.bias2 = bias2; p.gamma = gamma; p.M = M; p.eps = eps;
  static const int dbg = [] { const char* e = getenv("MMSAM_MLP_DBG"); return e ? atoi(e) : 0; }();
  p.dbg = dbg;
  static const int want_trace = getenv("MMSAM_MLP_TRACE") != nullptr;
  if (want_trace && !g_mlp_trace_buf) {
    cudaMalloc(&g_mlp_trace_buf, 3 * 64 * 12 * sizeof(long long));
    cudaMemset(g_mlp_trace_buf, 0, 3 * 64 * 12 * sizeof(long long));
  }
  p.trace = want_trace ? g_mlp_trace_buf : nullptr;
  cudaStream_t st = (cudaStream_t)stream;
  if (C == 96) return launch_mlp<96>(tmA, tmW1, tmW2, tmT, p, max_ctas, st);
  if (C == 192) return launch_mlp<192>(tmA, tmW1, tmW2, tmT, p, max_ctas, st);
  return launch_mlp<384>(tmA, tmW1, tmW2, tmT, p, max_ctas, st);
}
