// ConvNeXt block tail in ONE kernel:   t += gamma * ( GELU( LN(y) W1^T + b1 ) W2^T + b2 )
// (twin_convnext.py:98-132: norm -> pwconv1 -> GELU -> pwconv2 -> gamma -> residual), y = the 7x7 depthwise conv output
// (bf16 [M, C]), t = the tower's fp32 residual stream [M, C], hidden width 4C. The 4C-wide intermediate never leaves the
// SM: per 256-row tile the two GEMMs are chained through TENSOR MEMORY, 64 hidden units at a time.
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
//   * per chunk j of 64 hidden units:  G1(j): Hacc[j&1] = A . W1[64j..64j+63, :]^T   (SS form, N = 64, K = C)
//                                      epilogue-1: Hacc -> LN fold, bias, exact GELU -> bf16, written back over the
//                                                   accumulator columns it came from (tcgen05.st)
//                                      G2(j): O += H(j) . W2[:, 64j..64j+63]^T        (TS form: A operand from TMEM)
//     issue order G1(j+1), G2(j): the tensor core always has the next chunk's first GEMM while the GELU of chunk j runs.
//   * O (128 lanes x C fp32 columns) accumulates over all 4C/64 chunks; the drain adds b2, scales by gamma and ADDS the
//     result into t with a TMA reduction (cp.reduce.async.bulk.tensor .add, fp32): the residual is never fetched by the
//     SM, every element of t receives exactly one add (deterministic).
//   * weights stream through two 2-CTA rings (each CTA stages half of the N rows of either operand, the tensor core
//     reads the other half from the peer): L2 -> SM traffic per tile = |W1| + |W2| per PAIR.
// TMEM: O [0, C), Hacc[0] [C, C+64), Hacc[1] [C+64, C+128).
#include "common.cuh"
#include "cg2.cuh"
#include <cstdlib>

namespace mmsam {

template <int C> struct MlpCfg {
  static constexpr int HID = 4 * C;
  static constexpr int HN = 64;                       // hidden units per chunk
  static constexpr int NCH = HID / HN;
  static constexpr int KB1 = (C + 63) / 64;           // 64-wide k-blocks of GEMM1 (the last may be zero-padded by TMA)
  static constexpr int KS1 = C / 16;                  // k-steps of GEMM1
  static constexpr int NSPLIT = C > 256 ? 2 : 1;      // GEMM2 instruction N = C / NSPLIT (<= 256)
  static constexpr int N2 = C / NSPLIT;
  static constexpr int A_BYTES = KB1 * 128 * 128;
  static constexpr int W1_STAGE = KB1 * (HN / 2) * 128;     // per CTA: 32 weight rows x KB1 k-blocks
  static constexpr int W2_STAGE = (C / 2) * 128;            // per CTA: C / 2 weight rows x 64 hidden columns
  static constexpr int NS = C > 256 ? 2 : 4;
  static constexpr int STAGING = 8 * 4096;                  // one 32-row x 128-byte slab per epilogue warp
  static constexpr int SMEM_BYTES = A_BYTES + NS * (W1_STAGE + W2_STAGE) + STAGING + 256 + 1024;
  static constexpr int TMEM_COLS = C + 128 <= 256 ? 256 : 512;
  static constexpr int THREADS = 352;                       // 8 epilogue warps + TMA + MMA + relay
  static_assert(C % 32 == 0 && C <= 384, "C");
  static_assert(W1_STAGE % 1024 == 0 && (W2_STAGE / NSPLIT) % 1024 == 0, "swizzle atoms");
};

struct MlpParams {
  const float* colsum1;   // [4C]  sum_k W1'[n, k]  (W1' = W1 * ln_weight, as stored in the bf16 weight)
  const float* bias1;     // [4C]  b1 + W1 . ln_bias
  const float* bias2;     // [C]
  const float* gamma;     // [C] or null
  int M;
  float eps;
  long long* trace;   // perf debug (MMSAM_MLP_TRACE): clock64 stamps of pair 0, [role 0..2][chunk < 64][event < 8]
  int dbg;   // perf debug (env MMSAM_MLP_DBG): 1 no drain reduce, 2 no epilogue-1 math, 4 no G1 MMAs, 8 no G2 MMAs, 16 no weight loads
};

#define MLP_TRACE(role, gi, ev)                                                                                   \
  do {                                                                                                            \
    if (p.trace && (blockIdx.x >> 1) == 0 && (gi) < 64) p.trace[((role) * 64 + (gi)) * 8 + (ev)] = clock64();   \
  } while (0)

template <int C>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(MlpCfg<C>::THREADS, 1)
convnext_mlp_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW1,
                    const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmT, const MlpParams p) {
  using Cfg = MlpCfg<C>;
  constexpr int NS = Cfg::NS, NCH = Cfg::NCH;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sW1 = sA + Cfg::A_BYTES;
  uint8_t* sW2 = sW1 + NS * Cfg::W1_STAGE;
  uint8_t* staging = sW2 + NS * Cfg::W2_STAGE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging + Cfg::STAGING);
  uint64_t* w1_full = bars;                  // [NS] leader: both CTAs' TMA bytes
  uint64_t* w1_empty = bars + NS;            // [NS] per CTA (multicast commit)
  uint64_t* w2_full = bars + 2 * NS;         // [NS] leader
  uint64_t* w2_empty = bars + 3 * NS;        // [NS] per CTA
  uint64_t* a_full = bars + 4 * NS;          // per CTA: this CTA's A tile has landed
  uint64_t* a_ready = a_full + 1;            // leader: both CTAs' A tiles have landed (relay warps)
  uint64_t* a_empty = a_full + 2;            // per CTA: every G1 of the tile has read A
  uint64_t* hacc_full = a_full + 3;          // [2] per CTA: G1 accumulator complete
  uint64_t* h_ready = a_full + 5;            // [2] leader: 16 epilogue warps wrote H
  uint64_t* o_full = a_full + 7;             // per CTA
  uint64_t* o_empty = a_full + 8;            // leader: 16 epilogue warps drained O
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_full + 9);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int num_tiles = (p.M + 255) / 256;
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmW1); tma_prefetch_desc(&tmW2); tma_prefetch_desc(&tmT);
    for (int i = 0; i < NS; ++i) {
      mbar_init(&w1_full[i], 2); mbar_init(&w1_empty[i], 1);
      mbar_init(&w2_full[i], 2); mbar_init(&w2_empty[i], 1);
    }
    mbar_init(a_full, 1); mbar_init(a_ready, 2); mbar_init(a_empty, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&hacc_full[i], 1); mbar_init(&h_ready[i], 16); }
    mbar_init(o_full, 1); mbar_init(o_empty, 16);
    fence_barrier_init();
  }
  cluster_sync_all();
  if (warp == 9) tmem_alloc_cg2(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr uint32_t TM_H = C;               // Hacc[b] at TM_H + 64 b

  if (warp == 8) {
    // ---------------- TMA producer (both CTAs) ----------------
    if (lane == 0) {
      int s1 = 0, s2 = 0; uint32_t ph1 = 0, ph2 = 0; int ti = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs, ++ti) {
        const int m0 = tile * 256 + (int)rank * 128;
        mbar_wait(a_empty, (ti & 1) ^ 1);
        mbar_arrive_expect_tx(a_full, Cfg::A_BYTES);
#pragma unroll
        for (int kb = 0; kb < Cfg::KB1; ++kb) tma_load_2d(sA + kb * 16384, &tmA, a_full, kb * 64, m0);
        for (int j = 0; j < NCH; ++j) {
          {
            mbar_wait(&w1_empty[s1], ph1 ^ 1);
            const uint32_t lead = mapa_shared(smem_u32(&w1_full[s1]), 0);
            if (rank == 0) mbar_arrive_expect_tx(&w1_full[s1], (p.dbg & 16) ? 0 : 2 * Cfg::W1_STAGE);
            else mbar_arrive_cluster(lead);
            const uint32_t dst = smem_u32(sW1 + s1 * Cfg::W1_STAGE);
#pragma unroll
            for (int kb = 0; kb < Cfg::KB1 && !(p.dbg & 16); ++kb)
              tma_load_2d_cg2(dst + kb * (Cfg::HN / 2) * 128, &tmW1, lead, kb * 64, j * Cfg::HN + (int)rank * (Cfg::HN / 2));
            if (++s1 == NS) { s1 = 0; ph1 ^= 1; }
          }
          {
            mbar_wait(&w2_empty[s2], ph2 ^ 1);
            const uint32_t lead = mapa_shared(smem_u32(&w2_full[s2]), 0);
            if (rank == 0) mbar_arrive_expect_tx(&w2_full[s2], (p.dbg & 16) ? 0 : 2 * Cfg::W2_STAGE);
            else mbar_arrive_cluster(lead);
            const uint32_t dst = smem_u32(sW2 + s2 * Cfg::W2_STAGE);
#pragma unroll
            for (int h = 0; h < Cfg::NSPLIT && !(p.dbg & 16); ++h)
              tma_load_2d_cg2(dst + h * (Cfg::N2 / 2) * 128, &tmW2, lead, j * Cfg::HN, h * Cfg::N2 + (int)rank * (Cfg::N2 / 2));
            if (++s2 == NS) { s2 = 0; ph2 ^= 1; }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 10) {
    // ---------------- relay: tell the leader's MMA thread that this CTA's A tile has landed ----------------
    if (lane == 0) {
      const uint32_t lead = mapa_shared(smem_u32(a_ready), 0);
      int ti = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs, ++ti) {
        mbar_wait(a_full, ti & 1);
        mbar_arrive_cluster(lead);
      }
    }
    __syncwarp();
  } else if (warp == 9) {
    // ---------------- MMA issuer (leader CTA only) ----------------
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc1 = umma_idesc_bf16(256, Cfg::HN, 0, 0);
      constexpr uint32_t idesc2 = umma_idesc_bf16(256, Cfg::N2, 0, 0);
      int s1 = 0, s2 = 0; uint32_t ph1 = 0, ph2 = 0;
      const uint32_t a_addr = smem_u32(sA);
      auto issue_g2 = [&](int gp) {
        const int jp = gp % NCH, tp = gp / NCH, b = gp & 1;
        MLP_TRACE(0, gp, 3);
        mbar_wait(&h_ready[b], (gp >> 1) & 1);
        MLP_TRACE(0, gp, 4);
        mbar_wait(&w2_full[s2], ph2);
        if (jp == 0) mbar_wait(o_empty, (tp & 1) ^ 1);
        tc_fence_after();
        MLP_TRACE(0, gp, 5);
        const uint32_t w_addr = smem_u32(sW2 + s2 * Cfg::W2_STAGE);
        const uint32_t h_tmem = tmem_base + TM_H + b * 64;
#pragma unroll
        for (int kk = 0; kk < 4 && !(p.dbg & 8); ++kk) {
          // H k-steps 0,1 live in columns [0, 16) of the buffer, k-steps 2,3 in [32, 48) (each column half is rewritten
          // in place by the warps that read it)
          const uint32_t a_t = h_tmem + (kk < 2 ? kk * 8 : 32 + (kk - 2) * 8);
#pragma unroll
          for (int h = 0; h < Cfg::NSPLIT; ++h) {
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "setp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                ::"r"(tmem_base + h * Cfg::N2), "r"(a_t), "l"(umma_desc_sw128(w_addr + h * (Cfg::N2 / 2) * 128 + kk * 32)),
                  "r"(idesc2), "r"((jp | kk) != 0 ? 1u : 0u)
                : "memory");
          }
        }
        umma_commit_cg2(&w2_empty[s2]);
        MLP_TRACE(0, gp, 6);
        if (jp == NCH - 1) umma_commit_cg2(o_full);
        if (++s2 == NS) { s2 = 0; ph2 ^= 1; }
      };
      int g = 0, ti = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs, ++ti) {
        mbar_wait(a_ready, ti & 1);
        for (int j = 0; j < NCH; ++j, ++g) {
          MLP_TRACE(0, g, 0);
          mbar_wait(&w1_full[s1], ph1);
          tc_fence_after();
          MLP_TRACE(0, g, 1);
          const uint32_t w_addr = smem_u32(sW1 + s1 * Cfg::W1_STAGE);
          const uint32_t d_tmem = tmem_base + TM_H + (g & 1) * 64;
#pragma unroll
          for (int ks = 0; ks < Cfg::KS1 && !(p.dbg & 4); ++ks) {
            const int kb = ks >> 2, k = ks & 3;
            umma_f16_ss_cg2(d_tmem, umma_desc_sw128(a_addr + kb * 16384 + k * 32),
                            umma_desc_sw128(w_addr + kb * (Cfg::HN / 2) * 128 + k * 32), idesc1, ks != 0 ? 1u : 0u);
          }
          umma_commit_cg2(&w1_empty[s1]);
          umma_commit_cg2(&hacc_full[g & 1]);
          MLP_TRACE(0, g, 2);
          if (j == NCH - 1) umma_commit_cg2(a_empty);
          if (++s1 == NS) { s1 = 0; ph1 ^= 1; }
          if (g > 0) issue_g2(g - 1);
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
    const uint32_t slab = smem_u32(staging) + (uint32_t)warp * 4096;
    const uint32_t peer_slab = smem_u32(staging) + (uint32_t)(warp ^ 4) * 4096;
    const uint32_t h_ready_lead[2] = {mapa_shared(smem_u32(&h_ready[0]), 0), mapa_shared(smem_u32(&h_ready[1]), 0)};
    const uint32_t o_empty_lead = mapa_shared(smem_u32(o_empty), 0);
    constexpr int NP = C / 32;                                        // 32-column fp32 panels of O
    constexpr int NP0 = (NP + 1) / 2;
    const int pan_lo = half == 0 ? 0 : NP0, pan_hi = half == 0 ? NP0 : NP;
    int g = 0, ti = 0;
    for (int tile = pair; tile < num_tiles; tile += num_pairs, ++ti) {
      // ---- row statistics of y from the resident tile: this warp sums its half of the k-blocks ----
      if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the slab's last reduce has read it
      __syncwarp();
      mbar_wait(a_full, ti & 1);
      float s = 0.f, ss = 0.f;
      {
        constexpr int KBH = (Cfg::KB1 + 1) / 2;
        const int kb_lo = half == 0 ? 0 : KBH, kb_hi = half == 0 ? KBH : Cfg::KB1;
        const uint32_t a_row = smem_u32(sA) + row_l * 128;
        for (int kb = kb_lo; kb < kb_hi; ++kb) {
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            float f[8];
            unpack8(lds128(a_row + kb * 16384 + ((c ^ (row_l & 7)) << 4)), f);
#pragma unroll
            for (int i = 0; i < 8; ++i) { s += f[i]; ss = fmaf(f[i], f[i], ss); }
          }
        }
      }
      asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(slab + lane * 8), "f"(s), "f"(ss) : "memory");
      asm volatile("bar.sync %0, 64;" ::"r"(1 + quad) : "memory");
      {
        float s2, ss2;
        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(s2), "=f"(ss2) : "r"(peer_slab + lane * 8));
        s += s2; ss += ss2;
      }
      const float mean = s * (1.f / C);
      const float rstd = rsqrtf(fmaxf(ss * (1.f / C) - mean * mean, 0.f) + p.eps);
      const float nmean = -mean;

      // ---- per hidden chunk: Hacc -> LN fold + bias + GELU -> bf16 back into TMEM ----
      for (int j = 0; j < NCH; ++j, ++g) {
        const int b = g & 1;
        const int n0 = j * Cfg::HN + half * 32;
        float4 cs[8], bi[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          cs[i] = __ldg(reinterpret_cast<const float4*>(p.colsum1 + n0) + i);
          bi[i] = __ldg(reinterpret_cast<const float4*>(p.bias1 + n0) + i);
        }
        if (warp == 0 && lane == 0) MLP_TRACE(1 + rank, g, 0);
        mbar_wait(&hacc_full[b], (g >> 1) & 1);
        tc_fence_after();
        if (warp == 0 && lane == 0) MLP_TRACE(1 + rank, g, 1);
        const uint32_t t_addr = lane_addr + TM_H + b * 64 + half * 32;
        uint32_t r[32];
        tmem_ld_32x32b_x32(t_addr, r);
        tmem_ld_wait();
        if (warp == 0 && lane == 0) MLP_TRACE(1 + rank, g, 2);
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 8 && !(p.dbg & 2); ++i) {
          float v0 = fmaf(rstd, fmaf(nmean, cs[i].x, __uint_as_float(r[4 * i])), bi[i].x);
          float v1 = fmaf(rstd, fmaf(nmean, cs[i].y, __uint_as_float(r[4 * i + 1])), bi[i].y);
          float v2 = fmaf(rstd, fmaf(nmean, cs[i].z, __uint_as_float(r[4 * i + 2])), bi[i].z);
          float v3 = fmaf(rstd, fmaf(nmean, cs[i].w, __uint_as_float(r[4 * i + 3])), bi[i].w);
          gelu_erf2(v0, v1);
          gelu_erf2(v2, v3);
          pk[2 * i] = pack_bf16(v0, v1);
          pk[2 * i + 1] = pack_bf16(v2, v3);
        }
        if (warp == 0 && lane == 0) MLP_TRACE(1 + rank, g, 3);
        tmem_st_32x32b_x16(t_addr, pk);
        tmem_st_wait();
        if (warp == 0 && lane == 0) MLP_TRACE(1 + rank, g, 4);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(h_ready_lead[b]);
        if (warp == 0 && lane == 0) MLP_TRACE(1 + rank, g, 5);
      }

      // ---- drain: O -> (+ b2) * gamma -> slab -> TMA reduce-add into t ----
      mbar_wait(o_full, ti & 1);
      tc_fence_after();
      const int row0 = tile * 256 + (int)rank * 128 + quad * 32;
      for (int pi = pan_lo; pi < pan_hi; ++pi) {
        const int col0 = pi * 32;
        uint32_t r[32];
        tmem_ld_32x32b_x32(lane_addr + col0, r);
        tmem_ld_wait();
        if (pi == pan_hi - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(o_empty_lead);
        }
        uint4 out[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 b2 = __ldg(reinterpret_cast<const float4*>(p.bias2 + col0) + i);
          float4 gm = make_float4(1.f, 1.f, 1.f, 1.f);
          if (p.gamma) gm = __ldg(reinterpret_cast<const float4*>(p.gamma + col0) + i);
          out[i] = make_uint4(__float_as_uint((__uint_as_float(r[4 * i]) + b2.x) * gm.x),
                              __float_as_uint((__uint_as_float(r[4 * i + 1]) + b2.y) * gm.y),
                              __float_as_uint((__uint_as_float(r[4 * i + 2]) + b2.z) * gm.z),
                              __float_as_uint((__uint_as_float(r[4 * i + 3]) + b2.w) * gm.w));
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; ++i) sts128(slab + swz128(lane, i), out[i]);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0 && !(p.dbg & 1)) {
          asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                           reinterpret_cast<uint64_t>(&tmT)),
                       "r"(slab), "r"(col0), "r"(row0)
                       : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
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
  convnext_mlp_kernel<C><<<2 * pairs, Cfg::THREADS, Cfg::SMEM_BYTES, st>>>(tmA, tmW1, tmW2, tmT, p);
  MMSAM_LAUNCH_CHECK();
  return MMSAM_OK;
}

}  // namespace mmsam

static long long* g_mlp_trace_buf = nullptr;
// perf debug: copy the clock64 trace of the last traced launch to the host (3 x 64 x 8 values); returns 0 when tracing is off
extern "C" __attribute__((visibility("default"))) int mmsam_dbg_mlp_trace(long long* host_out) {
  if (!g_mlp_trace_buf) return 0;
  cudaDeviceSynchronize();
  cudaMemcpy(host_out, g_mlp_trace_buf, 3 * 64 * 8 * sizeof(long long), cudaMemcpyDeviceToHost);
  return 1;
}

// See include/mmsam_b200.h for the contract.
MMSAM_API int mmsam_convnext_mlp_bf16(const void* y, long long ldy, const void* W1, const float* colsum1, const float* bias1,
                                      const void* W2, const float* bias2, const float* gamma, float* t, long long ldt, int M,
                                      int C, float eps, int max_ctas, void* stream) {
  using namespace mmsam;
  if (M < 0) return MMSAM_ERR_BAD_ARG;
  if (M == 0) return MMSAM_OK;
  if (!y || !W1 || !colsum1 || !bias1 || !W2 || !bias2 || !t) return MMSAM_ERR_BAD_ARG;
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
  rc = mmsam_host::make_tmap_2d_bf16(&tmW1, W1, (uint64_t)HID, (uint64_t)C, (uint64_t)C, 32, 64);
  if (rc) return rc;
  rc = mmsam_host::make_tmap_2d_bf16(&tmW2, W2, (uint64_t)C, (uint64_t)HID, (uint64_t)HID, (uint32_t)(n2 / 2), 64);
  if (rc) return rc;
  {
    mmsam_host::EncodeTiledFn enc = mmsam_host::get_encode_tiled();
    cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)M};
    cuuint64_t strides[1] = {(cuuint64_t)ldt * 4};
    cuuint32_t box[2] = {32, 32};
    cuuint32_t estr[2] = {1, 1};
    if (!enc || enc(&tmT, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, t, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return MMSAM_ERR_DRIVER;
  }
  MlpParams p;
  p.colsum1 = colsum1; p.bias1 = bias1; p.bias2 = bias2; p.gamma = gamma; p.M = M; p.eps = eps;
  static const int dbg = [] { const char* e = getenv("MMSAM_MLP_DBG"); return e ? atoi(e) : 0; }();
  p.dbg = dbg;
  static const int want_trace = getenv("MMSAM_MLP_TRACE") != nullptr;
  if (want_trace && !g_mlp_trace_buf) {
    cudaMalloc(&g_mlp_trace_buf, 3 * 64 * 8 * sizeof(long long));
    cudaMemset(g_mlp_trace_buf, 0, 3 * 64 * 8 * sizeof(long long));
  }
  p.trace = want_trace ? g_mlp_trace_buf : nullptr;
  cudaStream_t st = (cudaStream_t)stream;
  if (C == 96) return launch_mlp<96>(tmA, tmW1, tmW2, tmT, p, max_ctas, st);
  if (C == 192) return launch_mlp<192>(tmA, tmW1, tmW2, tmT, p, max_ctas, st);
  return launch_mlp<384>(tmA, tmW1, tmW2, tmT, p, max_ctas, st);
}
