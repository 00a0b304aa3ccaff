// Fused multi-head attention with SAM's decomposed relative-position bias, head_dim 64,
// on tcgen05 tensor cores (TMEM accumulators, TMA-fed), flash-style: the score matrix never
// leaves the SM.
//
// Replaces Attention.forward + add_decomposed_rel_pos
// (segmentation/mmseg_custom/models/backbones/base/image_encoder.py:483-501, 587-623):
//     attn = softmax((q*scale) k^T + rel_h[q, kh] + rel_w[q, kw]);  out = attn v
// with rel_h[q,kh] = q . Rh[qh-kh+Kh-1], rel_w[q,kw] = q . Rw[qw-kw+Kw-1] computed from the
// UNSCALED q (image_encoder.py:492-495). Window blocks call it with Bp = B*windows, T = 196,
// (Kh,Kw) = (14,14) — the zero-padded window tokens are ordinary keys (no masking, image_encoder.py
// :504-526); global blocks with Bp = B, T = H*W.
//
// One CTA per SM, persistent over (batch', head, 128-query tile); 10 warps:
//   warp 8  TMA producer: Q tile, K/V 128-key blocks (SWIZZLE_128B boxes straight out of the
//           [Bp,T,3,nh,64] qkv GEMM output through a 4-D tensor map), rel-pos tables once per CTA.
//   warp 9  MMA issuer: G = Q.table^T (bias pre-products), S_j = Q.K_j^T into a double-buffered
//           TMEM tile, O_half += P_j[:, half].V_j[half] (V as an MN-major operand) into two TMEM accumulators.
//   warps 0-7  softmax: thread = query row (tcgen05.ld 32x32b), warps w / w+4 split each key block in halves
//           with independent online-softmax state. Scatter G into per-row bias rows in shared memory (Toeplitz
//           gather), then per key block: t = S*scale + bias, max / sum in base 2, P -> bf16 into the swizzled
//           K-major smem panel of the half. The output accumulates in TMEM across key blocks and is rescaled
//           only when a row maximum grows by more than 2^8 (lazy rescale); halves are merged per tile.
#include "common.cuh"
#include <cstdlib>
#include <type_traits>

namespace mmsam {

struct AttnParams {
  __nv_bfloat16* out;   // [Bp, T, nh*64], or [rows, nh*64] indexed through out_map
  const int* out_map;   // optional: destination row of each (bp, t) row, -1 = drop (window un-partition)
  int Bp, T, nh, Kh, Kw;
  int nh_rows, nw_rows;        // table rows (2K-1), 0 = no relative position bias
  int nh_pad, nw_pad;          // padded to a multiple of 16
  int kv_stages;               // 2 or 3
  int dbg;                     // perf-debug switches (MMSAM_ATT_DBG): 1 no bias scatter, 2 no bias add, 4 no pair barrier
  int q_bufs;                  // 1 or 2 Q tiles in shared memory (2: next tile's prologue overlaps this tile's tail)
  int bh_stride, bw_stride;    // floats per bias row in shared memory
  float scale_log2;            // scale * log2(e)
  int num_tiles, nqb;
};

static constexpr int ATT_BM = 128, ATT_BN = 128, ATT_D = 64;
static constexpr int TILE_BYTES = 128 * 64 * 2;  // one 128 x 64 bf16 SW128 tile
static constexpr float kLog2e = 1.4426950408889634f;

// TMEM column plan
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

static constexpr int TM_S = 0;      // 2 x 128: double-buffered score tile
static constexpr int TM_O = 256;    // 2 x 64: output accumulators of the two key halves
static constexpr int TM_G = 384;    // 128: bias pre-products
static constexpr float kRescaleThreshold = 8.f;   // log2 units: P stays <= 256 with a stale row maximum

// u = score * scale + bias (minus a per-segment constant kept in segb) for this warp's 64 key columns [k0h, k0h+64)
// of the score row, as 32 fp32 pairs; returns the row maximum of the full logits. Segments: columns [0,32) / [32,64).
__device__ __forceinline__ float ldb(const float* p) { return *p; }
__device__ __forceinline__ float ldb(const __half* p) { return __half2float(*p); }
__device__ __forceinline__ void stb(float* p, float v) { *p = v; }
__device__ __forceinline__ void stb(__half* p, float v) { *p = __float2half_rn(v); }

template <int KW, typename BT>
__device__ __forceinline__ float scores_to_logits(const uint32_t (&r)[64], u64 (&u)[32], float (&segb)[2], int k0h,
                                                  const BT* bh, const BT* bw, bool has_bias, float scale_log2,
                                                  int Kh, int Kw) {
  const u64 sc2 = pack2(scale_log2, scale_log2);
  float mx0 = -INFINITY, mx1 = -INFINITY;
  segb[0] = segb[1] = 0.f;
  if (!has_bias) {
#pragma unroll
    for (int i = 0; i < 32; ++i) u[i] = mul2(pack2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1])), sc2);
  } else if ((KW == 64 || KW == 32) && std::is_same<BT, float>::value) {
    // the key-row (kh) part of the bias is constant over a segment: it is folded into the exponent offset instead
    // of being added to every element; the key-column part is a 64-bit shared-memory load per pair
    if (KW == 64) { segb[0] = segb[1] = ldb(bh + (k0h >> 6)); }
    else { segb[0] = ldb(bh + (k0h >> 5)); segb[1] = ldb(bh + (k0h >> 5) + 1); }
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const float2 b = *reinterpret_cast<const float2*>(reinterpret_cast<const float*>(bw) + ((2 * i) & (KW - 1)));
      u[i] = fma2(pack2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1])), sc2, pack2(b.x, b.y));
    }
  } else if (KW == 14) {
    // SAM window (14 x 14 keys): k0h is 0, 64, 128 or 192 -> (kh, kw) of every column are compile-time constants.
    // The 7 key-column pairs of the row's bw entries are read once as 64-bit words, the <= 6 key-row entries this
    // half touches once as scalars; a column pair never straddles a key row (14 is even), so its bias is ONE packed
    // add of two registers instead of two shared-memory reads and two scalar adds per element.
    auto body = [&](auto k0c) {
      constexpr int K0 = decltype(k0c)::value;
      constexpr int KH0 = K0 / 14, KH1 = (K0 + 63) / 14 < 13 ? (K0 + 63) / 14 : 13;
      u64 bw2[7], bh2[KH1 - KH0 + 1];
#pragma unroll
      for (int t = 0; t < 7; ++t) bw2[t] = *(reinterpret_cast<const u64*>(bw) + t);
#pragma unroll
      for (int t = 0; t <= KH1 - KH0; ++t) { const float v = ldb(bh + KH0 + t); bh2[t] = pack2(v, v); }
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const int k0 = K0 + 2 * i;
        const u64 b2 = k0 < 196 ? add2(bw2[(k0 % 14) >> 1], bh2[(k0 / 14 < KH1 ? k0 / 14 : KH1) - KH0]) : 0ull;
        u[i] = fma2(pack2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1])), sc2, b2);
      }
    };
    switch (k0h >> 6) {
      case 0: body(std::integral_constant<int, 0>{}); break;
      case 1: body(std::integral_constant<int, 64>{}); break;
      case 2: body(std::integral_constant<int, 128>{}); break;
      default: body(std::integral_constant<int, 192>{}); break;
    }
  } else {
    // generic grid: branch-free running (kh, kw)
    int kh = k0h / Kw, kw = k0h - kh * Kw;
    float bb[2];
#pragma unroll
    for (int i = 0; i < 32; ++i) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int khc = kh < Kh ? kh : Kh - 1;
        bb[e] = ldb(bh + khc) + ldb(bw + kw);
        ++kw;
        const bool wrap = kw == Kw;
        kw = wrap ? 0 : kw;
        kh += wrap ? 1 : 0;
      }
      u[i] = fma2(pack2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1])), sc2, pack2(bb[0], bb[1]));
    }
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    float a0, a1, c0, c1;
    unpack2(u[i], a0, a1);
    unpack2(u[16 + i], c0, c1);
    mx0 = max3(mx0, a0, a1);
    mx1 = max3(mx1, c0, c1);
  }
  return fmaxf(mx0 + segb[0], mx1 + segb[1]);
}

// BT: element type of the per-row bias rows in shared memory (float; __half for key grids whose fp32 rows would not
// fit beside the K/V ring, e.g. a whole 1088 x 1920 MUSES frame = 68 x 120 tokens).
template <int KW, typename BT>
__global__ void __launch_bounds__(320, 1)
attention_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmTabH,
                 const __grid_constant__ CUtensorMap tmTabW, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  // keep the pointer derived from the __shared__ array (so loads compile to LDS, not generic LD)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  // layout: Q (2 buffers) | KV stages | P (2 panels) | tables | bias rows | (m, l) exchange | barriers
  uint8_t* sQ = smem;                       // q_bufs tiles: Q of the CTA's i-th tile lives in buffer i % q_bufs
  uint8_t* sKV = sQ + p.q_bufs * TILE_BYTES;
  uint8_t* sP = sKV + p.kv_stages * 2 * TILE_BYTES;
  uint8_t* sTab = sP + 2 * TILE_BYTES;
  const int tab_bytes = (p.nh_pad + p.nw_pad) * 128;
  float* sBias = reinterpret_cast<float*>(sTab + ((tab_bytes + 1023) & ~1023));
  const int bh_stride = p.bh_stride, bw_stride = p.bw_stride;  // row-private rows (strides: see the host code)
  BT* sBh = reinterpret_cast<BT*>(sBias);
  BT* sBw = sBh + ATT_BM * bh_stride;
  float* sML = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(sBw + ATT_BM * bw_stride) + 15) & ~(uintptr_t)15);   // [2 halves][128 rows][m, l]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sML + 2 * ATT_BM * 2);
  uint64_t* q_full = bars + 0;    // [2] at bars + 0, bars + 17
  uint64_t* q_empty = bars + 1;   // [2] at bars + 1, bars + 18
  uint64_t* g_full = bars + 2;
  uint64_t* g_empty = bars + 3;
  uint64_t* p_full = bars + 4;
  uint64_t* pv_done = bars + 5;
  uint64_t* tab_full = bars + 6;
  uint64_t* s_full = bars + 7;    // [2]
  uint64_t* s_empty = bars + 9;   // [2]
  uint64_t* kv_full = bars + 11;  // [3]
  uint64_t* kv_empty = bars + 14; // [3]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 19);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool has_bias = p.nh_rows > 0;
  const int nkb = (p.T + ATT_BN - 1) / ATT_BN;
  // G chunks of <= 128 table rows each: first the h table, then the w table
  const int nch_h = has_bias ? (p.nh_pad + 127) / 128 : 0;
  const int nch_w = has_bias ? (p.nw_pad + 127) / 128 : 0;
  const bool g_merged = has_bias && p.nh_pad + p.nw_pad <= 128;
  const int ng_chunks = g_merged ? 1 : nch_h + nch_w;

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    for (int i = 0; i < 7; ++i) mbar_init(&bars[i], 1);
    mbar_init(bars + 17, 1);
    mbar_init(bars + 18, 1);
    mbar_init(g_empty, 8);
    mbar_init(p_full, 8);
    for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], 8); }
    for (int i = 0; i < 3; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
    fence_barrier_init();
  }
  if (warp == 9) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 8) {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      if (has_bias) {
        mbar_arrive_expect_tx(tab_full, tab_bytes);
        tma_load_2d(sTab, &tmTabH, tab_full, 0, 0);
        tma_load_2d(sTab + p.nh_pad * 128, &tmTabW, tab_full, 0, 0);
      }
      int st = 0; uint32_t kph = 0; uint32_t it = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
        const int qb = tile % p.nqb;
        const int bh = tile / p.nqb;
        const int head = bh % p.nh, bp = bh / p.nh;
        const uint32_t qi = p.q_bufs == 2 ? (it & 1) : 0, qph = p.q_bufs == 2 ? ((it >> 1) & 1) : (it & 1);
        uint64_t* qf = q_full + qi * 17;
        mbar_wait(q_empty + qi * 17, qph ^ 1);
        mbar_arrive_expect_tx(qf, TILE_BYTES);
        asm volatile(
            "cp.async.bulk.tensor.4d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
            ::"r"(smem_u32(sQ + qi * TILE_BYTES)), "l"(reinterpret_cast<uint64_t>(&tmQKV)), "r"(smem_u32(qf)), "r"(0),
            "r"(head), "r"(qb * ATT_BM), "r"(bp)
            : "memory");
        for (int j = 0; j < nkb; ++j) {
          mbar_wait(&kv_empty[st], kph ^ 1);
          mbar_arrive_expect_tx(&kv_full[st], 2 * TILE_BYTES);
          uint8_t* sk = sKV + st * 2 * TILE_BYTES;
          asm volatile(
              "cp.async.bulk.tensor.4d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
              ::"r"(smem_u32(sk)), "l"(reinterpret_cast<uint64_t>(&tmQKV)), "r"(smem_u32(&kv_full[st])), "r"(0),
              "r"(p.nh + head), "r"(j * ATT_BN), "r"(bp)
              : "memory");
          asm volatile(
              "cp.async.bulk.tensor.4d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
              ::"r"(smem_u32(sk + TILE_BYTES)), "l"(reinterpret_cast<uint64_t>(&tmQKV)), "r"(smem_u32(&kv_full[st])),
              "r"(0), "r"(2 * p.nh + head), "r"(j * ATT_BN), "r"(bp)
              : "memory");
          if (++st == p.kv_stages) { st = 0; kph ^= 1; }
        }
      }
    }
  } else if (warp == 9) {
    // =========================== MMA issuer: the whole warp runs the loop, one elected lane issues (see elect_one) ===========================
    {
      constexpr uint32_t idesc_qk = umma_idesc_bf16(128, ATT_BN, 0, 0);
      constexpr uint32_t idesc_pv = umma_idesc_bf16(128, ATT_D, 0, 1);  // V is MN-major
      if (has_bias) { mbar_wait(tab_full, 0); }
      int st = 0; uint32_t kph = 0;       // kv ring (QK side)
      int st_pv = 0;                      // kv ring (PV side)
      uint32_t gph = 0, pph = 0;
      uint32_t g = 0;                     // global key-block counter (S buffer parity)
      uint32_t q_addr = smem_u32(sQ);     // Q buffer of the tile whose S / G products are being issued
      const uint32_t p_addr = smem_u32(sP);
      // O_half (+)= P[:, half] . V[half]: keys 0..63 of the block accumulate into O_a, keys 64..127 into O_b,
      // across all key blocks of the tile (the first block overwrites)
      auto issue_pv = [&](int stage, bool first) {
        mbar_wait(p_full, pph);
        pph ^= 1;
        tc_fence_after();
        const uint32_t v_addr = smem_u32(sKV + stage * 2 * TILE_BYTES + TILE_BYTES);
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < ATT_BN / 16; ++kk) {
            const uint32_t a = p_addr + (kk >> 2) * TILE_BYTES + (kk & 3) * 32;
            const uint32_t b = v_addr + kk * 16 * 128;
            umma_f16_ss(tmem + TM_O + (kk >> 2) * ATT_D, umma_desc_sw128(a), umma_desc_sw128(b), idesc_pv,
                        (first && (kk & 3) == 0) ? 0u : 1u);
          }
          umma_commit(pv_done);
          umma_commit(&kv_empty[stage]);
        }
        __syncwarp();
      };
      // S_j = Q . K_j^T into the double-buffered score tile
      auto issue_qk = [&]() {
        mbar_wait(&kv_full[st], kph);
        mbar_wait(&s_empty[g & 1], ((g >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t k_addr = smem_u32(sKV + st * 2 * TILE_BYTES);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16_ss(tmem + TM_S + (g & 1) * ATT_BN, umma_desc_sw128(q_addr + k * 32),
                        umma_desc_sw128(k_addr + k * 32), idesc_qk, k != 0 ? 1u : 0u);
          umma_commit(&s_full[g & 1]);
        }
        __syncwarp();
        if (++st == p.kv_stages) { st = 0; kph ^= 1; }
        ++g;
      };
      // Tile prologue: wait for Q, then S_0 (needs only Q and the first key block) and the bias pre-products
      // G = Q . table^T through TM_G — one chunk when both tables fit its 128 columns (SAM windows: 32 + 32 rows; the
      // tables sit back to back in shared memory), else <= 128 rows at a time. It is issued BEFORE the previous tile's
      // last P.V, so these products (and their hand-shakes with the softmax warps) overlap the previous tile's tail.
      auto prologue = [&](uint32_t it) {
        const uint32_t qi = p.q_bufs == 2 ? (it & 1) : 0, qph = p.q_bufs == 2 ? ((it >> 1) & 1) : (it & 1);
        mbar_wait(q_full + qi * 17, qph);
        tc_fence_after();
        q_addr = smem_u32(sQ + qi * TILE_BYTES);
        issue_qk();
        for (int ci = 0; ci < ng_chunks; ++ci) {
          const bool is_w = !g_merged && ci >= nch_h;
          const int c0 = g_merged ? 0 : (is_w ? ci - nch_h : ci) * 128;
          const int npad = g_merged ? p.nh_pad + p.nw_pad : (is_w ? p.nw_pad : p.nh_pad);
          const int n = npad - c0 < 128 ? npad - c0 : 128;
          mbar_wait(g_empty, gph ^ 1);
          tc_fence_after();
          const uint32_t t_addr = smem_u32(sTab) + ((is_w ? p.nh_pad : 0) + c0) * 128;
          const uint32_t idesc_g = umma_idesc_bf16(128, n, 0, 0);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16_ss(tmem + TM_G, umma_desc_sw128(q_addr + k * 32), umma_desc_sw128(t_addr + k * 32), idesc_g,
                          k != 0 ? 1u : 0u);
            umma_commit(g_full);
          }
          __syncwarp();
          gph ^= 1;
        }
        if (nkb == 1) {                                 // every MMA that reads this Q has been issued
          if (elect_one()) umma_commit(q_empty + qi * 17);
          __syncwarp();
        }
      };
      uint32_t it = 0;
      if ((int)blockIdx.x < p.num_tiles) prologue(0);
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
        // ---- main loop: QK_j issued ahead of PV_{j-1} ----
        for (int j = 1; j < nkb; ++j) {
          issue_qk();
          if (j == nkb - 1) {
            if (elect_one()) umma_commit(q_empty + (p.q_bufs == 2 ? (it & 1) : 0) * 17);
            __syncwarp();
          }
          issue_pv(st_pv, j == 1);
          if (++st_pv == p.kv_stages) st_pv = 0;
        }
        const bool more = tile + (int)gridDim.x < p.num_tiles;
        if (more && p.q_bufs == 2) prologue(it + 1);      // with one Q buffer the wait for Q would stall the last P.V
        issue_pv(st_pv, nkb == 1);
        if (++st_pv == p.kv_stages) st_pv = 0;
        if (more && p.q_bufs == 1) prologue(it + 1);
      }
    }
  } else {
    // =========================== softmax / output (warps 0..7) ===========================
    // Warps w and w+4 share TMEM lane quadrant w&3 (thread = query row) and split every 128-key block: half 0
    // takes keys 0..63, half 1 keys 64..127, each as an INDEPENDENT online softmax with its own row maximum /
    // sum and its own output accumulator in TMEM (O_a / O_b, accumulated by the tensor core across key blocks).
    // The two halves are merged once per tile. Two softmax warps per scheduler hide each other's MUFU / TMEM /
    // shared-memory latency; the accumulator is only touched when a row maximum grows by more than 2^8.
    const int quad = warp & 3, half = warp >> 2;
    const int row = quad * 32 + lane;  // TMEM lane == query row within the tile
    const uint32_t lane_addr = tmem + ((uint32_t)(quad * 32) << 16);
    BT* bh = sBh + row * bh_stride;
    BT* bw = sBw + row * bw_stride;
    uint32_t gph = 0, pvph = 0;
    uint32_t g = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int qb = tile % p.nqb;
      const int bhid = tile / p.nqb;
      const int head = bhid % p.nh, bp = bhid / p.nh;
      const int q = qb * ATT_BM + row;
      const int qh = q / p.Kw, qw = q - qh * p.Kw;
      // ---- scatter the bias pre-products into this row's bias rows (pre-scaled by log2 e); the two halves
      //      take alternate 16-column groups ----
      for (int ci = 0; ci < ng_chunks; ++ci) {
        const bool chunk_w = !g_merged && ci >= nch_h;
        const int c0 = g_merged ? 0 : (chunk_w ? ci - nch_h : ci) * 128;
        const int npad = g_merged ? p.nh_pad + p.nw_pad : (chunk_w ? p.nw_pad : p.nh_pad);
        const int n = npad - c0 < 128 ? npad - c0 : 128;
        mbar_wait(g_full, gph);
        gph ^= 1;
        tc_fence_after();
        for (int c = half * 16; c < n; c += 32) {
          // a 16-column group belongs to one table (both are padded to multiples of 16 rows)
          const bool is_w = g_merged ? c >= p.nh_pad : chunk_w;
          const int cl = g_merged && is_w ? c - p.nh_pad : c;
          const int K1 = is_w ? p.Kw : p.Kh;
          const int qpos = is_w ? qw : qh;
          BT* dst = is_w ? bw : bh;
          if (p.dbg & 1) continue;
          uint32_t r[16];
          __syncwarp();
          tmem_ld_32x32b_x16(lane_addr + TM_G + c, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int kpos = qpos + K1 - 1 - (c0 + cl + i);  // table row r = qpos - kpos + K1 - 1
            if (kpos >= 0 && kpos < K1) stb(dst + kpos, __uint_as_float(r[i]) * kLog2e);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(g_empty);
      }
      if (has_bias && !(p.dbg & 4)) asm volatile("bar.sync %0, 64;" ::"r"(1 + quad) : "memory");   // both halves' bias entries are in place
      float m_ref = -INFINITY, l_run = 0.f;
      for (int j = 0; j < nkb; ++j, ++g) {
        // ---- 1. this half's 64 score columns of key block j ----
        mbar_wait(&s_full[g & 1], (g >> 1) & 1);
        tc_fence_after();
        uint32_t r[64];
        __syncwarp();
        tmem_ld_32x32b_x32(lane_addr + TM_S + (g & 1) * ATT_BN + half * 64, r);
        tmem_ld_32x32b_x32(lane_addr + TM_S + (g & 1) * ATT_BN + half * 64 + 32, r + 32);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_empty[g & 1]);
        const int k0h = j * ATT_BN + half * 64;
        int nvalid = p.T - k0h;
        nvalid = nvalid < 0 ? 0 : (nvalid > 64 ? 64 : nvalid);
        u64 u[32];
        float segb[2];
        float m_blk = scores_to_logits<KW, BT>(r, u, segb, k0h, bh, bw, has_bias && !(p.dbg & 2), p.scale_log2, p.Kh, p.Kw);
        if (nvalid < 64) {  // ragged last key block: keys >= T do not exist
          m_blk = -INFINITY;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            float a0, a1;
            unpack2(u[i], a0, a1);
            a0 = 2 * i < nvalid ? a0 : -INFINITY;
            a1 = 2 * i + 1 < nvalid ? a1 : -INFINITY;
            u[i] = pack2(a0, a1);
            m_blk = fmaxf(m_blk, fmaxf(a0, a1) + segb[i >> 4]);
          }
        }
        // lazy rescale: keep the stale reference maximum unless the block exceeds it by more than 2^8
        float alpha = 1.f;
        if (m_blk > m_ref + kRescaleThreshold) {
          alpha = ex2(m_ref - m_blk);      // 0 on the first block (m_ref = -inf)
          m_ref = m_blk;
        }
        const float m_use = m_ref == -INFINITY ? 0.f : m_ref;
        const float n0 = segb[0] - m_use, n1 = segb[1] - m_use;
        const u64 neg[2] = {pack2(n0, n0), pack2(n1, n1)};
        u64 ls2[2] = {0ull, 0ull};
        uint32_t pk[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float a0, a1;
          unpack2(add2(u[i], neg[i >> 4]), a0, a1);
          const float p0 = ex2(a0), p1 = ex2(a1);
          ls2[i & 1] = add2(ls2[i & 1], pack2(p0, p1));
          pk[i] = pack_bf16(p0, p1);
        }
        float ls[4];
        unpack2(ls2[0], ls[0], ls[1]);
        unpack2(ls2[1], ls[2], ls[3]);
        l_run = l_run * alpha + ((ls[0] + ls[1]) + (ls[2] + ls[3]));
        // ---- 2. P_{j-1} V_{j-1} must be complete before P's shared-memory tile or the accumulator is touched ----
        if (j > 0) {
          mbar_wait(pv_done, pvph);
          pvph ^= 1;
          tc_fence_after();
          if (__any_sync(0xffffffffu, alpha != 1.f)) {
            // rare: rescale this half's accumulator rows (alpha = 1 for the rows that kept their maximum)
#pragma unroll
            for (int cc = 0; cc < ATT_D; cc += 32) {
              uint32_t o[32];
              tmem_ld_32x32b_x32(lane_addr + TM_O + half * ATT_D + cc, o);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
              tmem_st_32x32b_x32(lane_addr + TM_O + half * ATT_D + cc, o);
            }
            tmem_st_wait();
          }
          tc_fence_before();
        }
        // ---- 3. P_j -> this half's swizzled K-major bf16 panel, hand it to the MMA warp ----
#pragma unroll
        for (int c8 = 0; c8 < 8; ++c8) {
          const uint4 v = make_uint4(pk[4 * c8], pk[4 * c8 + 1], pk[4 * c8 + 2], pk[4 * c8 + 3]);
          *reinterpret_cast<uint4*>(sP + half * TILE_BYTES + row * 128 + ((c8 ^ (row & 7)) << 4)) = v;
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full);
      }
      // ---- tile epilogue: merge the two halves, normalise, store; this thread writes 32 of the 64 dims ----
      {
        mbar_wait(pv_done, pvph);
        pvph ^= 1;
        tc_fence_after();
        sML[(half * ATT_BM + row) * 2] = m_ref;
        sML[(half * ATT_BM + row) * 2 + 1] = l_run;
        asm volatile("bar.sync %0, 64;" ::"r"(1 + quad) : "memory");
        const float m_o = sML[((half ^ 1) * ATT_BM + row) * 2], l_o = sML[((half ^ 1) * ATT_BM + row) * 2 + 1];
        const float m = fmaxf(m_ref, m_o);
        const float s_self = ex2(m_ref - m), s_oth = ex2(m_o - m);   // exp2(-inf) = 0 for a half without keys
        const float inv = 1.f / (l_run * s_self + l_o * s_oth);
        const float sa = (half == 0 ? s_self : s_oth) * inv, sb = (half == 0 ? s_oth : s_self) * inv;
        uint32_t oa[32], ob[32];
        __syncwarp();
        tmem_ld_32x32b_x32(lane_addr + TM_O + half * 32, oa);
        tmem_ld_32x32b_x32(lane_addr + TM_O + ATT_D + half * 32, ob);
        tmem_ld_wait();
        tc_fence_before();
        float o[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] = sa * __uint_as_float(oa[i]) + sb * __uint_as_float(ob[i]);
        long long orow = (long long)bp * p.T + q;
        if (q < p.T && p.out_map) orow = p.out_map[orow];
        if (q < p.T && orow >= 0) {
          uint4* op = reinterpret_cast<uint4*>(p.out + orow * (p.nh * ATT_D) + head * ATT_D + half * 32);
#pragma unroll
          for (int i = 0; i < 4; ++i) op[i] = pack8(o + 8 * i);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

}  // namespace mmsam

int mmsam_attention_glb(const void* qkv, void* out, const int* out_row_map_dev, const void* tab_h, const void* tab_w, int Bp, int T,
                        int nh, int Kh, float scale, int max_ctas, cudaStream_t stream);   // attention_glb.cu
int mmsam_attention_win(const void* qkv, void* out, const int* out_row_map_dev, const void* tab_h, const void* tab_w, int Bp,
                        int nh, float scale, int max_ctas, cudaStream_t stream);   // attention_win.cu

// See include/mmsam_b200.h for the contract.
MMSAM_API int mmsam_attention_bf16(const void* qkv, void* out, const int* out_row_map_dev, const void* tab_h,
                                   const void* tab_w, int Bp, int T, int nh, int Kh, int Kw, float scale,
                                   int max_ctas, void* stream) {
  using namespace mmsam;
  if (Bp < 0 || T <= 0 || nh <= 0) return MMSAM_ERR_BAD_ARG;
  if (Bp == 0) return MMSAM_OK;
  if (!qkv || !out) return MMSAM_ERR_BAD_ARG;
  if ((((uintptr_t)qkv | (uintptr_t)out) & 15)) return MMSAM_ERR_BAD_ARG;
  const bool has_bias = tab_h != nullptr && tab_w != nullptr;
  if ((tab_h != nullptr) != (tab_w != nullptr)) return MMSAM_ERR_BAD_ARG;
  if (has_bias && (Kh <= 0 || Kw <= 0 || Kh * Kw != T)) return MMSAM_ERR_BAD_ARG;
  // SAM's 14 x 14 windows: the single-pass kernel (attention_win.cu)
  static const int use_win = [] { const char* e = getenv("MMSAM_ATT_WIN"); return e ? atoi(e) : 1; }();
  if (use_win && T == 196 && (!has_bias || (Kh == 14 && Kw == 14))) {
    if (has_bias && (((uintptr_t)tab_h | (uintptr_t)tab_w) & 15)) return MMSAM_ERR_BAD_ARG;
    return mmsam_attention_win(qkv, out, out_row_map_dev, tab_h, tab_w, Bp, nh, scale, max_ctas, (cudaStream_t)stream);
  }
  // 64-wide token grids (ViT-L global blocks at 1024^2): the two-tile ping-pong kernel (attention_glb.cu)
  static const int use_glb = [] { const char* e = getenv("MMSAM_ATT_GLB"); return e ? atoi(e) : 1; }();
  if (use_glb && has_bias && Kw == 64 && Kh >= 2 && Kh <= 64 && (Kh & 1) == 0) {
    if ((((uintptr_t)tab_h | (uintptr_t)tab_w) & 15)) return MMSAM_ERR_BAD_ARG;
    return mmsam_attention_glb(qkv, out, out_row_map_dev, tab_h, tab_w, Bp, T, nh, Kh, scale, max_ctas, (cudaStream_t)stream);
  }
  if (!has_bias) { Kh = 1; Kw = T; }
  AttnParams p;
  p.out = (__nv_bfloat16*)out;
  p.out_map = out_row_map_dev;
  p.Bp = Bp; p.T = T; p.nh = nh; p.Kh = Kh; p.Kw = Kw;
  p.nh_rows = has_bias ? 2 * Kh - 1 : 0;
  p.nw_rows = has_bias ? 2 * Kw - 1 : 0;
  p.nh_pad = (p.nh_rows + 15) & ~15;
  p.nw_pad = (p.nw_rows + 15) & ~15;
  if (p.nh_pad > 256 || p.nw_pad > 256) return MMSAM_ERR_UNSUPPORTED;
  p.scale_log2 = scale * kLog2e;
  p.nqb = (T + ATT_BM - 1) / ATT_BM;
  p.num_tiles = Bp * nh * p.nqb;
  const int tab_bytes = ((p.nh_pad + p.nw_pad) * 128 + 1023) & ~1023;
  // per-row bias rows in shared memory: bh rows are read as scalars (odd stride), bw rows as 64-bit pairs (even
  // stride with an odd number of pairs: conflict-free LDS.64 across the 32 rows of a warp)
  p.bh_stride = has_bias ? (Kh | 1) : 1;
  p.bw_stride = 2;
  if (has_bias) {
    p.bw_stride = (Kw + 1) & ~1;
    if (((p.bw_stride >> 1) & 1) == 0) p.bw_stride += 2;
  }
  int bias_elt = 4;
  int bias_bytes = ATT_BM * (p.bh_stride + p.bw_stride) * bias_elt + 32;
  int fixed = TILE_BYTES /*Q*/ + 2 * TILE_BYTES /*P*/ + tab_bytes + bias_bytes + 2 * ATT_BM * 2 * 4 /*(m,l)*/ + 256 /*barriers*/ + 1024 /*align*/ + 64;
  const int budget = 227 * 1024;
  int kw_mode = 0;
  if (has_bias && Kw == 64 && T % ATT_BN == 0) kw_mode = 64;
  else if (has_bias && Kw == 32 && T % ATT_BN == 0) kw_mode = 32;
  else if (has_bias && Kw == 14 && Kh == 14) kw_mode = 14;
  if (kw_mode == 0 && fixed + 2 * 2 * TILE_BYTES > budget) {
    // large generic key grid: half-precision bias rows (|bias| is O(1): 2^-11 relative, below the bf16 rounding of P)
    bias_elt = 2;
    fixed -= bias_bytes;
    bias_bytes = ATT_BM * (p.bh_stride + p.bw_stride) * bias_elt + 32;
    fixed += bias_bytes;
  }
  p.kv_stages = 3;
  if (fixed + 3 * 2 * TILE_BYTES > budget) p.kv_stages = 2;
  int smem_bytes = fixed + p.kv_stages * 2 * TILE_BYTES;
  if (smem_bytes > budget) return MMSAM_ERR_UNSUPPORTED;
  p.q_bufs = 1;
  static const int dbg_flags = [] { const char* e = getenv("MMSAM_ATT_DBG"); return e ? atoi(e) : 0; }();
  p.dbg = dbg_flags;
  if (smem_bytes + TILE_BYTES <= budget) { p.q_bufs = 2; smem_bytes += TILE_BYTES; }
  if (!has_bias) { p.Kh = 1; p.Kw = 1 << 30; }

  mmsam_host::EncodeTiledFn enc = mmsam_host::get_encode_tiled();
  if (!enc) return MMSAM_ERR_DRIVER;
  CUtensorMap tmQKV, tmH, tmW;
  {
    const uint64_t C3 = (uint64_t)3 * nh * ATT_D;
    cuuint64_t dims[4] = {(cuuint64_t)ATT_D, (cuuint64_t)(3 * nh), (cuuint64_t)T, (cuuint64_t)Bp};
    cuuint64_t strides[3] = {(cuuint64_t)ATT_D * 2, C3 * 2, (cuuint64_t)T * C3 * 2};
    cuuint32_t box[4] = {ATT_D, 1, ATT_BM, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    if (enc(&tmQKV, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(qkv), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return MMSAM_ERR_DRIVER;
  }
  if (has_bias) {
    if ((((uintptr_t)tab_h | (uintptr_t)tab_w) & 15)) return MMSAM_ERR_BAD_ARG;
    int rc = mmsam_host::make_tmap_2d_bf16(&tmH, tab_h, p.nh_pad, ATT_D, ATT_D, p.nh_pad, ATT_D);
    if (rc) return rc;
    rc = mmsam_host::make_tmap_2d_bf16(&tmW, tab_w, p.nw_pad, ATT_D, ATT_D, p.nw_pad, ATT_D);
    if (rc) return rc;
  } else {
    tmH = tmQKV; tmW = tmQKV;
  }
  if (max_ctas <= 0 || max_ctas > kNumSMs) max_ctas = kNumSMs;
  const int grid = p.num_tiles < max_ctas ? p.num_tiles : max_ctas;
#define ATT_LAUNCH(KWM, BTY)                                                                                         \
  do {                                                                                                          \
    MMSAM_SET_SMEM_ONCE((attention_kernel<KWM, BTY>), budget);                                                   \
    attention_kernel<KWM, BTY><<<grid, 320, smem_bytes, (cudaStream_t)stream>>>(tmQKV, tmH, tmW, p);                 \
  } while (0)
  if (kw_mode == 64) ATT_LAUNCH(64, float);
  else if (kw_mode == 32) ATT_LAUNCH(32, float);
  else if (kw_mode == 14) ATT_LAUNCH(14, float);
  else if (bias_elt == 4) ATT_LAUNCH(0, float);
  else ATT_LAUNCH(0, __half);
#undef ATT_LAUNCH
  MMSAM_LAUNCH_CHECK();
  return MMSAM_OK;
}
