// SAM global attention (64-wide token grids: 64 x 64 = 4096 tokens at 1024^2, head_dim 64) with the decomposed
// relative-position bias on tcgen05 tensor cores: TWO 128-query tiles per CTA in ping-pong, P kept in tensor memory.
// Replaces Attention.forward + add_decomposed_rel_pos (base/image_encoder.py:483-501, 587-623) for the 4 global blocks of
// ViT-L; the general flash kernel (attention.cu) keeps every other key grid.
//
// The general kernel measured 1.46 ms per launch (8 x 16 heads x 4096^2): 2880 cycles per 128 x 128 score tile against a
// 1024-cycle MUFU.EX2 floor, because (a) its 8 softmax warps work on ONE tile in lock-step, so their barrier round trips,
// TMEM loads and shared-memory hand-overs of P add up instead of overlapping, (b) the key-column bias is a 64-bit
// shared-memory load per score pair and the 67 KB of per-row bias rows leave room for only two K/V stages, (c) the online
// softmax is two dependent passes (row maximum, then exponentials) per key block. Here:
//   item = (batch, head, 256 query rows = 4 token rows): Q tiles A and B, K / V streamed once for both through 3 stages.
//   warps 0-3 / 4-7  softmax group of tile A / B, thread = query row, ALL 128 keys of the block (= 2 key rows x 64
//            key columns). The row's 64 key-column biases live in REGISTERS (packed pairs); the 2 key-row biases of a
//            block are two scalar shared-memory reads. One pass: p = exp2(s * scale + bw + bh - m_ref) with a STALE
//            reference maximum (exact maximum of the first block; later blocks only track their maximum and, should it
//            exceed m_ref by more than 2^16, rescale P, the running sum and O by an exact power of two - a path real
//            attention rows do not take). P (bf16) goes to 64 scratch columns of TENSOR MEMORY and O += P V reads it
//            from there (tcgen05.mma TS form) - no shared-memory panel, no proxy fence. Because P does not overwrite S,
//            the group releases S as soon as the block is in registers (s_free) and the tensor core computes the NEXT
//            block's scores while this block's exponentials run: the softmax warps wait ~250 cycles per block
//            (clock64 trace, tools/attn_glb_trace.py) instead of 1300 when P aliased S.
//   warp 9   MMA issuer (whole warp, one elected lane): per block and group  s_free -> Q.K^T(j+1),  p_full -> P.V(j).
//   warp 8   TMA producer.
// Bias set-up per item: G = Q . table^T for both axes and both tiles (4 MMAs of N = 128 into the idle S / O / scratch
// columns), scattered per row through the warp's 8 KB slice of shared memory: first the key-column values (Toeplitz
// gather, then into registers), then the key-row values, which stay there as bh[key row][lane].
// Measured (8 x 16 heads x 4096^2): 0.79 ms per launch = 693 TFLOP/s; the two groups' exponentials keep the MUFU units
// ~70 % busy (2048 of ~2900 cycles per key block), which is what bounds it now.
// TMEM: S_A [0,128), S_B [128,256), O_A [256,320), O_B [320,384), P_A [384,448), P_B [448,512) (= the bias scratch).
#include "common.cuh"
#include <cstdlib>

namespace mmsam {

namespace glb {
constexpr int BM = 128, BN = 128, D = 64, KW = 64;
constexpr int TILE = 128 * 64 * 2;
constexpr int NST = 3;
constexpr int TAB_BYTES = 2 * 128 * 128;                 // both tables, 128 rows each (127 used + a zero row)
constexpr int BH_BYTES = 8 * 64 * 32 * 4;                // per softmax warp: [64][32 lanes] fp32
constexpr int SMEM_BYTES = 2 * TILE + NST * 2 * TILE + TAB_BYTES + BH_BYTES + 256 + 1024;
constexpr int TM_S = 0, TM_O = 256, TM_X = 384;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kThresh = 16.f;
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory");
}  // namespace glb

struct GlbParams {
  __nv_bfloat16* out;
  const int* out_map;
  int Bp, T, nh, Kh;
  int nqp;             // 256-row query pairs per (batch, head)
  int num_items;
  float scale_log2;
  long long* trace;    // perf debug (MMSAM_ATT_TRACE): clock64 stamps of CTA 0, first item: [role 0..2][block < 64][event < 8]
};
#define GLB_TRACE(role, blk, ev)                                                                              \
  do {                                                                                                       \
    if (p.trace && blockIdx.x == 0 && it == 0 && (blk) < 64) p.trace[((role) * 64 + (blk)) * 8 + (ev)] = clock64(); \
  } while (0)

__device__ __forceinline__ float glb_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// One 32-column chunk of the score row. MAXONLY: running maximum of u = s * scale + bw (the key-row bias is added by the
// caller); else: p = exp2(u + nb), row sum, bf16 P written to the P columns of the chunk, maximum of u tracked on the side.
// The exponentials are issued as ONE block of 32 MUFU.EX2 (volatile, in place) between the arithmetic that feeds them and
// the sums / packs that consume them: with the consumers right behind each pair (what the compiler schedules by itself)
// every FADD2 / F2FP waits out the MUFU latency, and a softmax warp has only one other warp on its scheduler to hide it.
// exp2 on the FMA pipe for a pair of values <= ~16: Cody-Waite split x = n + f (n = round(x) through the 1.5 * 2^23 magic
// add, f in [-0.5, 0.5]), degree-3 minimax polynomial for 2^f (relative error 7.5e-5: P is rounded to bf16, 4e-3, right
// after), 2^n added into the exponent field. ~6 instructions per element on the FMA / ALU pipes against one MUFU.EX2 that
// occupies the 4-lane special-function unit for 8 cycles per warp. GLB_POLY_MASK (one bit per pair of a 32-column chunk)
// moves a fraction of every chunk's exponentials off that unit (the split FlashAttention-4 uses on this chip).
// MEASURED: 25 % of the exponentials through the polynomial (mask 0x8888) = 0.807 ms per launch against 0.78 - 0.79 with
// all of them on the MUFU (and 24 bytes of spills): with two softmax warps per scheduler the loop is bound by the
// latency of its dependent chains, not by the throughput of either pipe, and the polynomial adds a 9-deep chain.
// Default 0 = off; kept for a structure with more resident softmax warps.
#ifndef GLB_POLY_MASK
#define GLB_POLY_MASK 0x0000u
#endif
__device__ __forceinline__ void glb_ex2_poly2(float& x0, float& x1) {
  const u64 x2 = pack2(fmaxf(x0, -126.f), fmaxf(x1, -126.f));
  const u64 t = add2(x2, pack2(12582912.f, 12582912.f));
  const u64 n = add2(t, pack2(-12582912.f, -12582912.f));
  const u64 f = fma2(n, pack2(-1.f, -1.f), x2);
  u64 q = fma2(f, pack2(0.05517164617776871f, 0.05517164617776871f), pack2(0.2426111251115799f, 0.2426111251115799f));
  q = fma2(q, f, pack2(0.6932609677314758f, 0.6932609677314758f));
  q = fma2(q, f, pack2(0.9999280571937561f, 0.9999280571937561f));
  float q0, q1, t0, t1;
  unpack2(q, q0, q1);
  unpack2(t, t0, t1);
  x0 = __int_as_float(__float_as_int(q0) + (__float_as_int(t0) << 23));
  x1 = __int_as_float(__float_as_int(q1) + (__float_as_int(t1) << 23));
}

template <bool MAXONLY>
__device__ __forceinline__ void glb_chunk(const uint32_t (&r)[32], const u64* bw2, u64 sc2, u64 nb2, float& mx, u64 (&sum2)[2],
                                          uint32_t (&pk)[16]) {
  float e[32];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const u64 u = fma2(pack2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1])), sc2, bw2[i]);
    float a, b;
    unpack2(u, a, b);
    mx = max3(mx, a, b);
    if constexpr (!MAXONLY) unpack2(add2(u, nb2), e[2 * i], e[2 * i + 1]);
  }
  if constexpr (!MAXONLY) {
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (!((GLB_POLY_MASK >> (i >> 1)) & 1u)) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(e[i]));
#pragma unroll
    for (int i = 0; i < 16; ++i)
      if ((GLB_POLY_MASK >> i) & 1u) glb_ex2_poly2(e[2 * i], e[2 * i + 1]);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(sum2[i & 1]) : "l"(pack2(e[2 * i], e[2 * i + 1])));
      asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(pk[i]) : "f"(e[2 * i + 1]), "f"(e[2 * i]));
    }
  }
}

__global__ void __launch_bounds__(320, 1)
attention_glb_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmTabH,
                     const __grid_constant__ CUtensorMap tmTabW, const GlbParams p) {
  using namespace glb;
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sQ = smem;                               // tiles A, B
  uint8_t* sKV = sQ + 2 * TILE;                     // NST x (K, V)
  uint8_t* sTab = sKV + NST * 2 * TILE;             // Rh | Rw
  float* sBh = reinterpret_cast<float*>(sTab + TAB_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sBh) + BH_BYTES);
  uint64_t* q_full = bars + 0;
  uint64_t* q_empty = bars + 1;
  uint64_t* tab_full = bars + 2;
  uint64_t* g_full = bars + 3;      // bias pre-products are in tensor memory
  uint64_t* g_done = bars + 4;      // 8 warps: scattered
  uint64_t* o_free = bars + 5;      // 8 warps: the previous item's O has been read
  uint64_t* s_full = bars + 6;      // [2]
  uint64_t* p_full = bars + 8;      // [2] 4 warps each
  uint64_t* pv_done = bars + 10;    // [2]
  uint64_t* kv_full = bars + 12;    // [NST]
  uint64_t* kv_empty = bars + 15;   // [NST]
  uint64_t* s_free = bars + 18;     // [2] 4 warps each: the group holds its whole score row block in registers
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkb = p.T / BN;
  const int n_my = ((int)blockIdx.x < p.num_items) ? (p.num_items - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    mbar_init(q_full, 1); mbar_init(q_empty, 1); mbar_init(tab_full, 1);
    mbar_init(g_full, 1); mbar_init(g_done, 8); mbar_init(o_free, 8);
    for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&p_full[i], 4); mbar_init(&pv_done[i], 1); mbar_init(&s_free[i], 4); }
    for (int i = 0; i < NST; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
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
      if (elect_one()) {
        mbar_arrive_expect_tx(tab_full, TAB_BYTES);
        tma_load_2d(sTab, &tmTabH, tab_full, 0, 0);
        tma_load_2d(sTab + 128 * 128, &tmTabW, tab_full, 0, 0);
      }
      __syncwarp();
      int st = 0; uint32_t kph = 0;
      for (int it = 0; it < n_my; ++it) {
        const int item = blockIdx.x + it * gridDim.x;
        const int qp = item % p.nqp, bh = item / p.nqp;
        const int head = bh % p.nh, bp = bh / p.nh;
        mbar_wait(q_empty, (it & 1) ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(q_full, 2 * TILE);
          tma_load_4d(sQ, &tmQKV, q_full, 0, head, qp * 256, bp);
          tma_load_4d(sQ + TILE, &tmQKV, q_full, 0, head, qp * 256 + 128, bp);        // rows >= T: zero-filled
        }
        __syncwarp();
        for (int j = 0; j < nkb; ++j) {
          mbar_wait(&kv_empty[st], kph ^ 1);
          if (elect_one()) {
            mbar_arrive_expect_tx(&kv_full[st], 2 * TILE);
            uint8_t* sk = sKV + st * 2 * TILE;
            tma_load_4d(sk, &tmQKV, &kv_full[st], 0, p.nh + head, j * BN, bp);
            tma_load_4d(sk + TILE, &tmQKV, &kv_full[st], 0, 2 * p.nh + head, j * BN, bp);
          }
          __syncwarp();
          if (++st == NST) { st = 0; kph ^= 1; }
        }
      }
    }
  } else if (warp == 9) {
    // =========================== MMA issuer: the whole warp runs the loop, one elected lane issues ===========================
    {
      constexpr uint32_t idesc_qk = umma_idesc_bf16(128, BN, 0, 0);
      constexpr uint32_t idesc_pv = umma_idesc_bf16(128, D, 0, 1);      // A = P in tensor memory, B = V MN-major
      mbar_wait(tab_full, 0);
      const uint32_t q_addr[2] = {smem_u32(sQ), smem_u32(sQ + TILE)};
      const uint32_t t_addr = smem_u32(sTab);
      int st = 0; uint32_t kph = 0;
      uint32_t pph[2] = {0, 0};
      uint32_t sfree_ph[2] = {0, 0};
      int trace_it = 0, trace_j = 0;
      auto issue_qk = [&](int g, uint32_t k_addr) {
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16_ss(tmem + TM_S + g * BN, umma_desc_sw128(q_addr[g] + k * 32), umma_desc_sw128(k_addr + k * 32), idesc_qk,
                        k != 0 ? 1u : 0u);
          umma_commit(&s_full[g]);
        }
        __syncwarp();
      };
      auto issue_pv = [&](int g, uint32_t v_addr, bool first) {
        mbar_wait(&p_full[g], pph[g]);
        pph[g] ^= 1;
        tc_fence_after();
        if (lane == 0 && p.trace && blockIdx.x == 0 && trace_it == 0 && trace_j < 64) p.trace[(2 * 64 + trace_j) * 8 + g * 2] = clock64();
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < BN / 16; ++kk)
            umma_f16_ts(tmem + TM_O + g * D, tmem + TM_X + g * 64 + kk * 8, umma_desc_sw128(v_addr + kk * 16 * 128), idesc_pv,
                        (first && kk == 0) ? 0u : 1u);
          umma_commit(&pv_done[g]);
        }
        __syncwarp();
      };
      for (int it = 0; it < n_my; ++it) {
        // ---- bias pre-products of both tiles and both axes: Gw(A) -> S_A, Gh(A) -> scratch, Gw(B) -> S_B, Gh(B) -> O ----
        mbar_wait(q_full, it & 1);
        mbar_wait(o_free, (it & 1) ^ 1);
        tc_fence_after();
        const uint32_t gdst[4] = {tmem + TM_S, tmem + TM_X, tmem + TM_S + BN, tmem + TM_O};
        if (elect_one()) {
#pragma unroll
          for (int gi = 0; gi < 4; ++gi) {
            const uint32_t tab = t_addr + ((gi & 1) ? 0 : 128 * 128);        // even: key-column table (Rw), odd: key-row table (Rh)
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16_ss(gdst[gi], umma_desc_sw128(q_addr[gi >> 1] + k * 32), umma_desc_sw128(tab + k * 32), idesc_qk, k != 0 ? 1u : 0u);
          }
          umma_commit(g_full);
        }
        __syncwarp();
        mbar_wait(g_done, it & 1);
        tc_fence_after();
        // ---- key blocks. P lives in the scratch columns, not over S: Q.K^T of block j + 1 is issued as soon as the group
        //      holds block j in registers (s_free), long before its exponentials are done - the softmax warps never wait for the
        //      tensor core in steady state; P.V of block j follows when P is written. ----
        trace_it = it;
        uint32_t sfph[2] = {sfree_ph[0], sfree_ph[1]};
        {
          mbar_wait(&kv_full[st], kph);
          const uint32_t k_addr = smem_u32(sKV + st * 2 * TILE);
          issue_qk(0, k_addr);
          issue_qk(1, k_addr);
          if (nkb == 1) {
            if (elect_one()) umma_commit(q_empty);
            __syncwarp();
          }
        }
        for (int j = 0; j < nkb; ++j) {
          trace_j = j;
          const int st_next = st + 1 == NST ? 0 : st + 1;
          const uint32_t kph_next = st + 1 == NST ? kph ^ 1 : kph;
          const bool more = j + 1 < nkb;
          if (lane == 0) GLB_TRACE(2, j, 4);
          if (more) mbar_wait(&kv_full[st_next], kph_next);
          if (lane == 0) GLB_TRACE(2, j, 5);
          const uint32_t k_next = smem_u32(sKV + st_next * 2 * TILE);
          const uint32_t v_addr = smem_u32(sKV + st * 2 * TILE + TILE);
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            mbar_wait(&s_free[g], sfph[g]);
            sfph[g] ^= 1;
            if (more) {
              tc_fence_after();
              issue_qk(g, k_next);
            }
            if (lane == 0) GLB_TRACE(2, j, 1 + 2 * g);
            issue_pv(g, v_addr, j == 0);
          }
          if (elect_one()) {
            umma_commit(&kv_empty[st]);
            if (j + 2 == nkb) umma_commit(q_empty);       // every MMA that reads this item's Q has been issued
          }
          __syncwarp();
          st = st_next; kph = kph_next;
        }
        sfree_ph[0] = sfph[0]; sfree_ph[1] = sfph[1];
      }
    }
  } else {
    // =========================== softmax groups (warps 0-3: tile A, warps 4-7: tile B) ===========================
    const int g = warp >> 2, wq = warp & 3;
    const int row = wq * 32 + lane;                      // TMEM lane == query row inside the tile
    const uint32_t lane_addr = tmem + ((uint32_t)(wq * 32) << 16);
    const uint32_t s_addr = lane_addr + TM_S + g * BN;
    const uint32_t o_addr = lane_addr + TM_O + g * D;
    const uint32_t p_addr = lane_addr + TM_X + g * 64;      // P (bf16 pairs): 64 scratch columns per group
    float* slice = sBh + warp * (64 * 32) + lane;        // this warp's [64][lane] area, this lane's column
    const u64 sc2 = pack2(p.scale_log2, p.scale_log2);
    const int qw = (wq & 1) * 32 + lane;                 // token column of this row (tile rows = 2 token rows x 64)
    uint32_t sph = 0;                                    // s_full parity
    uint32_t npv = 0;                                    // P.V products of this group committed so far (pv_done phases)
    for (int it = 0; it < n_my; ++it) {
      const int item = blockIdx.x + it * gridDim.x;
      const int qp = item % p.nqp, bhid = item / p.nqp;
      const int head = bhid % p.nh, bp = bhid / p.nh;
      const int q = qp * 256 + g * 128 + row;
      const int qh = qp * 4 + g * 2 + (wq >> 1);         // token row (warp-uniform)
      // ---- bias set-up ----
      mbar_wait(g_full, it & 1);
      tc_fence_after();
      u64 bw2[32];
      {
        const uint32_t gw_addr = s_addr;                                             // Gw of this tile sits in its S columns
        const uint32_t gh_addr = lane_addr + (g == 0 ? TM_X : TM_O);
        // key-column values: bw[kw] = Gw[qw - kw + 63]  (image_encoder.py:609-621), i.e. table row c holds kw = qw + 63 - c
#pragma unroll 1
        for (int c0 = 0; c0 < 128; c0 += 32) {
          uint32_t r[32];
          tmem_ld_32x32b_x32(gw_addr + c0, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int kw = qw + 63 - (c0 + i);
            if (kw >= 0 && kw < 64) slice[kw * 32] = __uint_as_float(r[i]) * kLog2e;
          }
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 32; ++i) bw2[i] = pack2(slice[(2 * i) * 32], slice[(2 * i + 1) * 32]);
        __syncwarp();
        // key-row values: bh[kh] = Gh[qh - kh + Kh - 1], kept in the slice as [kh][lane]
#pragma unroll 1
        for (int c0 = 0; c0 < 128; c0 += 32) {
          uint32_t r[32];
          tmem_ld_32x32b_x32(gh_addr + c0, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int kh = qh + p.Kh - 1 - (c0 + i);
            if (kh >= 0 && kh < p.Kh) slice[kh * 32] = __uint_as_float(r[i]) * kLog2e;
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(g_done);
      }
      float m_ref = 0.f, l_run = 0.f;
      for (int j = 0; j < nkb; ++j) {
        if (wq == 0 && lane == 0) GLB_TRACE(g, j, 0);
        mbar_wait(&s_full[g], sph);
        sph ^= 1;
        tc_fence_after();
        if (wq == 0 && lane == 0) GLB_TRACE(g, j, 1);
        const float bh0 = slice[(2 * j) * 32], bh1 = slice[(2 * j + 1) * 32];
        uint32_t ra[32], rb[32], pka[16], pkb[16];
        u64 sum2[2] = {pack2(0.f, 0.f), pack2(0.f, 0.f)};
        float mx0 = -INFINITY, mx1 = -INFINITY;
        if (j == 0) {
          // exact row maximum of the first block (no reference yet)
          tmem_ld_32x32b_x32(s_addr, ra);
          tmem_ld_wait();
          tmem_ld_32x32b_x32(s_addr + 32, rb);
          glb_chunk<true>(ra, bw2, sc2, 0ull, mx0, sum2, pka);
          tmem_ld_wait();
          tmem_ld_32x32b_x32(s_addr + 64, ra);
          glb_chunk<true>(rb, bw2 + 16, sc2, 0ull, mx0, sum2, pka);
          tmem_ld_wait();
          tmem_ld_32x32b_x32(s_addr + 96, rb);
          glb_chunk<true>(ra, bw2, sc2, 0ull, mx1, sum2, pka);
          tmem_ld_wait();
          glb_chunk<true>(rb, bw2 + 16, sc2, 0ull, mx1, sum2, pka);
          m_ref = fmaxf(mx0 + bh0, mx1 + bh1);
          mx0 = mx1 = -INFINITY;
        }
        const float n0 = bh0 - m_ref, n1 = bh1 - m_ref;
        const u64 nb0 = pack2(n0, n0), nb1 = pack2(n1, n1);
        tmem_ld_32x32b_x32(s_addr, ra);
        tmem_ld_wait();
        tmem_ld_32x32b_x32(s_addr + 32, rb);
        glb_chunk<false>(ra, bw2, sc2, nb0, mx0, sum2, pka);
        tmem_ld_wait();
        tmem_ld_32x32b_x32(s_addr + 64, ra);
        glb_chunk<false>(rb, bw2 + 16, sc2, nb0, mx0, sum2, pkb);
        if (j > 0) {
          // P of the previous block must have been consumed before it is overwritten; two chunks into this block its P.V
          // (issued when that P was handed over) has normally completed
          if (wq == 0 && lane == 0) GLB_TRACE(g, j, 4);
          mbar_wait(&pv_done[g], (npv - 1) & 1);
          tc_fence_after();
          if (wq == 0 && lane == 0) GLB_TRACE(g, j, 5);
        }
        tmem_st_32x32b_x16(p_addr, pka);
        tmem_st_32x32b_x16(p_addr + 16, pkb);
        tmem_ld_wait();
        tmem_ld_32x32b_x32(s_addr + 96, rb);
        glb_chunk<false>(ra, bw2, sc2, nb1, mx1, sum2, pka);
        tmem_st_32x32b_x16(p_addr + 32, pka);
        tmem_ld_wait();
        // the whole block is in registers: the tensor core may overwrite S with the next block's scores
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_free[g]);
        glb_chunk<false>(rb, bw2 + 16, sc2, nb1, mx1, sum2, pkb);
        tmem_st_32x32b_x16(p_addr + 48, pkb);
        float ls0, ls1, ls2, ls3;
        unpack2(sum2[0], ls0, ls1);
        unpack2(sum2[1], ls2, ls3);
        float lsum = (ls0 + ls1) + (ls2 + ls3);
        const float m_blk = fmaxf(mx0 + bh0, mx1 + bh1);
        if (__any_sync(0xffffffffu, m_blk > m_ref + kThresh)) {
          // rare: this block's logits exceed the reference by more than 2^16. Rescale by an exact power of two: P of this block,
          // its sum, the running sum and the accumulator (rows that are fine use a factor of 1).
          const float delta = m_blk > m_ref + kThresh ? ceilf(m_blk - m_ref) : 0.f;
          const float f = glb_ex2(-delta);
          tmem_st_wait();
#pragma unroll 1
          for (int c0 = 0; c0 < 64; c0 += 32) {
            uint32_t pr[32];
            tmem_ld_32x32b_x32(p_addr + c0, pr);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) pr[i] = pack_bf16(bf16lo(pr[i]) * f, bf16hi(pr[i]) * f);
            tmem_st_32x32b_x32(p_addr + c0, pr);
          }
          lsum *= f;
          l_run *= f;
          m_ref += delta;
          if (j > 0) {
            mbar_wait(&pv_done[g], (npv - 1) & 1);       // P.V of the previous block has left the accumulator alone
            tc_fence_after();
#pragma unroll 1
            for (int c0 = 0; c0 < D; c0 += 32) {
              uint32_t o[32];
              tmem_ld_32x32b_x32(o_addr + c0, o);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * f);
              tmem_st_32x32b_x32(o_addr + c0, o);
            }
          }
        }
        l_run += lsum;
        if (wq == 0 && lane == 0) GLB_TRACE(g, j, 2);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[g]);
        if (wq == 0 && lane == 0) GLB_TRACE(g, j, 3);
        ++npv;
      }
      // ---- item epilogue: O / l -> bf16 -> global ----
      mbar_wait(&pv_done[g], (npv - 1) & 1);
      tc_fence_after();
      {
        const float inv = 1.f / l_run;
        uint32_t oa[32], ob[32];
        tmem_ld_32x32b_x32(o_addr, oa);
        tmem_ld_32x32b_x32(o_addr + 32, ob);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(o_free);
        long long orow = (long long)bp * p.T + q;
        if (q < p.T && p.out_map) orow = p.out_map[orow];
        if (q < p.T && orow >= 0) {
          uint4* op = reinterpret_cast<uint4*>(p.out + orow * (p.nh * D) + head * D);
          float o[8];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
#pragma unroll
            for (int k = 0; k < 8; ++k) o[k] = __uint_as_float(oa[8 * i + k]) * inv;
            op[i] = pack8(o);
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
#pragma unroll
            for (int k = 0; k < 8; ++k) o[k] = __uint_as_float(ob[8 * i + k]) * inv;
            op[4 + i] = pack8(o);
          }
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

static long long* g_glb_trace_buf = nullptr;
// perf debug: copy the clock64 trace of the last traced launch to the host (3 x 64 x 8 values); returns 0 when tracing is off
extern "C" __attribute__((visibility("default"))) int mmsam_dbg_attn_glb_trace(long long* host_out) {
  if (!g_glb_trace_buf) return 0;
  cudaDeviceSynchronize();
  cudaMemcpy(host_out, g_glb_trace_buf, 3 * 64 * 8 * sizeof(long long), cudaMemcpyDeviceToHost);
  return 1;
}

// qkv bf16 [Bp, T, 3, nh, 64] -> out; T = Kh * 64 tokens, Kh even and <= 64, both tables zero-padded to 128 rows.
int mmsam_attention_glb(const void* qkv, void* out, const int* out_row_map_dev, const void* tab_h, const void* tab_w, int Bp, int T,
                        int nh, int Kh, float scale, int max_ctas, cudaStream_t stream) {
  using namespace mmsam;
  GlbParams p;
  p.out = (__nv_bfloat16*)out;
  p.out_map = out_row_map_dev;
  p.Bp = Bp; p.T = T; p.nh = nh; p.Kh = Kh;
  p.nqp = (T + 255) / 256;
  p.num_items = Bp * nh * p.nqp;
  p.scale_log2 = scale * glb::kLog2e;
  static const int want_trace = getenv("MMSAM_ATT_TRACE") != nullptr;
  if (want_trace && !g_glb_trace_buf) {
    cudaMalloc(&g_glb_trace_buf, 3 * 64 * 8 * sizeof(long long));
    cudaMemset(g_glb_trace_buf, 0, 3 * 64 * 8 * sizeof(long long));
  }
  p.trace = want_trace ? g_glb_trace_buf : nullptr;
  mmsam_host::EncodeTiledFn enc = mmsam_host::get_encode_tiled();
  if (!enc) return MMSAM_ERR_DRIVER;
  CUtensorMap tmQKV, tmH, tmW;
  {
    const uint64_t C3 = (uint64_t)3 * nh * glb::D;
    cuuint64_t dims[4] = {(cuuint64_t)glb::D, (cuuint64_t)(3 * nh), (cuuint64_t)T, (cuuint64_t)Bp};
    cuuint64_t strides[3] = {(cuuint64_t)glb::D * 2, C3 * 2, (cuuint64_t)T * C3 * 2};
    cuuint32_t box[4] = {glb::D, 1, 128, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    if (enc(&tmQKV, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(qkv), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return MMSAM_ERR_DRIVER;
  }
  // tables: [pad16(2K-1), 64] bf16, zero rows beyond 2K-1; boxes of 128 rows (rows past the buffer are zero-filled by TMA)
  int rc = mmsam_host::make_tmap_2d_bf16(&tmH, tab_h, (uint64_t)((2 * Kh - 1 + 15) & ~15), 64, 64, 128, 64);
  if (rc) return rc;
  rc = mmsam_host::make_tmap_2d_bf16(&tmW, tab_w, 128, 64, 64, 128, 64);
  if (rc) return rc;
  if (max_ctas <= 0 || max_ctas > kNumSMs) max_ctas = kNumSMs;
  const int grid = p.num_items < max_ctas ? p.num_items : max_ctas;
  MMSAM_SET_SMEM_ONCE((attention_glb_kernel), glb::SMEM_BYTES);
  cudaError_t le = mmsam_host::launch_pdl(attention_glb_kernel, dim3(grid), dim3(320), glb::SMEM_BYTES, stream, tmQKV, tmH, tmW, p);
  if (le != cudaSuccess) return (int)le;
  MMSAM_LAUNCH_CHECK();
  return MMSAM_OK;
}
