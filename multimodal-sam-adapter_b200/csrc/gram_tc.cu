// Gram matrices over pixels on the tensor cores:
//     S[b, i, j] += sum_pix X[b, pix, qoff + i] * X[b, pix, koff + j]        (+ squared column norms)
// AttentionBase q @ k^T over HW with F.normalize'd q / k (segmentation/mmseg_custom/models/backbones/
// adapter_modules_multimodal_mix_mod_new_in_twin_convnext_new.py:98-103) and GFFM's cross-modal energies
// (:250-254). X is a channels-last bf16 map [B, HW, ld], so both operands of the contraction are "MN-major"
// (channels contiguous, the reduction runs over rows = pixels): a TMA box {64 channels, 64 pixels} lands in shared
// memory exactly as one SWIZZLE_128B MN-major UMMA atom, M = 128 / N = 128|256 are 2 / 2|4 such atoms side by side
// (descriptor leading-byte-offset = atom size).
//
// The SIMT version this replaces (neck.cu: fp32 FMA, 64x64 tiles) ran at 0.3-0.6 ms per launch, 4.1 ms per step for
// maps that take 50 us to stream: 8.6-19 G FMA per launch do not belong on CUDA cores.
//
// CTA = (128 x BNJ output tile, pixel chunk, image): warp 4 TMA producer, warp 5 MMA issuer (fp32 accumulators in
// TMEM: G, and when norms are requested A^T A / B^T B whose diagonals are the squared norms), warps 0-3 epilogue
// (tcgen05.ld -> plain stores of this pixel chunk's PARTIAL sums: S_part [chunk][b][n][n], nq/nk_part [chunk][b][n];
// the consumers (neck.cu: gfe_weff_kernel / gffm_softmax_kernel) add the chunks in a fixed order, so the result does
// not depend on the CTA schedule — no floating-point atomics anywhere). Split over pixel chunks so that every SM has
// work; HBM-bound: X is read from DRAM once, tiles of the same chunk share it through L2.
#include "common.cuh"
#include <cstdlib>

namespace mmsam {

static constexpr int GR_KP = 64;                 // pixels per stage
static constexpr int GR_ATOM = GR_KP * 128;      // bytes of one {64 ch, 64 px} atom
static constexpr int GR_STAGES = 4;

__device__ __forceinline__ uint64_t umma_desc_sw128_mn(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3ffffu) >> 4);   // start address
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16;   // leading byte offset: between 64-element atoms along M/N
  d |= (uint64_t)(1024 >> 4) << 32;               // stride byte offset: between 8-row groups along K (pixels)
  d |= (uint64_t)1 << 46;                         // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                         // SWIZZLE_128B
  return d;
}

struct GramParams {
  float* S; float* nq; float* nk;
  int qoff, koff, n, HW, chunk, blk, nti, ntj;
};

template <int BNJ>
__global__ void __launch_bounds__(192, 1)
gram_tc_kernel(const __grid_constant__ CUtensorMap tmX, const GramParams p) {
  constexpr int NA = 2, NB = BNJ / 64;
  constexpr int STAGE_BYTES = (NA + NB) * GR_ATOM;
  constexpr int TM_G = 0, TM_NQ = 256, TM_NK = 384;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + GR_STAGES * STAGE_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + GR_STAGES;
  uint64_t* done = bars + 2 * GR_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * GR_STAGES + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ti = blockIdx.x / p.ntj, tj = blockIdx.x % p.ntj;
  const int i0 = ti * 128, j0 = tj * BNJ;
  if (p.blk > 0) {  // skip tiles that do not touch the block diagonal (uniform over the CTA, before any barrier)
    const int i1 = min(i0 + 127, p.n - 1), j1 = min(j0 + BNJ - 1, p.n - 1);
    if (i1 / p.blk < j0 / p.blk || j1 / p.blk < i0 / p.blk) return;
  }
  // squared norms of the q / k columns: on the diagonal tiles (never skipped by the block-diagonal test; BNJ == 128
  // whenever norms are requested, so tile (t, t) holds q columns and k columns [128 t, 128 t + 128))
  const bool want_nq = p.nq != nullptr && ti == tj;
  const bool want_nk = p.nk != nullptr && ti == tj;
  const int b = blockIdx.z;
  const int p0 = blockIdx.y * p.chunk, p1 = min(p0 + p.chunk, p.HW);
  const int nst = (p1 - p0) / GR_KP;

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tmX);
    for (int i = 0; i < GR_STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(done, 1);
    fence_barrier_init();
  }
  if (warp == 5) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 4) {
    {   // TMA producer: the whole warp runs the loop, one elected lane issues (see elect_one)
      int s = 0; uint32_t ph = 0;
      for (int it = 0; it < nst; ++it) {
        mbar_wait(&empty[s], ph ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&full[s], STAGE_BYTES);
          uint8_t* st = smem + s * STAGE_BYTES;
          const int prow = b * p.HW + p0 + it * GR_KP;
#pragma unroll
          for (int a = 0; a < NA; ++a) tma_load_2d(st + a * GR_ATOM, &tmX, &full[s], p.qoff + i0 + a * 64, prow);
#pragma unroll
          for (int a = 0; a < NB; ++a) tma_load_2d(st + (NA + a) * GR_ATOM, &tmX, &full[s], p.koff + j0 + a * 64, prow);
        }
        __syncwarp();
        if (++s == GR_STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 5) {
    {   // the whole warp runs the loop, one elected lane issues (see elect_one)
      constexpr uint32_t idesc_g = umma_idesc_bf16(128, BNJ, 1, 1);
      constexpr uint32_t idesc_n = umma_idesc_bf16(128, 128, 1, 1);
      int s = 0; uint32_t ph = 0;
      for (int it = 0; it < nst; ++it) {
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + s * STAGE_BYTES);
        const uint32_t b_addr = a_addr + NA * GR_ATOM;
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < GR_KP / 16; ++ks) {
            const uint64_t da = umma_desc_sw128_mn(a_addr + ks * 16 * 128, GR_ATOM);
            const uint64_t db = umma_desc_sw128_mn(b_addr + ks * 16 * 128, GR_ATOM);
            const uint32_t acc = (it | ks) != 0 ? 1u : 0u;
            umma_f16_ss(tmem + TM_G, da, db, idesc_g, acc);
            if (want_nq) umma_f16_ss(tmem + TM_NQ, da, da, idesc_n, acc);
            if (want_nk) umma_f16_ss(tmem + TM_NK, db, db, idesc_n, acc);
          }
          umma_commit(&empty[s]);
        }
        __syncwarp();
        if (++s == GR_STAGES) { s = 0; ph ^= 1; }
      }
      if (elect_one()) umma_commit(done);
      __syncwarp();
    }
  } else {
    // ---------------- epilogue: thread = output row i ----------------
    mbar_wait(done, 0);
    tc_fence_after();
    const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
    const int li = warp * 32 + lane;          // row inside the tile
    const int gi = i0 + li;
    const long long slot = (long long)blockIdx.y * gridDim.z + b;      // (chunk, image)
    float* Sb = p.S + slot * p.n * p.n;
    if (nst > 0) {
      // A thread holds one ROW of the tile: stored as is, every store instruction would touch 32 different lines
      // (17 us of sector writes for a 128 x 256 tile). The 32 x 32 block of a warp goes through shared memory (the
      // pipeline stages are free once `done` has fired; row pitch 36 floats: conflict-free both ways) and leaves as
      // 16-byte pieces, 8 lanes per row = full 128-byte lines.
      const bool vec4 = (p.blk & 3) == 0;
      const uint32_t tb = smem_u32(smem) + (uint32_t)warp * (32 * 36 * 4);
#pragma unroll 1
      for (int c = 0; c < BNJ; c += 32) {
        if (j0 + c >= p.n) break;
        uint32_t r[32];
        __syncwarp();
        tmem_ld_32x32b_x32(lane_addr + TM_G + c, r);
        tmem_ld_wait();
        if (vec4) {
#pragma unroll
          for (int k = 0; k < 8; ++k)
            asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(tb + (uint32_t)(lane * 36 + k * 4) * 4u), "r"(r[4 * k]),
                         "r"(r[4 * k + 1]), "r"(r[4 * k + 2]), "r"(r[4 * k + 3]) : "memory");
          __syncwarp();
          const int cg = j0 + c + (lane & 7) * 4;
#pragma unroll
          for (int rr = 0; rr < 8; ++rr) {
            const int row = rr * 4 + (lane >> 3);
            const int gr = i0 + warp * 32 + row;
            uint4 v;
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                         : "r"(tb + (uint32_t)(row * 36 + (lane & 7) * 4) * 4u));
            if (gr < p.n && cg < p.n && (p.blk == 0 || gr / p.blk == cg / p.blk))
              *reinterpret_cast<uint4*>(Sb + (long long)gr * p.n + cg) = v;
          }
        } else if (gi < p.n) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int gj = j0 + c + j;
            if (gj < p.n && (p.blk == 0 || gi / p.blk == gj / p.blk)) Sb[(long long)gi * p.n + gj] = __uint_as_float(r[j]);
          }
        }
      }
      if (want_nq) {   // diagonal of A^T A: element (li, li) sits in 32-column chunk `warp`, register `lane`
        uint32_t r[32];
        __syncwarp();
        tmem_ld_32x32b_x32(lane_addr + TM_NQ + warp * 32, r);
        tmem_ld_wait();
        float v = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) v = j == lane ? __uint_as_float(r[j]) : v;
        if (gi < p.n) p.nq[slot * p.n + gi] = v;
      }
      if (want_nk) {
        uint32_t r[32];
        __syncwarp();
        tmem_ld_32x32b_x32(lane_addr + TM_NK + warp * 32, r);
        tmem_ld_wait();
        float v = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) v = j == lane ? __uint_as_float(r[j]) : v;
        const int gj = j0 + li;
        if (gj < p.n) p.nk[slot * p.n + gj] = v;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

template <int BNJ>
static int launch_gram_tc(const CUtensorMap& tm, const GramParams& p, dim3 grid, cudaStream_t st) {
  constexpr int smem = GR_STAGES * (2 + BNJ / 64) * GR_ATOM + 1024 + 256;
  MMSAM_SET_SMEM_ONCE(gram_tc_kernel<BNJ>, smem);
  gram_tc_kernel<BNJ><<<grid, 192, smem, st>>>(tm, p);
  MMSAM_LAUNCH_CHECK();
  return MMSAM_OK;
}

}  // namespace mmsam

// Pixel chunking of the tensor-core path. It depends on the map size and the matrix size ONLY (not on the batch): an
// image's partial sums are then the same whatever batch it sits in, which makes the whole forward batch-invariant bit
// for bit. The chunk count is chosen so that (output tiles that run) x (chunks) ~ 18 CTAs per image (one wave of the
// 148 SMs at the bench batch of 8; measured against 36 and 72: tools/bench_gram.py): few chunks for the large matrices of the deep levels — every chunk costs a full
// [n, n] fp32 partial that the consumer has to read back — many for the 96-channel level whose single tile would
// otherwise leave the machine empty. At least 4 pipeline stages of 64 pixels per CTA. Returns the number of chunks,
// 0 when the shape does not fit this path.
int mmsam_gram_tc_plan(int n, int B, int HW, int norms, int* chunk_out) {
  using namespace mmsam;
  (void)B;
  if (HW % GR_KP != 0 || (long long)B * HW > 0x7fffffffLL) return 0;
  static const int target = [] { const char* e = getenv("MMSAM_GRAM_TARGET"); return e ? atoi(e) : 18; }();
  const int bnj = (norms || n <= 128) ? 128 : 256;
  const int nti = (n + 127) / 128, ntj = (n + bnj - 1) / bnj;
  int tiles = nti * ntj;
  if (norms && 3 * nti - 2 < tiles) tiles = 3 * nti - 2;      // per-head blocks: only tiles on the block diagonal run
  int nch = target / tiles;          // rounded down: tiles x chunks x 8 images stay within one wave of 148 CTAs
  if (nch < 1) nch = 1;
  int chunk = (HW + nch - 1) / nch;
  if (chunk < 4 * GR_KP) chunk = 4 * GR_KP;
  if (chunk > HW) chunk = HW;
  chunk = (chunk + GR_KP - 1) / GR_KP * GR_KP;
  if (chunk_out) *chunk_out = chunk;
  return (HW + chunk - 1) / chunk;
}

// Tensor-core path of mmsam_gram_bf16 (see neck.cu for the entry point and the SIMT fallback).
// Returns MMSAM_ERR_UNSUPPORTED when the shape does not fit (caller falls back).
int mmsam_gram_tc(const void* X, long long ld, int qoff, int koff, int n, int B, int HW, int blk, float* S, float* nq,
                  float* nk, cudaStream_t st) {
  using namespace mmsam;
  int chunk = 0;
  const int nchunks = mmsam_gram_tc_plan(n, B, HW, nq != nullptr, &chunk);
  if (nchunks == 0) return MMSAM_ERR_UNSUPPORTED;
  mmsam_host::EncodeTiledFn enc = mmsam_host::get_encode_tiled();
  if (!enc) return MMSAM_ERR_DRIVER;
  CUtensorMap tm;
  cuuint64_t dims[2] = {(cuuint64_t)ld, (cuuint64_t)B * HW};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64, GR_KP};
  cuuint32_t estr[2] = {1, 1};
  if (enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(X), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return MMSAM_ERR_DRIVER;
  const bool norms = nq != nullptr;
  const int bnj = (norms || n <= 128) ? 128 : 256;
  GramParams p;
  p.S = S; p.nq = nq; p.nk = nk; p.qoff = qoff; p.koff = koff; p.n = n; p.HW = HW; p.blk = blk;
  p.nti = (n + 127) / 128;
  p.ntj = (n + bnj - 1) / bnj;
  p.chunk = chunk;
  dim3 grid(p.nti * p.ntj, nchunks, B);
  if (bnj == 128) return launch_gram_tc<128>(tm, p, grid, st);
  return launch_gram_tc<256>(tm, p, grid, st);
}
