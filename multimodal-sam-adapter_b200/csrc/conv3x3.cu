// Grouped 3x3 convolution (stride 1, zero pad 1, no bias) on channels-last bf16 maps as an implicit GEMM
// on tcgen05 tensor cores — no im2col buffer.
//
// Serves the fusion neck's grouped convs (segmentation/mmseg_custom/models/backbones/
// adapter_modules_multimodal_mix_mod_new_in_twin_convnext_new.py): AttentionBase.qkv2
// (Conv2d(3c, 3c, 3, padding=1, groups=32), :84) and Mlp.dwconv (Conv2d(2c, 2c, 3, padding=1,
// groups=c), :118-119). Any groups count works, including 1 (dense 3x3).
//
// Tile = 128 output pixels (a 16-row x 8-column spatial patch of one image) x 64 output channels. The A operand of
// filter tap (dy, dx) is the same patch shifted by (dy-1, dx-1). ONE 4-D TMA box {64 ch, 8 w, 18 h, 1 b} per column
// shift dx serves the three taps of that column: the box is 18 image rows of 8 pixels = 18 SWIZZLE_128B atoms of 8
// rows x 128 B, so the row shift dy is a start address dy * 1024 B further into the box — still on an atom boundary,
// the descriptor needs nothing else. The hardware's out-of-bounds zero fill does the spatial zero padding and the
// channel tail. Because the conv is grouped, the 64 output channels of a tile only read a short window of input
// channels (the groups they belong to): the K loop is KC 64-channel blocks of that window x 3 column shifts (one
// pipeline unit each: the 18 KB box + the three 8 KB weight blocks of its taps, 12 MMAs of 128 x 64 x 16), against
// weight blocks pre-packed per (n-tile, tap, k-block) with zeros outside each group. Same warp specialisation as
// gemm.cu (TMA producer warp, one MMA-issuer lane, 8 epilogue warps, 2 TMEM accumulators).
//
// History (tools/bench_conv3x3.py, batch 8): with one shifted 16 KB box per TAP (8 x 16 patches, 9 boxes per k-block)
// the kernel took ~0.33 us per (tap, k-block) = ~600 cycles per 128-row box whatever the grouping: the per-row request
// rate of the box loads — not bytes or FLOPs — bounded it, and the same input rows were fetched nine times. Sharing a box
// between the three row shifts loads 3 x 144 instead of 9 x 128 rows per k-block. Halving the k-blocks per tile (n-tile
// stride below) pays in full; keeping the weights resident in shared memory (tried: n-tile-major ranges, 72 KB of tap
// blocks) does not.
#include "common.cuh"
#include <cstdlib>

namespace mmsam {

struct ConvParams {
  __nv_bfloat16* out;
  int B, H, W, Cin, Cout;
  int cg_in, cg_out;   // channels per group
  int KC;              // 64-channel k-blocks per tap
  int NS;              // output channels per n-tile (<= 64; the MMA tile stays 64 wide, the tail columns carry zero weights)
  int tiles_x, tiles_y, num_n, num_tiles;
};

static constexpr int CV_BN = 64, CV_STAGES = 4;
static constexpr int CV_TH = 16, CV_TW = 8;                                  // patch: rows x columns
static constexpr int CV_A_BYTES = (CV_TH + 2) * CV_TW * 64 * 2;              // 18 atoms of 8 pixels x 64 channels
static constexpr int CV_B_BYTES = CV_BN * 64 * 2;                            // one tap's weight block
static constexpr int CV_STAGE_BYTES = CV_A_BYTES + 3 * CV_B_BYTES;           // unit = one column shift of one k-block
static constexpr int CV_EPI_PITCH = 80;   // bytes per staged row (64 B of data): 16-byte stores of a quarter warp hit distinct banks
static constexpr int CV_SMEM = CV_STAGES * CV_STAGE_BYTES + 1024 + 256 + 8 * 32 * CV_EPI_PITCH;

__global__ void __launch_bounds__(320, 1)
conv3x3_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, const ConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + CV_STAGES * CV_STAGE_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + CV_STAGES;
  uint64_t* tfull = bars + 2 * CV_STAGES;
  uint64_t* tempty = bars + 2 * CV_STAGES + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * CV_STAGES + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_k = 3 * p.KC;          // pipeline units per tile

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW);
    for (int i = 0; i < CV_STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 8); }
    fence_barrier_init();
  }
  if (warp == 9) tmem_alloc(tmem_slot, 2 * CV_BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // tile -> (n tile fastest, then x, y, b)
  auto decode = [&](int tile, int& nt, int& x0, int& y0, int& b) {
    nt = tile % p.num_n;
    int t = tile / p.num_n;
    x0 = (t % p.tiles_x) * CV_TW;
    t /= p.tiles_x;
    y0 = (t % p.tiles_y) * CV_TH;
    b = t / p.tiles_y;
  };

  if (warp == 8) {
    {   // TMA producer: the whole warp runs the loop, one elected lane issues (see elect_one: a TMA issued from divergent code
        // costs a ~90-cycle ELECT / BRA.U.ANY round trip, and this kernel's k-blocks are only ~256 tensor-core cycles long)
      int s = 0; uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        int nt, x0, y0, b;
        decode(tile, nt, x0, y0, b);
        // first input channel of the first group touched, aligned down to 8 channels: a TMA box must start on
        // a 16-byte boundary of the innermost dimension
        const int kwin = (((nt * p.NS) / p.cg_out) * p.cg_in) & ~7;
        for (int kc = 0; kc < p.KC; ++kc) {
          for (int dx = 0; dx < 3; ++dx) {
            mbar_wait(&empty[s], ph ^ 1);
            if (elect_one()) {
              mbar_arrive_expect_tx(&full[s], CV_STAGE_BYTES);
              uint8_t* sa = smem + s * CV_STAGE_BYTES;
              asm volatile(
                  "cp.async.bulk.tensor.4d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                  ::"r"(smem_u32(sa)), "l"(reinterpret_cast<uint64_t>(&tmX)), "r"(smem_u32(&full[s])),
                  "r"(kwin + kc * 64), "r"(x0 + dx - 1), "r"(y0 - 1), "r"(b)
                  : "memory");
#pragma unroll
              for (int dy = 0; dy < 3; ++dy)
                tma_load_2d(sa + CV_A_BYTES + dy * CV_B_BYTES, &tmW, &full[s], 0, ((nt * 9 + dy * 3 + dx) * p.KC + kc) * 64);
            }
            __syncwarp();
            if (++s == CV_STAGES) { s = 0; ph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 9) {
    {   // the whole warp runs the loop, one elected lane issues (see elect_one)
      constexpr uint32_t idesc = umma_idesc_bf16(128, CV_BN, 0, 0);
      int s = 0; uint32_t ph = 0; int it = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        const uint32_t aph = (it >> 1) & 1;
        mbar_wait(&tempty[acc], aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * CV_BN;
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + s * CV_STAGE_BYTES);
          const uint32_t b_addr = a_addr + CV_A_BYTES;
          if (elect_one()) {
#pragma unroll
            for (int dy = 0; dy < 3; ++dy)        // row shift: dy atoms (8 pixels x 128 B) into the box
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_f16_ss(d_tmem, umma_desc_sw128(a_addr + dy * (CV_TW * 128) + k * 32),
                            umma_desc_sw128(b_addr + dy * CV_B_BYTES + k * 32), idesc, (kb | dy | k) != 0 ? 1u : 0u);
            umma_commit(&empty[s]);
          }
          __syncwarp();
          if (++s == CV_STAGES) { s = 0; ph ^= 1; }
        }
        if (elect_one()) umma_commit(&tfull[acc]);
        __syncwarp();
      }
    }
  } else {
    const int quad = warp & 3, half = warp >> 2;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      int nt, x0, y0, b;
      decode(tile, nt, x0, y0, b);
      const int acc = it & 1;
      const uint32_t aph = (it >> 1) & 1;
      mbar_wait(&tfull[acc], aph);
      tc_fence_after();
      uint32_t r[32];
      tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(quad * 32) << 16) + acc * CV_BN + half * 32, r);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
      // Unaligned / partial tiles: row-per-thread scalar stores cost one L1 wavefront per row and instruction; instead the warp parks its 32 rows
      // x 32 columns in a private shared-memory slab and writes them back two rows per instruction (16 lanes x 4 B
      // each), whatever the alignment of the tile's first channel (NS = 54 or 36 starts rows on odd 4-byte words).
      const int col = nt * p.NS + half * 32;
      // this warp's columns [col, col + nval): inside the tile's NS channels and inside the map
      int nval = p.NS - half * 32;
      nval = nval > 32 ? 32 : nval;
      nval = col + nval > p.Cout ? p.Cout - col : nval;
      if (nval == 32 && (col & 7) == 0) {
        // aligned full tile (NS = 64): four 16-byte stores per thread are cheaper than the slab round trip
        const int rr = quad * 32 + lane;
        const int y = y0 + (rr >> 3), x = x0 + (rr & 7);
        if (y < p.H && x < p.W) {
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
          uint4* op = reinterpret_cast<uint4*>(p.out + (((long long)b * p.H + y) * p.W + x) * p.Cout + col);
#pragma unroll
          for (int j = 0; j < 4; ++j) op[j] = pack8(v + 8 * j);
        }
        continue;
      }
      uint8_t* slab = smem + CV_STAGES * CV_STAGE_BYTES + 256 + warp * 32 * CV_EPI_PITCH;
      {
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
#pragma unroll
        for (int j = 0; j < 4; ++j) *reinterpret_cast<uint4*>(slab + lane * CV_EPI_PITCH + j * 16) = pack8(v + 8 * j);
      }
      __syncwarp();
      const int sub = lane & 15, rsel = lane >> 4;
#pragma unroll 4
      for (int i = 0; i < 16; ++i) {
        const int rl = 2 * i + rsel;                 // row of this warp's 32
        const int rr = quad * 32 + rl;
        const int y = y0 + (rr >> 3), x = x0 + (rr & 7);
        if (y < p.H && x < p.W && 2 * sub < nval) {
          const uint32_t w2 = *reinterpret_cast<const uint32_t*>(slab + rl * CV_EPI_PITCH + sub * 4);
          __nv_bfloat16* op = p.out + (((long long)b * p.H + y) * p.W + x) * p.Cout + col + 2 * sub;
          if (2 * sub + 1 < nval && ((col & 1) == 0)) *reinterpret_cast<uint32_t*>(op) = w2;
          else {
            op[0] = __ushort_as_bfloat16((unsigned short)(w2 & 0xffffu));
            if (2 * sub + 1 < nval) op[1] = __ushort_as_bfloat16((unsigned short)(w2 >> 16));
          }
        }
      }
      __syncwarp();     // the slab is rewritten by the next tile
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * CV_BN);
  }
}

}  // namespace mmsam

namespace mmsam {
// 64-channel k-blocks per tap when the output channels are tiled NS at a time: the widest input-channel window
// (first channel rounded down to a 16-byte boundary) any tile's groups span.
static int conv3x3_kblocks_for(int Cin, int Cout, int groups, int NS) {
  const int cgi = Cin / groups, cgo = Cout / groups;
  int kc = 1;
  for (int n0 = 0; n0 < Cout; n0 += NS) {
    const int n1 = (n0 + NS < Cout ? n0 + NS : Cout) - 1;
    const int g0 = n0 / cgo, g1 = n1 / cgo;
    const int len = (g1 + 1) * cgi - ((g0 * cgi) & ~7);
    const int need = (len + 63) / 64;
    if (need > kc) kc = need;
  }
  return kc;
}
}  // namespace mmsam

// Output channels per n-tile. The kernel is bound by the L2 -> shared-memory fill (9 taps x KC blocks of 24 KB per
// tile), so the stride that minimises (number of n-tiles) x KC wins: e.g. 9 channels per group (qkv2 of the 256^2
// level) -> 54 = 6 whole groups whose window fits ONE 64-channel block, instead of 64 outputs spanning 72 inputs.
MMSAM_API int mmsam_conv3x3_nstride(int Cin, int Cout, int groups) {
  if (groups <= 0 || Cin <= 0 || Cout <= 0 || Cin % groups || Cout % groups) return -1;
  if (const char* e = getenv("MMSAM_CONV3X3_NS")) {          // perf-debug: force the n-tile stride (64 = the old tiling)
    const int v = atoi(e);
    if (v >= 16 && v <= 64 && !(v & 1)) return v;
  }
  int best = 64;
  long long best_cost = -1;
  for (int ns = 64; ns >= 16; ns -= 2) {
    const long long cost = (long long)((Cout + ns - 1) / ns) * mmsam::conv3x3_kblocks_for(Cin, Cout, groups, ns);
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = ns; }
  }
  return best;
}

// Number of 64-channel k-blocks per tap for a given grouping (host helper, also used by the weight packer).
MMSAM_API int mmsam_conv3x3_kblocks(int Cin, int Cout, int groups) {
  const int ns = mmsam_conv3x3_nstride(Cin, Cout, groups);
  if (ns < 0) return -1;
  return mmsam::conv3x3_kblocks_for(Cin, Cout, groups, ns);
}

// x [B,H,W,Cin] bf16, w_packed bf16 [(ceil(Cout/NS) * 9 * KC) * 64, 64] with NS = mmsam_conv3x3_nstride(): block
// (nt, tap, kc) holds W[co = nt*NS + r, ci = kwin(nt) + kc*64 + c, tap] (0 outside co's group and for r >= NS),
// kwin(nt) = first input channel of the first group the tile touches rounded down to 8; out [B,H,W,Cout] bf16.
MMSAM_API int mmsam_conv3x3_bf16(const void* x, const void* w_packed, void* out, int B, int H, int W, int Cin,
                                 int Cout, int groups, int max_ctas, void* stream) {
  using namespace mmsam;
  if (B < 0 || H <= 0 || W <= 0 || Cin <= 0 || Cout <= 0 || (Cin & 7) || (Cout & 7)) return MMSAM_ERR_BAD_ARG;
  const int KC = mmsam_conv3x3_kblocks(Cin, Cout, groups);
  if (KC < 0) return MMSAM_ERR_BAD_ARG;
  if (B == 0) return MMSAM_OK;
  if (!x || !w_packed || !out || (((uintptr_t)x | (uintptr_t)w_packed | (uintptr_t)out) & 15)) return MMSAM_ERR_BAD_ARG;
  ConvParams p;
  p.out = (__nv_bfloat16*)out;
  p.B = B; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.cg_in = Cin / groups; p.cg_out = Cout / groups; p.KC = KC;
  p.NS = mmsam_conv3x3_nstride(Cin, Cout, groups);
  p.tiles_x = (W + CV_TW - 1) / CV_TW; p.tiles_y = (H + CV_TH - 1) / CV_TH; p.num_n = (Cout + p.NS - 1) / p.NS;
  p.num_tiles = p.num_n * p.tiles_x * p.tiles_y * B;
  mmsam_host::EncodeTiledFn enc = mmsam_host::get_encode_tiled();
  if (!enc) return MMSAM_ERR_DRIVER;
  CUtensorMap tmX, tmW;
  {
    cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)Cin * 2, (cuuint64_t)W * Cin * 2, (cuuint64_t)H * W * Cin * 2};
    cuuint32_t box[4] = {64, CV_TW, CV_TH + 2, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    if (enc(&tmX, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return MMSAM_ERR_DRIVER;
  }
  int rc = mmsam_host::make_tmap_2d_bf16(&tmW, w_packed, (uint64_t)p.num_n * 9 * KC * 64, 64, 64, 64, 64);
  if (rc) return rc;
  MMSAM_SET_SMEM_ONCE(conv3x3_kernel, CV_SMEM);
  if (max_ctas <= 0 || max_ctas > kNumSMs) max_ctas = kNumSMs;
  const int grid = p.num_tiles < max_ctas ? p.num_tiles : max_ctas;
  conv3x3_kernel<<<grid, 320, CV_SMEM, (cudaStream_t)stream>>>(tmX, tmW, p);
  MMSAM_LAUNCH_CHECK();
  return MMSAM_OK;
}
