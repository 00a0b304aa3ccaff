// Depthwise KxK convolution (stride 1, "same" zero padding) over channels-last bf16 maps.
//
// Serves ConvNeXtBlock.depthwise_conv 7x7 (segmentation/mmseg_custom/models/backbones/base/
// twin_convnext.py:98-101), ConvFFN's DWConv 3x3 applied with SHARED weights to the three token
// grids 128^2 | 64^2 | 32^2 of the concatenated sequence, fused with the GELU that follows it
// (adapter_modules_multimodal_mix_mod_new_in_twin_convnext_new.py:446-471), and MobileNetV2's
// 3x3 depthwise conv + ReLU6 (:281-295).
//
// HBM-bound. A CTA owns a TH x TW pixel tile x 64 channels: the (TH+K-1) x (TW+K-1) x 64 input halo
// is staged in shared memory with 16-byte loads (coalesced along C), weights for the 64 channels sit in
// shared memory as fp32 [K*K][64]. A thread owns one channel pair and 4 adjacent output columns for
// all TH rows: per filter row it keeps the K weights in registers and slides over 4+K-1 staged inputs,
// so the inner loop is pure fp32 FMA (K*K per output, the CUDA-core floor) with conflict-free 4-byte
// shared-memory reads.
#include "common.cuh"
#include <cstdlib>

namespace mmsam {

struct DwGrid {
  long long in_off, out_off;  // element offsets of this grid inside a batch item
  int H, W;
};
struct DwParams {
  const void* x;      // bf16, or fp32 (XF32: the ConvNeXt towers keep their residual stream in fp32; it is rounded to bf16 once,
                      // while being staged)
  __nv_bfloat16* y;
  const float* w;     // [K*K][C] fp32 (tap-major)
  const float* bias;  // [C] or null
  long long in_bstride, out_bstride;  // elements between batch items
  int C, B, ngrids, act;              // act: 0 none, 1 gelu, 3 relu6
  DwGrid g[3];
  int tiles_x[3], tiles_y[3], tile_start[4];
};

// SF32: the halo is staged as fp32 (converted once while staging) so that the inner loop reads its packed fp32x2 operand
// with one LDS.64 and no bf16 -> fp32 unpacking (2 ALU instructions per loaded value, ~10 % of the kernel's issue slots;
// the kernel is issue-bound: 67 % issue-active with the FMA pipe at 38 %). Used for the 7x7 (97 KB + weights, 2 CTAs / SM).
template <int K, int TH, int TW, bool XF32, bool SF32>
__global__ void __launch_bounds__(256, 2)
dwconv_kernel(const DwParams p) {
  constexpr int R = K / 2;
  constexpr int IH = TH + K - 1, IW = TW + K - 1;
  constexpr int CB = 64;  // channels per CTA
  constexpr int XO = 4;   // adjacent output columns per thread
  constexpr int SB = SF32 ? 4 : 2;   // bytes per staged element
  static_assert(TW == 32, "thread map assumes 8 column groups of 4");
  extern __shared__ __align__(16) uint8_t dw_smem[];
  __nv_bfloat16* s_in = reinterpret_cast<__nv_bfloat16*>(dw_smem);                 // [IH*IW][CB] (bf16 or fp32)
  float* s_inf = reinterpret_cast<float*>(dw_smem);
  float* s_w = reinterpret_cast<float*>(dw_smem + IH * IW * CB * SB);              // [K*K][CB]

  // which grid / tile
  int t = blockIdx.x, gi = 0;
  while (gi + 1 < p.ngrids && t >= p.tile_start[gi + 1]) ++gi;
  t -= p.tile_start[gi];
  const int tx = t % p.tiles_x[gi], ty = t / p.tiles_x[gi];
  const int H = p.g[gi].H, W = p.g[gi].W;
  const int c0 = blockIdx.y * CB;
  const int b = blockIdx.z;
  const long long xoff = (long long)b * p.in_bstride + p.g[gi].in_off;
  __nv_bfloat16* yout = p.y + (long long)b * p.out_bstride + p.g[gi].out_off;
  const int y0 = ty * TH, x0 = tx * TW;
  const int cvalid = p.C - c0 < CB ? p.C - c0 : CB;  // multiple of 8

  for (int i = threadIdx.x; i < K * K * CB; i += blockDim.x) {
    const int c = i % CB;
    s_w[i] = c < cvalid ? p.w[(i / CB) * p.C + c0 + c] : 0.f;
  }
  // stage the halo: 8 threads x 16 B cover the 64 channels of one pixel
  for (int i = threadIdx.x; i < IH * IW * (CB / 8); i += blockDim.x) {
    const int cv = i % (CB / 8);
    const int pix = i / (CB / 8);
    const int iy = y0 + pix / IW - R, ix = x0 + pix % IW - R;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (iy >= 0 && iy < H && ix >= 0 && ix < W && cv * 8 < cvalid) {
      const long long e = xoff + ((long long)iy * W + ix) * p.C + c0 + cv * 8;
      if constexpr (XF32) {
        const float4 fa = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.x) + e));
        const float4 fb = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.x) + e + 4));
        // the operands of the path are bf16: round the fp32 stream once here (SF32 and bf16 staging then agree bit for bit)
        v = make_uint4(pack_bf16(fa.x, fa.y), pack_bf16(fa.z, fa.w), pack_bf16(fb.x, fb.y), pack_bf16(fb.z, fb.w));
      } else {
        v = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.x) + e));
      }
    }
    if constexpr (SF32) {
      *reinterpret_cast<float4*>(&s_inf[pix * CB + cv * 8]) = make_float4(bf16lo(v.x), bf16hi(v.x), bf16lo(v.y), bf16hi(v.y));
      *reinterpret_cast<float4*>(&s_inf[pix * CB + cv * 8 + 4]) = make_float4(bf16lo(v.z), bf16hi(v.z), bf16lo(v.w), bf16hi(v.w));
    } else {
      *reinterpret_cast<uint4*>(&s_in[(pix * CB) + cv * 8]) = v;
    }
  }
  __syncthreads();

  // thread -> (channel pair cp, group of 4 output columns xg); all TH rows of the tile
  const int cp = threadIdx.x & 31;
  const int xg = threadIdx.x >> 5;
  const int c = c0 + cp * 2;
  // packed fp32x2 accumulators: .x = channel c, .y = channel c + 1 (one FFMA2 per tap and output instead of two FFMA;
  // three-register FFMA issues every 2 cycles per scheduler, FFMA2 every 3 for twice the work)
  u64 acc[TH][XO];
  {
    float b0 = 0.f, b1 = 0.f;
    if (p.bias && cp * 2 < cvalid) { b0 = p.bias[c]; b1 = p.bias[c + 1]; }
#pragma unroll
    for (int r = 0; r < TH; ++r)
#pragma unroll
      for (int xo = 0; xo < XO; ++xo) acc[r][xo] = pack2(b0, b1);
  }
  const uint32_t* s_in32 = reinterpret_cast<const uint32_t*>(s_in);
#pragma unroll
  for (int ky = 0; ky < K; ++ky) {
    u64 w[K];
#pragma unroll
    for (int kx = 0; kx < K; ++kx) w[kx] = *reinterpret_cast<const u64*>(&s_w[(ky * K + kx) * CB + cp * 2]);
#pragma unroll
    for (int r = 0; r < TH; ++r) {
      u64 in[XO + K - 1];
#pragma unroll
      for (int i = 0; i < XO + K - 1; ++i) {
        if constexpr (SF32) {
          in[i] = *reinterpret_cast<const u64*>(&s_inf[((r + ky) * IW + xg * XO + i) * CB + cp * 2]);
        } else {
          const uint32_t v = s_in32[((r + ky) * IW + xg * XO + i) * (CB / 2) + cp];
          in[i] = pack2(bf16lo(v), bf16hi(v));
        }
      }
#pragma unroll
      for (int xo = 0; xo < XO; ++xo)
#pragma unroll
        for (int kx = 0; kx < K; ++kx) acc[r][xo] = fma2(in[xo + kx], w[kx], acc[r][xo]);
    }
  }
  if (cp * 2 >= cvalid) return;
#pragma unroll
  for (int r = 0; r < TH; ++r) {
    const int oy = y0 + r;
    if (oy >= H) break;
#pragma unroll
    for (int xo = 0; xo < XO; ++xo) {
      const int ox = x0 + xg * XO + xo;
      if (ox >= W) continue;
      float a0, a1;
      unpack2(acc[r][xo], a0, a1);
      if (p.act == 1) { a0 = gelu_erf(a0); a1 = gelu_erf(a1); }
      else if (p.act == 3) { a0 = fminf(fmaxf(a0, 0.f), 6.f); a1 = fminf(fmaxf(a1, 0.f), 6.f); }
      *reinterpret_cast<uint32_t*>(yout + ((long long)oy * W + ox) * p.C + c) = pack_bf16(a0, a1);
    }
  }
}

template <int K, int TH, int TW, bool XF32, bool SF32>
static int launch_dw(const DwParams& p, dim3 grid, cudaStream_t st) {
  constexpr int smem = (TH + K - 1) * (TW + K - 1) * 64 * (SF32 ? 4 : 2) + K * K * 64 * 4;
  MMSAM_SET_SMEM_ONCE((dwconv_kernel<K, TH, TW, XF32, SF32>), smem);
  dwconv_kernel<K, TH, TW, XF32, SF32><<<grid, 256, smem, st>>>(p);
  MMSAM_LAUNCH_CHECK();
  return MMSAM_OK;
}

// ---- 7x7 on the fp32 residual stream of the ConvNeXt towers (72 launches per step) ---------------------------------------
// The generic kernel above is bound by shared-memory bandwidth, not by the FMA pipe: with 4 output columns per thread it
// reads 10 staged fp32 pairs (LDS.64 = 2 wavefronts each) per 28 packed FMAs, 23.5 wavefronts against 21 FMA-pipe cycles per
// SM, and its 64-channel CTAs waste a third of the second channel block at C = 96. This kernel:
//   * thread = one channel pair x 8 adjacent output columns x 4 rows: 14 loads per 56 FFMA2 (31.5 wavefronts against 42
//     FMA-pipe cycles) -> bound by the FMA pipe, the floor of a depthwise conv on CUDA cores;
//   * CTA = 32 channels (divides 96 / 192 / 384 / 768) x 32 columns x 4 RG rows, 64 RG threads;
//   * the (4 RG + 6) x 38 x 32-channel fp32 halo arrives by ONE 4-D TMA box whose out-of-bounds elements are zero-filled
//     by the copy engine ("same" padding, ragged right / bottom tiles); fp32 operands are used as they are (no bf16
//     rounding of the stream);
//   * the bf16 output tile is staged over the consumed halo and leaves by ONE TMA store (clipped at the map's edges).
template <int RG>
__global__ void __launch_bounds__(64 * RG, RG == 4 ? 2 : 3)
dwconv7_tma_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY, const float* __restrict__ wt,
                   const float* __restrict__ bias, int C, int tiles_x) {
  constexpr int K = 7, CB = 32, XO = 8, TH = 4, TW = 32;
  constexpr int ROWS = TH * RG, IH = ROWS + K - 1, IW = TW + K - 1;
  extern __shared__ __align__(128) uint8_t dw_smem[];
  uint8_t* base = dw_smem + ((128u - (smem_u32(dw_smem) & 127u)) & 127u);
  float* s_in = reinterpret_cast<float*>(base);                                  // [IH][IW][CB] fp32
  float* s_w = reinterpret_cast<float*>(base + IH * IW * CB * 4);                // [K*K][CB]
  uint64_t* bar = reinterpret_cast<uint64_t*>(base + IH * IW * CB * 4 + K * K * CB * 4);
  const int tx = blockIdx.x % tiles_x, ty = blockIdx.x / tiles_x;
  const int c0 = blockIdx.y * CB, b = blockIdx.z;
  const int y0 = ty * ROWS, x0 = tx * TW;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  pdl_wait();
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar, IH * IW * CB * 4);
    tma_load_4d(s_in, &tmX, bar, c0, x0 - K / 2, y0 - K / 2, b);
  }
  for (int i = threadIdx.x; i < K * K * CB; i += blockDim.x) s_w[i] = wt[(i / CB) * C + c0 + (i % CB)];
  const int cp = threadIdx.x & 15, xg = (threadIdx.x >> 4) & 3, rg = threadIdx.x >> 6;
  u64 acc[TH][XO];
  {
    const float b0 = bias ? bias[c0 + cp * 2] : 0.f, b1 = bias ? bias[c0 + cp * 2 + 1] : 0.f;
#pragma unroll
    for (int r = 0; r < TH; ++r)
#pragma unroll
      for (int xo = 0; xo < XO; ++xo) acc[r][xo] = pack2(b0, b1);
  }
  __syncthreads();
  mbar_wait(bar, 0);
  const float* in0 = s_in + ((rg * TH) * IW + xg * XO) * CB + cp * 2;
#pragma unroll
  for (int ky = 0; ky < K; ++ky) {
    u64 w[K];
#pragma unroll
    for (int kx = 0; kx < K; ++kx) w[kx] = *reinterpret_cast<const u64*>(&s_w[(ky * K + kx) * CB + cp * 2]);
#pragma unroll
    for (int r = 0; r < TH; ++r) {
      u64 in[XO + K - 1];
#pragma unroll
      for (int i = 0; i < XO + K - 1; ++i) in[i] = *reinterpret_cast<const u64*>(in0 + ((r + ky) * IW + i) * CB);
#pragma unroll
      for (int xo = 0; xo < XO; ++xo)
#pragma unroll
        for (int kx = 0; kx < K; ++kx) acc[r][xo] = fma2(in[xo + kx], w[kx], acc[r][xo]);
    }
  }
  __syncthreads();                      // every warp is done with the halo: reuse it as the [ROWS][TW][CB] bf16 output tile
  uint32_t* s_out = reinterpret_cast<uint32_t*>(base);
#pragma unroll
  for (int r = 0; r < TH; ++r)
#pragma unroll
    for (int xo = 0; xo < XO; ++xo) {
      float a0, a1;
      unpack2(acc[r][xo], a0, a1);
      s_out[((rg * TH + r) * TW + xg * XO + xo) * (CB / 2) + cp] = pack_bf16(a0, a1);
    }
  fence_proxy_async();
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(&tmY)),
                 "r"(smem_u32(s_out)), "r"(c0), "r"(x0), "r"(y0), "r"(b)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  }
}

template <int RG>
static int launch_dw7_tma(const CUtensorMap& tmX, const CUtensorMap& tmY, const float* w, const float* bias, int C, int H, int W,
                          int B, cudaStream_t st) {
  constexpr int smem = (4 * RG + 6) * 38 * 32 * 4 + 49 * 32 * 4 + 16 + 128;
  MMSAM_SET_SMEM_ONCE((dwconv7_tma_kernel<RG>), smem);
  const int tiles_x = (W + 31) / 32, tiles_y = (H + 4 * RG - 1) / (4 * RG);
  cudaError_t le = mmsam_host::launch_pdl(dwconv7_tma_kernel<RG>, dim3(tiles_x * tiles_y, C / 32, B), dim3(64 * RG), smem, st, tmX, tmY, w, bias, C, tiles_x);
  if (le != cudaSuccess) return (int)le;
  MMSAM_LAUNCH_CHECK();
  return MMSAM_OK;
}

// fp32 [B][H][W][C] -> bf16 [B][H][W][C] through the TMA kernel; batch strides in elements
static int dwconv7_tma(const float* x, __nv_bfloat16* y, const float* w, const float* bias, int B, int C, int H, int W,
                       long long in_bstride, long long out_bstride, cudaStream_t st) {
  mmsam_host::EncodeTiledFn enc = mmsam_host::get_encode_tiled();
  if (!enc) return MMSAM_ERR_DRIVER;
  static const int rg_env = [] { const char* e = getenv("MMSAM_DW_RG"); return e ? atoi(e) : 0; }();
  // rows per CTA: 8 (128 threads, 3 CTAs / SM) measured 3-10 % faster than 16 (256 threads, 2 CTAs / SM) at every stage of
  // the step (110 / 60 / 33 / 21 us vs 114 / 64 / 37 / 22): 22.4 TFMA/s at stage 0 = the FFMA2 issue rate of the chip.
  // A persistent 256-thread CTA per SM with the halo and the tap block double-buffered (what made the 3x3 kernel below
  // 20 % faster) measured 115 / 63 / 37 / 21 us: this kernel waits for its FMA pipe, not for its loads.
  const int rg = rg_env == 4 ? 4 : 2;
  CUtensorMap tmX, tmY;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  {
    cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)in_bstride * 4};
    cuuint32_t box[4] = {32, 38, (cuuint32_t)(4 * rg + 6), 1};
    if (enc(&tmX, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(x), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return MMSAM_ERR_DRIVER;
  }
  {
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)out_bstride * 2};
    cuuint32_t box[4] = {32, 32, (cuuint32_t)(4 * rg), 1};
    if (enc(&tmY, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, y, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return MMSAM_ERR_DRIVER;
  }
  return rg == 2 ? launch_dw7_tma<2>(tmX, tmY, w, bias, C, H, W, B, st) : launch_dw7_tma<4>(tmX, tmY, w, bias, C, H, W, B, st);
}

// ---- 3x3 on bf16 maps (ConvFFN's DWConv over the three token grids + GELU, MobileNetV2's depthwise + ReLU6) -----------------
// Same structure as the 7x7 above: one 4-D TMA box per tile (zero-filled halo), thread = channel pair x 8 columns x 4 rows,
// 32-channel CTAs, TMA store of the bf16 tile. The halo is staged as bf16 (the copy engine cannot convert), unpacked on
// load; the fused activation runs on the packed pair. Up to three grids per launch (one tensor-map pair each).
struct Dw3Maps { CUtensorMap x[3], y[3]; };
struct Dw3Params {
  const float* w;      // [9][C] tap-major
  const float* bias;   // [C] or null
  int C, act, ngrids, B;
  int tiles_x[3], tile_start[4];
};

// Persistent: a CTA walks items (tile, 32-channel block, image; channel block fastest, so the CTAs that run together
// read neighbouring 64-byte pieces of the same pixels) with the halo double-buffered — the TMA load of the next item is in
// flight while the current one is computed (a one-shot CTA per tile exposed the load latency with two CTAs per SM:
// 2.6 - 3.1 TB/s). The output tile is staged over the halo it came from; the next load into that buffer waits for
// the bulk store to have read it.
__global__ void __launch_bounds__(256, 2)
dwconv3_tma_kernel(const __grid_constant__ Dw3Maps maps, const Dw3Params p) {
  constexpr int K = 3, CB = 32, XO = 8, TH = 4, TW = 32, RG = 4;
  constexpr int ROWS = TH * RG, IH = ROWS + K - 1, IW = TW + K - 1;
  constexpr int HALO = (IH * IW * CB * 2 + 127) / 128 * 128;
  extern __shared__ __align__(128) uint8_t dw_smem[];
  uint8_t* base = dw_smem + ((128u - (smem_u32(dw_smem) & 127u)) & 127u);
  float* s_w = reinterpret_cast<float*>(base + 2 * HALO);                          // [9][CB]
  uint64_t* bar = reinterpret_cast<uint64_t*>(base + 2 * HALO + K * K * CB * 4);   // [2]
  const int ncb = p.C / CB, ntile = p.tile_start[3];
  const int items = ntile * ncb * p.B;
  auto decode = [&](int item, int& gi, int& x0, int& y0, int& c0, int& b) {
    const int r = item / ncb;
    c0 = (item - r * ncb) * CB;
    b = r / ntile;
    int t = r - b * ntile;
    gi = 0;
    while (gi + 1 < p.ngrids && t >= p.tile_start[gi + 1]) ++gi;
    t -= p.tile_start[gi];
    const int ty = t / p.tiles_x[gi];
    x0 = (t - ty * p.tiles_x[gi]) * TW;
    y0 = ty * ROWS;
  };
  auto issue = [&](int item, int buf) {       // one thread
    int gi, x0, y0, c0, b;
    decode(item, gi, x0, y0, c0, b);
    mbar_arrive_expect_tx(&bar[buf], IH * IW * CB * 2);
    tma_load_4d(base + buf * HALO, &maps.x[gi], &bar[buf], c0, x0 - 1, y0 - 1, b);
  };
  if (threadIdx.x == 0) {
    mbar_init(&bar[0], 1); mbar_init(&bar[1], 1);
    fence_barrier_init();
  }
  pdl_wait();
  if (threadIdx.x == 0 && (int)blockIdx.x < items) issue(blockIdx.x, 0);
  const int cp = threadIdx.x & 15, xg = (threadIdx.x >> 4) & 3, rg = threadIdx.x >> 6;
  int it = 0;
  for (int item = blockIdx.x; item < items; item += gridDim.x, ++it) {
    const int buf = it & 1;
    int gi, x0, y0, c0, b;
    decode(item, gi, x0, y0, c0, b);
    __syncthreads();                    // every thread is done with s_w and with the other buffer (item - 1)
    if (threadIdx.x == 0 && item + (int)gridDim.x < items) {
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");      // the store of item - 1 has read that buffer
      issue(item + gridDim.x, buf ^ 1);
    }
    for (int i = threadIdx.x; i < K * K * CB; i += blockDim.x) s_w[i] = p.w[(i / CB) * p.C + c0 + (i % CB)];
    u64 acc[TH][XO];
    {
      const float b0 = p.bias ? p.bias[c0 + cp * 2] : 0.f, b1 = p.bias ? p.bias[c0 + cp * 2 + 1] : 0.f;
#pragma unroll
      for (int r = 0; r < TH; ++r)
#pragma unroll
        for (int xo = 0; xo < XO; ++xo) acc[r][xo] = pack2(b0, b1);
    }
    __syncthreads();
    mbar_wait(&bar[buf], (it >> 1) & 1);
    const uint32_t* s_in = reinterpret_cast<const uint32_t*>(base + buf * HALO);      // [IH][IW][CB / 2] bf16 pairs
    const uint32_t* in0 = s_in + ((rg * TH) * IW + xg * XO) * (CB / 2) + cp;
#pragma unroll
    for (int ky = 0; ky < K; ++ky) {
      u64 w[K];
#pragma unroll
      for (int kx = 0; kx < K; ++kx) w[kx] = *reinterpret_cast<const u64*>(&s_w[(ky * K + kx) * CB + cp * 2]);
#pragma unroll
      for (int r = 0; r < TH; ++r) {
        u64 in[XO + K - 1];
#pragma unroll
        for (int i = 0; i < XO + K - 1; ++i) {
          const uint32_t v = in0[((r + ky) * IW + i) * (CB / 2)];
          in[i] = pack2(bf16lo(v), bf16hi(v));
        }
#pragma unroll
        for (int xo = 0; xo < XO; ++xo)
#pragma unroll
          for (int kx = 0; kx < K; ++kx) acc[r][xo] = fma2(in[xo + kx], w[kx], acc[r][xo]);
      }
    }
    __syncthreads();                      // every warp is done with the halo: reuse it as the [ROWS][TW][CB] bf16 output tile
    uint32_t* s_out = reinterpret_cast<uint32_t*>(base + buf * HALO);
#pragma unroll
    for (int r = 0; r < TH; ++r)
#pragma unroll
      for (int xo = 0; xo < XO; ++xo) {
        float a0, a1;
        unpack2(acc[r][xo], a0, a1);
        if (p.act == 1) gelu_erf2(a0, a1);
        else if (p.act == 3) { a0 = fminf(fmaxf(a0, 0.f), 6.f); a1 = fminf(fmaxf(a1, 0.f), 6.f); }
        s_out[((rg * TH + r) * TW + xg * XO + xo) * (CB / 2) + cp] = pack_bf16(a0, a1);
      }
    fence_proxy_async();
    __syncthreads();
    if (threadIdx.x == 0) {
      asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                       reinterpret_cast<uint64_t>(&maps.y[gi])),
                   "r"(smem_u32(s_out)), "r"(c0), "r"(x0), "r"(y0), "r"(b)
                   : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  }
  if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

static int dwconv3_tma(const __nv_bfloat16* x, __nv_bfloat16* y, const float* w, const float* bias, int B, int C, int ngrids,
                       const DwGrid* g, long long in_bstride, long long out_bstride, int act, cudaStream_t st) {
  mmsam_host::EncodeTiledFn enc = mmsam_host::get_encode_tiled();
  if (!enc) return MMSAM_ERR_DRIVER;
  Dw3Maps maps;
  Dw3Params p;
  p.w = w; p.bias = bias; p.C = C; p.act = act; p.ngrids = ngrids;
  int total = 0;
  cuuint32_t estr[4] = {1, 1, 1, 1};
  for (int i = 0; i < 3; ++i) {
    const int gi = i < ngrids ? i : 0;          // unused slots repeat grid 0 (a valid descriptor, never dereferenced)
    const int H = g[gi].H, W = g[gi].W;
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t sx[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)in_bstride * 2};
    cuuint64_t sy[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)out_bstride * 2};
    cuuint32_t bx[4] = {32, 34, 18, 1}, by[4] = {32, 32, 16, 1};
    if (enc(&maps.x[i], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<__nv_bfloat16*>(x) + g[gi].in_off, dims, sx, bx, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return MMSAM_ERR_DRIVER;
    if (enc(&maps.y[i], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, y + g[gi].out_off, dims, sy, by, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return MMSAM_ERR_DRIVER;
    p.tiles_x[i] = i < ngrids ? (W + 31) / 32 : 0;
    p.tile_start[i] = total;
    if (i < ngrids) total += p.tiles_x[i] * ((H + 15) / 16);
  }
  p.tile_start[3] = total;
  constexpr int smem = 2 * ((18 * 34 * 32 * 2 + 127) / 128 * 128) + 9 * 32 * 4 + 16 + 128;
  MMSAM_SET_SMEM_ONCE(dwconv3_tma_kernel, smem);
  p.B = B;
  const long long items = (long long)total * (C / 32) * B;
  if (items > 0x7fffffffLL) return MMSAM_ERR_UNSUPPORTED;
  const int grid = items < 2 * kNumSMs ? (int)items : 2 * kNumSMs;
  cudaError_t le = mmsam_host::launch_pdl(dwconv3_tma_kernel, dim3(grid), dim3(256), smem, st, maps, p);
  if (le != cudaSuccess) return (int)le;
  MMSAM_LAUNCH_CHECK();
  return MMSAM_OK;
}

}  // namespace mmsam

// See include/mmsam_b200.h for the contract.
MMSAM_API int mmsam_dwconv(const void* x, int x_dtype, void* y, const float* w_tap_major, const float* bias, int B,
                                int C, int ksize, int ngrids, const int* grid_hw_host,
                                const long long* grid_in_off_host, const long long* grid_out_off_host,
                                long long in_bstride, long long out_bstride, int act, void* stream) {
  using namespace mmsam;
  if (B < 0 || C <= 0 || (C & 7) || ngrids < 1 || ngrids > 3) return MMSAM_ERR_BAD_ARG;
  if (ksize != 3 && ksize != 7) return MMSAM_ERR_UNSUPPORTED;
  if (x_dtype != MMSAM_BF16 && x_dtype != MMSAM_F32) return MMSAM_ERR_BAD_DTYPE;
  if (x_dtype == MMSAM_F32 && ksize != 7) return MMSAM_ERR_UNSUPPORTED;    // fp32 input: the ConvNeXt 7x7 only
  if (act != 0 && act != 1 && act != 3) return MMSAM_ERR_BAD_ARG;
  if (B == 0) return MMSAM_OK;
  if (!x || !y || !w_tap_major || !grid_hw_host) return MMSAM_ERR_BAD_ARG;
  if ((((uintptr_t)x | (uintptr_t)y) & 15) || (in_bstride & 7) || (out_bstride & 7)) return MMSAM_ERR_BAD_ARG;
  DwParams p;
  p.x = x; p.y = (__nv_bfloat16*)y; p.w = w_tap_major; p.bias = bias;
  p.in_bstride = in_bstride; p.out_bstride = out_bstride; p.C = C; p.B = B; p.ngrids = ngrids; p.act = act;
  const int TH = ksize == 7 ? 4 : 8, TW = 32;
  int total = 0;
  for (int i = 0; i < 3; ++i) {
    if (i < ngrids) {
      p.g[i].H = grid_hw_host[2 * i]; p.g[i].W = grid_hw_host[2 * i + 1];
      p.g[i].in_off = grid_in_off_host ? grid_in_off_host[i] : 0;
      p.g[i].out_off = grid_out_off_host ? grid_out_off_host[i] : 0;
      if (p.g[i].H <= 0 || p.g[i].W <= 0 || (p.g[i].in_off & 7) || (p.g[i].out_off & 7)) return MMSAM_ERR_BAD_ARG;
      p.tiles_x[i] = (p.g[i].W + TW - 1) / TW;
      p.tiles_y[i] = (p.g[i].H + TH - 1) / TH;
    } else {
      p.g[i].H = p.g[i].W = 0; p.g[i].in_off = p.g[i].out_off = 0; p.tiles_x[i] = p.tiles_y[i] = 0;
    }
    p.tile_start[i] = total;
    total += p.tiles_x[i] * p.tiles_y[i];
  }
  p.tile_start[3] = total;
  dim3 grid(total, (C + 63) / 64, B);
  cudaStream_t st = (cudaStream_t)stream;
  static const int variant = [] { const char* e = getenv("MMSAM_DW_VARIANT"); return e ? atoi(e) : 1; }();   // 0: bf16 staging
  if (ksize == 7 && x_dtype == MMSAM_F32 && ngrids == 1 && C % 32 == 0 && act == 0 && variant != 0 && variant != 2 &&
      (long long)p.g[0].W * C % 4 == 0 && (in_bstride & 3) == 0)
    return dwconv7_tma(reinterpret_cast<const float*>(x) + p.g[0].in_off, p.y + p.g[0].out_off, w_tap_major, bias, B, C, p.g[0].H,
                       p.g[0].W, in_bstride, out_bstride, st);
  if (ksize == 7) {
    if (variant == 0) return x_dtype == MMSAM_F32 ? launch_dw<7, 4, 32, true, false>(p, grid, st) : launch_dw<7, 4, 32, false, false>(p, grid, st);
    return x_dtype == MMSAM_F32 ? launch_dw<7, 4, 32, true, true>(p, grid, st) : launch_dw<7, 4, 32, false, true>(p, grid, st);
  }
  if (variant != 0 && variant != 2 && C % 32 == 0) {
    bool ok = true;      // TMA needs 16-byte global strides
    for (int i = 0; i < ngrids; ++i) ok = ok && ((long long)p.g[i].W * C % 8 == 0);
    if (ok) return dwconv3_tma(reinterpret_cast<const __nv_bfloat16*>(x), p.y, w_tap_major, bias, B, C, ngrids, p.g, in_bstride, out_bstride, act, st);
  }
  return launch_dw<3, 8, 32, false, false>(p, grid, st);
}
