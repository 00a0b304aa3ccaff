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
  if (ksize == 7) {
    if (variant == 0) return x_dtype == MMSAM_F32 ? launch_dw<7, 4, 32, true, false>(p, grid, st) : launch_dw<7, 4, 32, false, false>(p, grid, st);
    return x_dtype == MMSAM_F32 ? launch_dw<7, 4, 32, true, true>(p, grid, st) : launch_dw<7, 4, 32, false, true>(p, grid, st);
  }
  return launch_dw<3, 8, 32, false, false>(p, grid, st);
}
