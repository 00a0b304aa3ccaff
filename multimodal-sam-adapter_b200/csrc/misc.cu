// Small HBM-bound layout / resampling kernels around the GEMMs.
#include "common.cuh"
#include <cmath>
#include <cstdlib>

MMSAM_API int mmsam_arch(void) { return 100; }

namespace mmsam {

// NCHW fp32 image -> patch-major bf16 rows [(b, py, px), (c, ky, kx)]: turns the non-overlapping
// strided convs (ViT patch embed 16x16/s16, image_encoder.py:662-671; ConvNeXt stem 4x4/s4,
// twin_convnext.py:295-312) into GEMMs against the flattened conv weight [Cout, C*p*p].
// A thread owns one (patch, channel, patch row): p contiguous floats in, p contiguous bf16 out (16-byte loads, 8/16-byte
// stores); consecutive threads walk (c, ky) of one patch, so a warp writes one contiguous span of the output row and reads
// 16/64-byte pieces that neighbouring patches complete to full sectors. (One element per thread scattered 2-byte stores
// over 8 sectors per warp instruction: 208 us per launch against a 25 us HBM floor.)
template <int P>
__global__ void __launch_bounds__(256)
patchify_kernel(const float* __restrict__ img, __nv_bfloat16* __restrict__ out, int B, int Ctot, int c_off,
                int C, int H, int W) {
  const int PW = W / P, PH = H / P;
  const long long total = (long long)B * PH * PW * C * P;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int ky = (int)(idx % P);
    long long t = idx / P;
    const int c = (int)(t % C); t /= C;
    const int px = (int)(t % PW); t /= PW;
    const int py = (int)(t % PH);
    const int b = (int)(t / PH);
    const float4* src = reinterpret_cast<const float4*>(img + (((long long)b * Ctot + c_off + c) * H + py * P + ky) * W + px * P);
    __nv_bfloat16* dst = out + (((long long)b * PH + py) * PW + px) * (C * P * P) + (c * P + ky) * P;
#pragma unroll
    for (int j = 0; j < P / 4; ++j) {
      const float4 v = __ldg(src + j);
      uint2 o;
      o.x = pack_bf16(v.x, v.y);
      o.y = pack_bf16(v.z, v.w);
      *reinterpret_cast<uint2*>(dst + 4 * j) = o;
    }
  }
}
// any patch size / alignment: one element per thread
__global__ void __launch_bounds__(256)
patchify_generic_kernel(const float* __restrict__ img, __nv_bfloat16* __restrict__ out, int B, int Ctot, int c_off,
                        int C, int H, int W, int p) {
  const int PW = W / p, PH = H / p;
  const long long total = (long long)B * PH * PW * C * p * p;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int kx = (int)(idx % p);
    long long t = idx / p;
    const int px = (int)(t % PW); t /= PW;
    const int ky = (int)(t % p); t /= p;
    const int c = (int)(t % C); t /= C;
    const int py = (int)(t % PH);
    const int b = (int)(t / PH);
    const float v = __ldg(img + (((long long)b * Ctot + c_off + c) * H + py * p + ky) * W + px * p + kx);
    out[(((long long)b * PH + py) * PW + px) * (C * p * p) + (c * p + ky) * p + kx] = __float2bfloat16_rn(v);
  }
}

// The step before the path (pipelines/transform.py:2796-2806 Normalize_multimodal + formatting.py ImageToTensor): one
// modality's decoded image, HWC uint8, -> (v * inv255 - mean[c]) / std[c] written as channels [c_off, c_off + C) of the
// fp32 NCHW network input. A thread owns 4 consecutive pixels of a row: 4 * C byte loads, one float4 store per channel.
__global__ void __launch_bounds__(256)
normalize_u8_kernel(const uint8_t* __restrict__ img, float* __restrict__ out, int B, int H, int W, int C, int Ctot,
                    int c_off, float m0, float m1, float m2, float m3, float r0, float r1, float r2, float r3, float pre) {
  const float mean[4] = {m0, m1, m2, m3}, rstd[4] = {r0, r1, r2, r3};
  const int W4 = W >> 2;
  const long long total = (long long)B * H * W4;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int x4 = (int)(idx % W4);
    long long t = idx / W4;
    const int y = (int)(t % H);
    const int b = (int)(t / H);
    const uint8_t* src = img + (((long long)b * H + y) * W + x4 * 4) * C;
    for (int c = 0; c < C; ++c) {
      float4 v;
      v.x = ((float)src[c] * pre - mean[c]) * rstd[c];
      v.y = ((float)src[C + c] * pre - mean[c]) * rstd[c];
      v.z = ((float)src[2 * C + c] * pre - mean[c]) * rstd[c];
      v.w = ((float)src[3 * C + c] * pre - mean[c]) * rstd[c];
      *reinterpret_cast<float4*>(out + (((long long)b * Ctot + c_off + c) * H + y) * W + x4 * 4) = v;
    }
  }
}

// out[b,y,x,c] = (base[b,y,x,c] + bilinear(src[b])[y,x,c]) * scale[c] + shift[c]   (NHWC bf16)
// PyTorch bilinear, align_corners=False. Serves the ViT-feature fusion + eval BatchNorm at the end of
// the backbone (..._new.py:326-337) and the head's resize-into-concat (segformer_head.py:55-61).
// Grid = (x * channel-vector chunks, blocks of RA_YB rows, batch): a thread keeps its (x, 8 channels) for RA_YB consecutive
// output rows — one 32-bit division per thread instead of three 64-bit ones per element, horizontal corners / weights and
// the BatchNorm scale / shift (registers) computed once per RA_YB outputs (see resize_sum_affine_kernel).
static constexpr int RA_YB = 8;
// 8 consecutive channels of a source pixel; SF32: the source map is fp32 (the ViT token stream)
template <bool SF32>
__device__ __forceinline__ void ra_load8(const void* base, long long elem_off, float* f) {
  if constexpr (SF32) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + elem_off));
    const float4 b = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + elem_off + 4));
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  } else {
    unpack8(__ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(base) + elem_off)), f);
  }
}
template <bool SF32>
__global__ void __launch_bounds__(256, 3)
resize_add_affine_kernel(const void* __restrict__ src, const __nv_bfloat16* __restrict__ base,
                         const float* __restrict__ scale, const float* __restrict__ shift,
                         __nv_bfloat16* __restrict__ out, int B, int Hs, int Ws, int Ho, int Wo, int C,
                         long long src_bstride, long long base_bstride, long long out_bstride, long long ldo,
                         long long lds, long long ldb, float rh, float rw) {
  const uint32_t CV = (uint32_t)C >> 3;
  const uint32_t i = blockIdx.x * 256u + threadIdx.x;
  if (i >= (uint32_t)Wo * CV) return;
  const uint32_t x = i / CV, cv = i - x * CV;
  const int b = blockIdx.z;
  const int y_begin = blockIdx.y * RA_YB, y_end = min(y_begin + RA_YB, Ho);
  float sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { sc[j] = 1.f; sh[j] = 0.f; }
  if (scale) {
    const float4 a0 = __ldg(reinterpret_cast<const float4*>(scale + cv * 8)), a1 = __ldg(reinterpret_cast<const float4*>(scale + cv * 8 + 4));
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(shift + cv * 8)), b1 = __ldg(reinterpret_cast<const float4*>(shift + cv * 8 + 4));
    sc[0] = a0.x; sc[1] = a0.y; sc[2] = a0.z; sc[3] = a0.w; sc[4] = a1.x; sc[5] = a1.y; sc[6] = a1.z; sc[7] = a1.w;
    sh[0] = b0.x; sh[1] = b0.y; sh[2] = b0.z; sh[3] = b0.w; sh[4] = b1.x; sh[5] = b1.y; sh[6] = b1.z; sh[7] = b1.w;
  }
  const bool same = Hs == Ho && Ws == Wo;
  float sx = (x + 0.5f) * rw - 0.5f;
  sx = sx < 0.f ? 0.f : sx;
  int x0 = same ? (int)x : (int)sx;
  x0 = x0 > Ws - 1 ? Ws - 1 : x0;
  const int x1 = x0 < Ws - 1 ? x0 + 1 : x0;
  const float lx = same ? 0.f : sx - x0;
  const long long sb = b * src_bstride + cv * 8;            // element offset of this thread's channels in image b
  const long long so0 = (long long)x0 * lds, so1 = (long long)x1 * lds;
  const __nv_bfloat16* bp = base ? base + b * base_bstride + ((long long)y_begin * Wo + x) * ldb + cv * 8 : nullptr;
  __nv_bfloat16* op = out + b * out_bstride + ((long long)y_begin * Wo + x) * ldo + cv * 8;
  // The two source rows of the current output row, already interpolated in x, stay in registers: when up-sampling, the
  // RA_YB consecutive output rows of a thread share them (x4: a new source row every fourth output), and the kernel was
  // bound by the L1 traffic of four 32-byte taps per 16-byte output (675 us for the 1.07 GB level against a 336 us
  // HBM floor). out = (1 - ly) * row0 + ly * row1 with row_k = (1 - lx) * src[y_k, x0] + lx * src[y_k, x1].
  float row0[8], row1[8];
  int cy0 = -1, cy1 = -1;                       // source rows held in row0 / row1
  auto load_row = [&](int ys, float* r) {
    const long long ro = sb + (long long)ys * Ws * lds;
    float a[8], c[8];
    ra_load8<SF32>(src, ro + so0, a);
    ra_load8<SF32>(src, ro + so1, c);
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = fmaf(lx, c[j] - a[j], a[j]);
  };
  // the base rows of all RA_YB outputs are requested up front: with one 16-byte load in flight per thread the kernel
  // ran at half the HBM rate (2048 threads x 16 B per SM = 4.8 MB in flight chip-wide against ~1.5 us of loaded latency)
  uint4 gb[RA_YB];
  if (bp) {
#pragma unroll
    for (int k = 0; k < RA_YB; ++k)
      gb[k] = (y_begin + k < y_end) ? __ldg(reinterpret_cast<const uint4*>(bp + (long long)k * Wo * ldb)) : make_uint4(0, 0, 0, 0);
  }
#pragma unroll
  for (int k = 0; k < RA_YB; ++k, op += (long long)Wo * ldo) {
    const int y = y_begin + k;
    if (y >= y_end) break;
    float f[8];
    if (same) {
      ra_load8<SF32>(src, sb + (long long)y * Ws * lds + so0, f);
    } else {
      float sy = (y + 0.5f) * rh - 0.5f;
      sy = sy < 0.f ? 0.f : sy;
      int y0 = (int)sy;
      y0 = y0 > Hs - 1 ? Hs - 1 : y0;
      const int y1 = y0 < Hs - 1 ? y0 + 1 : y0;
      const float ly = sy - y0;
      if (y0 != cy0) {
        if (y0 == cy1) {
#pragma unroll
          for (int j = 0; j < 8; ++j) row0[j] = row1[j];
        } else {
          load_row(y0, row0);
        }
        cy0 = y0;
      }
      if (y1 != cy1) {
        if (y1 == y0) {
#pragma unroll
          for (int j = 0; j < 8; ++j) row1[j] = row0[j];
        } else {
          load_row(y1, row1);
        }
        cy1 = y1;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = fmaf(ly, row1[j] - row0[j], row0[j]);
    }
    if (bp) {
      float g[8];
      unpack8(gb[k], g);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] += g[j];
    }
    if (scale) {
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = fmaf(f[j], sc[j], sh[j]);
    }
    *reinterpret_cast<uint4*>(op) = pack8(f);
  }
}

// out[b,y,x,:] = act((base[b,y,x,:] + sum_k bilinear(src_k[b])[y,x,:]) * scale + shift), dense NHWC bf16, up to 3 sources
// of different resolutions. The Segformer head's fusion conv is linear and so is the bilinear resize in front of it:
//   fusion(concat_i resize(y_i)) = sum_i resize(W_i y_i)      (decode_heads/segformer_head.py:55-64)
// so each level's slice of the fusion weight is applied at the level's OWN resolution and this kernel adds the
// up-sampled partial sums, the folded BatchNorm shift and the ReLU: the [B, H/4 * W/4, 4 * 512] concat is never built.
struct ResizeSumParams {
  const __nv_bfloat16* base;
  const __nv_bfloat16* src[3];
  int Hs[3], Ws[3];
  float rh[3], rw[3];
  int nsrc;
  const float* scale;
  const float* shift;
  __nv_bfloat16* out;
  int B, Ho, Wo, C, relu;
};
// Grid = (x * channel-vector chunks, row blocks of RS_YB rows, batch): a thread keeps its (x, 8 channels) for RS_YB
// consecutive output rows, so the 64-bit index arithmetic (three divisions by run-time sizes per element in a flat
// grid-stride loop: ~600 instructions per element under ncu, issue-bound at 1.5 TB/s), the horizontal corner / weight
// computation of every source and the BatchNorm scale / shift loads are paid once per RS_YB outputs.
static constexpr int RS_YB = 8;
__global__ void __launch_bounds__(256)
resize_sum_affine_kernel(const ResizeSumParams p) {
  const uint32_t CV = (uint32_t)p.C >> 3;
  const uint32_t i = blockIdx.x * 256u + threadIdx.x;
  if (i >= (uint32_t)p.Wo * CV) return;
  const uint32_t x = i / CV, cv = i - x * CV;
  const int b = blockIdx.z;
  const int y_begin = blockIdx.y * RS_YB, y_end = min(y_begin + RS_YB, p.Ho);
  float sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { sc[j] = 1.f; sh[j] = 0.f; }
  if (p.scale) {
    const float4 a0 = __ldg(reinterpret_cast<const float4*>(p.scale + cv * 8)), a1 = __ldg(reinterpret_cast<const float4*>(p.scale + cv * 8 + 4));
    sc[0] = a0.x; sc[1] = a0.y; sc[2] = a0.z; sc[3] = a0.w; sc[4] = a1.x; sc[5] = a1.y; sc[6] = a1.z; sc[7] = a1.w;
  }
  if (p.shift) {
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.shift + cv * 8)), b1 = __ldg(reinterpret_cast<const float4*>(p.shift + cv * 8 + 4));
    sh[0] = b0.x; sh[1] = b0.y; sh[2] = b0.z; sh[3] = b0.w; sh[4] = b1.x; sh[5] = b1.y; sh[6] = b1.z; sh[7] = b1.w;
  }
  // horizontal corners and weights of every source, and the element offset of the two corner columns
  int xo0[3], xo1[3];
  float lxs[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    xo0[k] = xo1[k] = 0; lxs[k] = 0.f;
    if (k < p.nsrc) {
      const int Ws = p.Ws[k];
      float sx = (x + 0.5f) * p.rw[k] - 0.5f;
      sx = sx < 0.f ? 0.f : sx;
      int x0 = (int)sx;
      x0 = x0 > Ws - 1 ? Ws - 1 : x0;
      const int x1 = x0 < Ws - 1 ? x0 + 1 : x0;
      lxs[k] = sx - x0;
      xo0[k] = x0 * p.C + cv * 8;
      xo1[k] = x1 * p.C + cv * 8;
    }
  }
  const long long pix_row = (long long)p.Wo * p.C;
  const long long off0 = ((long long)b * p.Ho + y_begin) * pix_row + (long long)x * p.C + cv * 8;
  const __nv_bfloat16* bp = p.base ? p.base + off0 : nullptr;
  __nv_bfloat16* op = p.out + off0;
  for (int y = y_begin; y < y_end; ++y, op += pix_row) {
    u64 f2[4] = {0ull, 0ull, 0ull, 0ull};
    if (bp) {
      const uint4 bv = __ldg(reinterpret_cast<const uint4*>(bp));
      bp += pix_row;
      f2[0] = pack2(bf16lo(bv.x), bf16hi(bv.x)); f2[1] = pack2(bf16lo(bv.y), bf16hi(bv.y));
      f2[2] = pack2(bf16lo(bv.z), bf16hi(bv.z)); f2[3] = pack2(bf16lo(bv.w), bf16hi(bv.w));
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      if (k < p.nsrc) {
        const int Hs = p.Hs[k], Ws = p.Ws[k];
        float sy = (y + 0.5f) * p.rh[k] - 0.5f;
        sy = sy < 0.f ? 0.f : sy;
        int y0 = (int)sy;
        y0 = y0 > Hs - 1 ? Hs - 1 : y0;
        const int y1 = y0 < Hs - 1 ? y0 + 1 : y0;
        const float ly = sy - y0, lx = lxs[k];
        const __nv_bfloat16* r0 = p.src[k] + ((long long)b * Hs + y0) * Ws * p.C;
        const __nv_bfloat16* r1 = p.src[k] + ((long long)b * Hs + y1) * Ws * p.C;
        const uint4 q00 = __ldg(reinterpret_cast<const uint4*>(r0 + xo0[k]));
        const uint4 q01 = __ldg(reinterpret_cast<const uint4*>(r0 + xo1[k]));
        const uint4 q10 = __ldg(reinterpret_cast<const uint4*>(r1 + xo0[k]));
        const uint4 q11 = __ldg(reinterpret_cast<const uint4*>(r1 + xo1[k]));
        const float w00 = (1.f - ly) * (1.f - lx), w01 = (1.f - ly) * lx, w10 = ly * (1.f - lx), w11 = ly * lx;
        const u64 k00 = pack2(w00, w00), k01 = pack2(w01, w01), k10 = pack2(w10, w10), k11 = pack2(w11, w11);
        const uint32_t* c00 = reinterpret_cast<const uint32_t*>(&q00);
        const uint32_t* c01 = reinterpret_cast<const uint32_t*>(&q01);
        const uint32_t* c10 = reinterpret_cast<const uint32_t*>(&q10);
        const uint32_t* c11 = reinterpret_cast<const uint32_t*>(&q11);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          f2[j] = fma2(pack2(bf16lo(c00[j]), bf16hi(c00[j])), k00, f2[j]);
          f2[j] = fma2(pack2(bf16lo(c01[j]), bf16hi(c01[j])), k01, f2[j]);
          f2[j] = fma2(pack2(bf16lo(c10[j]), bf16hi(c10[j])), k10, f2[j]);
          f2[j] = fma2(pack2(bf16lo(c11[j]), bf16hi(c11[j])), k11, f2[j]);
        }
      }
    }
    float f[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) unpack2(f2[j], f[2 * j], f[2 * j + 1]);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      f[j] = fmaf(f[j], sc[j], sh[j]);
      if (p.relu) f[j] = fmaxf(f[j], 0.f);
    }
    *reinterpret_cast<uint4*>(op) = pack8(f);
  }
}

// Staged variant for the Segformer head: three sources exactly 2x / 4x / 8x smaller than the output, C % 64 == 0.
// The kernel above reads 12 corner vectors per 16-byte output through L1 (13x its output: 586 us at the head's shape, the
// L1 gather ceiling). Caching x-interpolated source rows in registers with run-time row bookkeeping made every row
// change a dependent global load (629 us), and with the windows staged in shared memory the bookkeeping itself cost 325
// instructions per output (593 us, issue-bound under ncu). With integer factors and row tiles aligned to 8 the pattern
// is static: output row i of a tile samples source rows floor((i + 0.5) / F - 0.5) and the next one, relative to the
// tile. So: a CTA owns a tile of 8 rows x 32 columns x 64 channels, stages the source windows it samples in shared
// memory (coordinates clamped to the map: the clamped duplicates reproduce F.interpolate's edge handling), and a
// thread = (column, 8 channels) walks each source's window rows ONCE — two x-interpolated rows in registers, the
// contributions of the output rows between them accumulated with compile-time weights. 3 + 4 + 6 shared-memory row
// fetches per 8 outputs instead of 96 corner loads; the base rows are requested up front.
static constexpr int RT_Y = 8, RT_X = 32, RT_C = 64;
struct ResizeStagedParams {
  ResizeSumParams q;
  int nc[3], off[3];      // window columns / byte offset in shared memory (window rows: RT_Y / F + 2)
};
template <int F>
__device__ __forceinline__ void rs_source(const uint8_t* win, int nc, int xa, int xb, float lx, int cvl, u64 (&acc)[RT_Y][4]) {
  constexpr int NR = RT_Y / F + 2;
  const u64 l1 = pack2(lx, lx), l0 = pack2(1.f - lx, 1.f - lx);
  u64 prev[4], cur[4];
#pragma unroll
  for (int r = 0; r < NR; ++r) {
    const uint8_t* rp = win + (r * nc) * 128 + cvl * 16;
    const uint4 q0 = *reinterpret_cast<const uint4*>(rp + xa * 128);
    const uint4 q1 = *reinterpret_cast<const uint4*>(rp + xb * 128);
    const uint32_t* e0 = reinterpret_cast<const uint32_t*>(&q0);
    const uint32_t* e1 = reinterpret_cast<const uint32_t*>(&q1);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      prev[j] = cur[j];
      cur[j] = fma2(pack2(bf16lo(e1[j]), bf16hi(e1[j])), l1, mul2(pack2(bf16lo(e0[j]), bf16hi(e0[j])), l0));
    }
    if (r > 0) {
      // output rows i with floor((i + 0.5) / F - 0.5) + 1 == r - 1 lie between window rows r - 1 (prev) and r (cur)
#pragma unroll
      for (int i = 0; i < RT_Y; ++i) {
        const int num = 2 * i + 1 + F;                      // (i + 0.5) / F - 0.5 + 1 = num / (2 F)
        if (num / (2 * F) == r - 1) {
          const float ly = (float)(num % (2 * F)) / (float)(2 * F);
          const u64 w1 = pack2(ly, ly), w0 = pack2(1.f - ly, 1.f - ly);
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fma2(cur[j], w1, fma2(prev[j], w0, acc[i][j]));
        }
      }
    }
  }
}
template <int F0, int F1, int F2>
__global__ void __launch_bounds__(256, 2)
resize_sum_staged_kernel(const ResizeStagedParams ps) {
  extern __shared__ __align__(16) uint8_t rs_smem[];
  const ResizeSumParams& p = ps.q;
  const int ncb = p.C / RT_C;
  const int tx = blockIdx.x / ncb, cb = blockIdx.x - tx * ncb;
  const int x_t = tx * RT_X, y_t = blockIdx.y * RT_Y, b = blockIdx.z;
  const int cvl = threadIdx.x & 7, xl = threadIdx.x >> 3;          // 8 channel vectors (128 B) x 32 columns
  const int c0 = cb * RT_C + cvl * 8;
  const int x = x_t + xl;
  constexpr int FS[3] = {F0, F1, F2};
  // ---- stage the windows: rows y_t / F - 1 ... y_t / F + RT_Y / F, columns from the first one the tile samples ----
  int wx0[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    float sx = (x_t + 0.5f) * p.rw[k] - 0.5f;
    sx = sx < 0.f ? 0.f : sx;
    wx0[k] = min((int)sx, p.Ws[k] - 1);
    const int wy0 = y_t / FS[k] - 1;
    const int n = (RT_Y / FS[k] + 2) * ps.nc[k] * 8;
    for (int i = threadIdx.x; i < n; i += 256) {
      const int v = i & 7, px = i >> 3;
      const int ry = px / ps.nc[k], rx = px - ry * ps.nc[k];
      const int sy = min(max(wy0 + ry, 0), p.Hs[k] - 1), sxx = min(wx0[k] + rx, p.Ws[k] - 1);
      const uint4 val = __ldg(reinterpret_cast<const uint4*>(p.src[k] + (((long long)b * p.Hs[k] + sy) * p.Ws[k] + sxx) * p.C + cb * RT_C + v * 8));
      *reinterpret_cast<uint4*>(rs_smem + ps.off[k] + (px * 8 + v) * 16) = val;
    }
  }
  const bool active = x < p.Wo;
  const int xc = active ? x : p.Wo - 1;
  const long long pix_row = (long long)p.Wo * p.C;
  const long long off0 = ((long long)b * p.Ho + y_t) * pix_row + (long long)xc * p.C + c0;
  uint4 gb[RT_Y];
#pragma unroll
  for (int i = 0; i < RT_Y; ++i)
    gb[i] = (p.base && y_t + i < p.Ho) ? __ldg(reinterpret_cast<const uint4*>(p.base + off0 + i * pix_row)) : make_uint4(0, 0, 0, 0);
  // horizontal corners (window-relative) and weights of every source
  int xa[3], xb[3];
  float lxs[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    float sx = (xc + 0.5f) * p.rw[k] - 0.5f;
    sx = sx < 0.f ? 0.f : sx;
    int x0 = (int)sx;
    x0 = x0 > p.Ws[k] - 1 ? p.Ws[k] - 1 : x0;
    const int x1 = x0 < p.Ws[k] - 1 ? x0 + 1 : x0;
    lxs[k] = sx - x0;
    xa[k] = x0 - wx0[k]; xb[k] = x1 - wx0[k];
  }
  __syncthreads();
  u64 acc[RT_Y][4];
#pragma unroll
  for (int i = 0; i < RT_Y; ++i) {
    const uint4 bv = gb[i];
    acc[i][0] = pack2(bf16lo(bv.x), bf16hi(bv.x)); acc[i][1] = pack2(bf16lo(bv.y), bf16hi(bv.y));
    acc[i][2] = pack2(bf16lo(bv.z), bf16hi(bv.z)); acc[i][3] = pack2(bf16lo(bv.w), bf16hi(bv.w));
  }
  rs_source<F0>(rs_smem + ps.off[0], ps.nc[0], xa[0], xb[0], lxs[0], cvl, acc);
  rs_source<F1>(rs_smem + ps.off[1], ps.nc[1], xa[1], xb[1], lxs[1], cvl, acc);
  rs_source<F2>(rs_smem + ps.off[2], ps.nc[2], xa[2], xb[2], lxs[2], cvl, acc);
  float sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { sc[j] = 1.f; sh[j] = 0.f; }
  if (p.scale) {
    const float4 a0 = __ldg(reinterpret_cast<const float4*>(p.scale + c0)), a1 = __ldg(reinterpret_cast<const float4*>(p.scale + c0 + 4));
    sc[0] = a0.x; sc[1] = a0.y; sc[2] = a0.z; sc[3] = a0.w; sc[4] = a1.x; sc[5] = a1.y; sc[6] = a1.z; sc[7] = a1.w;
  }
  if (p.shift) {
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.shift + c0)), b1 = __ldg(reinterpret_cast<const float4*>(p.shift + c0 + 4));
    sh[0] = b0.x; sh[1] = b0.y; sh[2] = b0.z; sh[3] = b0.w; sh[4] = b1.x; sh[5] = b1.y; sh[6] = b1.z; sh[7] = b1.w;
  }
  __nv_bfloat16* op = p.out + off0;
#pragma unroll
  for (int i = 0; i < RT_Y; ++i, op += pix_row) {
    if (y_t + i >= p.Ho) break;
    float f[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) unpack2(acc[i][j], f[2 * j], f[2 * j + 1]);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      f[j] = fmaf(f[j], sc[j], sh[j]);
      if (p.relu) f[j] = fmaxf(f[j], 0.f);
    }
    if (active) *reinterpret_cast<uint4*>(op) = pack8(f);
  }
}

// labels[b,y,x] = argmax_c bilinear(logits[b])[y,x,c]  (first maximum wins, like torch.argmax).
// Replaces resize(logits -> image size) + softmax + argmax of EncoderDecoder.encode_decode_test /
// whole_inference_dim(_cut) / simple_test (encoder_decoder.py:96-117, 329-414, 449, 477); softmax is
// monotone so it is skipped; the crop of whole_inference_dim_cut is the (Hc, Wc) output window.
__global__ void __launch_bounds__(256)
upsample_argmax_kernel(const float* __restrict__ logits, uint8_t* __restrict__ labels, int B, int hs, int ws,
                       int ldl, int ncls, int Ho, int Wo, int Hc, int Wc, float rh, float rw) {
  const long long total = (long long)B * Hc * Wc;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(idx % Wc);
    long long t = idx / Wc;
    const int y = (int)(t % Hc);
    const int b = (int)(t / Hc);
    float sy = (y + 0.5f) * rh - 0.5f, sx = (x + 0.5f) * rw - 0.5f;
    sy = sy < 0.f ? 0.f : sy;
    sx = sx < 0.f ? 0.f : sx;
    int y0 = (int)sy, x0 = (int)sx;
    y0 = y0 > hs - 1 ? hs - 1 : y0;
    x0 = x0 > ws - 1 ? ws - 1 : x0;
    const int y1 = y0 < hs - 1 ? y0 + 1 : y0, x1 = x0 < ws - 1 ? x0 + 1 : x0;
    const float ly = sy - y0, lx = sx - x0;
    const float w00 = (1.f - ly) * (1.f - lx), w01 = (1.f - ly) * lx, w10 = ly * (1.f - lx), w11 = ly * lx;
    const float* lb = logits + (long long)b * hs * ws * ldl;
    const float* p00 = lb + ((long long)y0 * ws + x0) * ldl;
    const float* p01 = lb + ((long long)y0 * ws + x1) * ldl;
    const float* p10 = lb + ((long long)y1 * ws + x0) * ldl;
    const float* p11 = lb + ((long long)y1 * ws + x1) * ldl;
    float best = -INFINITY;
    int arg = 0;
    for (int c = 0; c < ncls; c += 4) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(p00 + c));
      const float4 bq = __ldg(reinterpret_cast<const float4*>(p01 + c));
      const float4 cq = __ldg(reinterpret_cast<const float4*>(p10 + c));
      const float4 d = __ldg(reinterpret_cast<const float4*>(p11 + c));
      const float v0 = w00 * a.x + w01 * bq.x + w10 * cq.x + w11 * d.x;
      const float v1 = w00 * a.y + w01 * bq.y + w10 * cq.y + w11 * d.y;
      const float v2 = w00 * a.z + w01 * bq.z + w10 * cq.z + w11 * d.z;
      const float v3 = w00 * a.w + w01 * bq.w + w10 * cq.w + w11 * d.w;
      if (v0 > best) { best = v0; arg = c; }
      if (c + 1 < ncls && v1 > best) { best = v1; arg = c + 1; }
      if (c + 2 < ncls && v2 > best) { best = v2; arg = c + 2; }
      if (c + 3 < ncls && v3 > best) { best = v3; arg = c + 3; }
    }
    labels[((long long)b * Hc + y) * Wc + x] = (uint8_t)arg;
  }
}

// Same result, source window staged in shared memory. The flat kernel above is bound by L1 wavefronts: every one of its
// 28 16-byte loads per pixel touches ~9 distinct source pixels per warp. Here a CTA owns UA_RY output rows x 256 output
// columns, copies the few source rows / columns they interpolate from (coalesced, pixel pitch padded to 144 B so that the
// ~9 source pixels a warp reads side by side fall into different banks) and every thread walks its column down the rows.
static constexpr int UA_RY = 4, UA_PITCH = 36;   // floats per staged pixel (32 logits + 4 pad)
__global__ void __launch_bounds__(256)
upsample_argmax_staged_kernel(const float* __restrict__ logits, uint8_t* __restrict__ labels, int hs, int ws, int ldl,
                              int ncls, int Hc, int Wc, float rh, float rw, int cols_s, int rows_s) {
  extern __shared__ __align__(16) float ua_s[];     // [rows_s][cols_s][UA_PITCH]
  const int b = blockIdx.z;
  const int y_begin = blockIdx.y * UA_RY, y_end = min(y_begin + UA_RY, Hc);
  const int x_begin = blockIdx.x * 256;
  // first source row / column any pixel of the tile reads (the clamped floor of its source coordinate)
  float fy = (y_begin + 0.5f) * rh - 0.5f, fx = (x_begin + 0.5f) * rw - 0.5f;
  fy = fy < 0.f ? 0.f : fy;
  fx = fx < 0.f ? 0.f : fx;
  const int ys0 = min((int)fy, hs - 1), xs0 = min((int)fx, ws - 1);
  const int nvec = ldl >> 2;                         // float4 per source pixel (ldl <= 32)
  const float* lb = logits + (long long)b * hs * ws * ldl;
  for (int i = threadIdx.x; i < rows_s * cols_s * nvec; i += 256) {
    const int v = i % nvec, px = i / nvec;
    const int cx = px % cols_s, cy = px / cols_s;
    const int sy = min(ys0 + cy, hs - 1), sx = min(xs0 + cx, ws - 1);
    *reinterpret_cast<float4*>(ua_s + (cy * cols_s + cx) * UA_PITCH + v * 4) =
        __ldg(reinterpret_cast<const float4*>(lb + ((long long)sy * ws + sx) * ldl) + v);
  }
  __syncthreads();
  const int x = x_begin + threadIdx.x;
  if (x >= Wc) return;
  float sxf = (x + 0.5f) * rw - 0.5f;
  sxf = sxf < 0.f ? 0.f : sxf;
  int x0 = (int)sxf;
  x0 = x0 > ws - 1 ? ws - 1 : x0;
  const int x1 = x0 < ws - 1 ? x0 + 1 : x0;
  const float lx = sxf - x0;
  const int cx0 = x0 - xs0, cx1 = x1 - xs0;
  for (int y = y_begin; y < y_end; ++y) {
    float syf = (y + 0.5f) * rh - 0.5f;
    syf = syf < 0.f ? 0.f : syf;
    int y0 = (int)syf;
    y0 = y0 > hs - 1 ? hs - 1 : y0;
    const int y1 = y0 < hs - 1 ? y0 + 1 : y0;
    const float ly = syf - y0;
    const float w00 = (1.f - ly) * (1.f - lx), w01 = (1.f - ly) * lx, w10 = ly * (1.f - lx), w11 = ly * lx;
    const float* p00 = ua_s + ((y0 - ys0) * cols_s + cx0) * UA_PITCH;
    const float* p01 = ua_s + ((y0 - ys0) * cols_s + cx1) * UA_PITCH;
    const float* p10 = ua_s + ((y1 - ys0) * cols_s + cx0) * UA_PITCH;
    const float* p11 = ua_s + ((y1 - ys0) * cols_s + cx1) * UA_PITCH;
    float best = -INFINITY;
    int arg = 0;
    for (int c = 0; c < ncls; c += 4) {
      const float4 a = *reinterpret_cast<const float4*>(p00 + c);
      const float4 bq = *reinterpret_cast<const float4*>(p01 + c);
      const float4 cq = *reinterpret_cast<const float4*>(p10 + c);
      const float4 d = *reinterpret_cast<const float4*>(p11 + c);
      const float v0 = w00 * a.x + w01 * bq.x + w10 * cq.x + w11 * d.x;
      const float v1 = w00 * a.y + w01 * bq.y + w10 * cq.y + w11 * d.y;
      const float v2 = w00 * a.z + w01 * bq.z + w10 * cq.z + w11 * d.z;
      const float v3 = w00 * a.w + w01 * bq.w + w10 * cq.w + w11 * d.w;
      if (v0 > best) { best = v0; arg = c; }
      if (c + 1 < ncls && v1 > best) { best = v1; arg = c + 1; }
      if (c + 2 < ncls && v2 > best) { best = v2; arg = c + 2; }
      if (c + 3 < ncls && v3 > best) { best = v3; arg = c + 3; }
    }
    labels[((long long)b * Hc + y) * Wc + x] = (uint8_t)arg;
  }
}

// confusion[gt, pred] += 1 over all pixels with gt != ignore (the device-side form of
// intersect_and_union, mmseg_custom/apis/evaluation/metrics_micro.py:26-86).
__global__ void __launch_bounds__(256)
confusion_kernel(const uint8_t* __restrict__ pred, const uint8_t* __restrict__ gt, unsigned long long* conf,
                 long long n, int ncls, int ignore) {
  extern __shared__ unsigned int s_conf[];
  for (int i = threadIdx.x; i < ncls * ncls; i += blockDim.x) s_conf[i] = 0;
  __syncthreads();
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n;
       idx += (long long)gridDim.x * blockDim.x) {
    const int g = gt[idx], p = pred[idx];
    if (g != ignore && g < ncls && p < ncls) atomicAdd(&s_conf[g * ncls + p], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < ncls * ncls; i += blockDim.x)
    if (s_conf[i]) atomicAdd(&conf[i], (unsigned long long)s_conf[i]);
}

static inline unsigned grid_for(long long total, int per_block = 256, int waves = 8) {
  long long b = (total + per_block - 1) / per_block;
  const long long cap = (long long)kNumSMs * waves;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (unsigned)b;
}

}  // namespace mmsam

MMSAM_API int mmsam_patchify_f32(const float* img, void* out, int B, int Ctot, int c_off, int C, int H, int W,
                                 int p, void* stream) {
  using namespace mmsam;
  if (B < 0 || C <= 0 || p <= 0 || H % p || W % p || c_off < 0 || c_off + C > Ctot) return MMSAM_ERR_BAD_ARG;
  if (B == 0) return MMSAM_OK;
  if (!img || !out) return MMSAM_ERR_BAD_ARG;
  const long long total = (long long)B * C * H * W;
  const bool vec = (((uintptr_t)img) & 15) == 0 && (((uintptr_t)out) & 7) == 0 && (W & 3) == 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (vec && p == 4) patchify_kernel<4><<<grid_for(total / 4), 256, 0, st>>>(img, (__nv_bfloat16*)out, B, Ctot, c_off, C, H, W);
  else if (vec && p == 16) patchify_kernel<16><<<grid_for(total / 16), 256, 0, st>>>(img, (__nv_bfloat16*)out, B, Ctot, c_off, C, H, W);
  else patchify_generic_kernel<<<grid_for(total), 256, 0, st>>>(img, (__nv_bfloat16*)out, B, Ctot, c_off, C, H, W, p);
  MMSAM_LAUNCH_CHECK();
  return MMSAM_OK;
}

MMSAM_API int mmsam_normalize_u8(const void* img_hwc_u8, float* out_nchw, int B, int H, int W, int C, int Ctot, int c_off,
                                 const float* mean_host, const float* std_host, float prescale, void* stream) {
  using namespace mmsam;
  if (B < 0 || H <= 0 || W <= 0 || C <= 0 || C > 4 || (W & 3) || c_off < 0 || c_off + C > Ctot) return MMSAM_ERR_BAD_ARG;
  if (B == 0) return MMSAM_OK;
  if (!img_hwc_u8 || !out_nchw || !mean_host || !std_host || (((uintptr_t)out_nchw) & 15)) return MMSAM_ERR_BAD_ARG;
  float m[4] = {0, 0, 0, 0}, r[4] = {1, 1, 1, 1};
  for (int c = 0; c < C; ++c) {
    if (std_host[c] == 0.f) return MMSAM_ERR_BAD_ARG;
    m[c] = mean_host[c];
    r[c] = 1.f / std_host[c];
  }
  const long long total = (long long)B * H * (W / 4);
  normalize_u8_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>((const uint8_t*)img_hwc_u8, out_nchw, B, H, W, C, Ctot, c_off,
                                                                       m[0], m[1], m[2], m[3], r[0], r[1], r[2], r[3], prescale);
  MMSAM_LAUNCH_CHECK();
  return MMSAM_OK;
}

MMSAM_API int mmsam_resize_add_affine(const void* src, int src_dtype, const void* base, const float* scale,
                                           const float* shift, void* out, int B, int Hs, int Ws, int Ho, int Wo,
                                           int C, long long src_bstride, long long lds, long long base_bstride,
                                           long long ldb, long long out_bstride, long long ldo, void* stream) {
  using namespace mmsam;
  if (B < 0 || C <= 0 || (C & 7) || Hs <= 0 || Ws <= 0 || Ho <= 0 || Wo <= 0) return MMSAM_ERR_BAD_ARG;
  if ((lds | ldb | ldo | src_bstride | base_bstride | out_bstride) & 7) return MMSAM_ERR_BAD_ARG;
  if ((scale == nullptr) != (shift == nullptr)) return MMSAM_ERR_BAD_ARG;
  if (src_dtype != MMSAM_BF16 && src_dtype != MMSAM_F32) return MMSAM_ERR_BAD_DTYPE;
  if (B == 0) return MMSAM_OK;
  if (!src || !out) return MMSAM_ERR_BAD_ARG;
  if ((((uintptr_t)src | (uintptr_t)out | (uintptr_t)base | (uintptr_t)scale | (uintptr_t)shift) & 15)) return MMSAM_ERR_BAD_ARG;
  if (B > 65535 || (Ho + RA_YB - 1) / RA_YB > 65535 || (long long)Wo * (C / 8) > (1ll << 31) - 256) return MMSAM_ERR_UNSUPPORTED;
  dim3 grid((unsigned)(((long long)Wo * (C / 8) + 255) / 256), (unsigned)((Ho + RA_YB - 1) / RA_YB), (unsigned)B);
  if (src_dtype == MMSAM_F32)
    resize_add_affine_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(
        src, (const __nv_bfloat16*)base, scale, shift, (__nv_bfloat16*)out, B, Hs, Ws, Ho, Wo, C,
        src_bstride, base_bstride, out_bstride, ldo, lds, ldb, (float)Hs / (float)Ho, (float)Ws / (float)Wo);
  else
    resize_add_affine_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(
        src, (const __nv_bfloat16*)base, scale, shift, (__nv_bfloat16*)out, B, Hs, Ws, Ho, Wo, C,
        src_bstride, base_bstride, out_bstride, ldo, lds, ldb, (float)Hs / (float)Ho, (float)Ws / (float)Wo);
  MMSAM_LAUNCH_CHECK();
  return MMSAM_OK;
}

MMSAM_API int mmsam_resize_sum_affine_bf16(const void* base, int nsrc, const void* src0, const void* src1,
                                           const void* src2, const int* src_hw_host, const float* scale,
                                           const float* shift, int relu, void* out, int B, int Ho, int Wo, int C,
                                           void* stream) {
  using namespace mmsam;
  if (B < 0 || C <= 0 || (C & 7) || Ho <= 0 || Wo <= 0 || nsrc < 0 || nsrc > 3) return MMSAM_ERR_BAD_ARG;
  if (B == 0) return MMSAM_OK;
  if (!out || (nsrc > 0 && !src_hw_host) || (!base && nsrc == 0)) return MMSAM_ERR_BAD_ARG;
  const void* srcs[3] = {src0, src1, src2};
  ResizeSumParams p;
  p.base = (const __nv_bfloat16*)base; p.nsrc = nsrc; p.scale = scale; p.shift = shift; p.out = (__nv_bfloat16*)out;
  p.B = B; p.Ho = Ho; p.Wo = Wo; p.C = C; p.relu = relu;
  uintptr_t align = (uintptr_t)base | (uintptr_t)out | (uintptr_t)scale | (uintptr_t)shift;
  for (int k = 0; k < 3; ++k) {
    p.src[k] = nullptr; p.Hs[k] = p.Ws[k] = 1; p.rh[k] = p.rw[k] = 1.f;
    if (k < nsrc) {
      if (!srcs[k]) return MMSAM_ERR_BAD_ARG;
      p.src[k] = (const __nv_bfloat16*)srcs[k];
      p.Hs[k] = src_hw_host[2 * k]; p.Ws[k] = src_hw_host[2 * k + 1];
      if (p.Hs[k] <= 0 || p.Ws[k] <= 0) return MMSAM_ERR_BAD_ARG;
      p.rh[k] = (float)p.Hs[k] / (float)Ho; p.rw[k] = (float)p.Ws[k] / (float)Wo;
      align |= (uintptr_t)srcs[k];
    }
  }
  if (align & 15) return MMSAM_ERR_BAD_ARG;
  if (B > 65535 || (Ho + RS_YB - 1) / RS_YB > 65535 || (long long)Wo * (C / 8) > (1ll << 31) - 256 ||
      (long long)Wo * C >= (1ll << 31))
    return MMSAM_ERR_UNSUPPORTED;
  if (nsrc == 3 && base && (C % RT_C) == 0 && Ho == 2 * p.Hs[0] && Wo == 2 * p.Ws[0] && Ho == 4 * p.Hs[1] && Wo == 4 * p.Ws[1] &&
      Ho == 8 * p.Hs[2] && Wo == 8 * p.Ws[2] && getenv("MMSAM_RESIZE_SUM_LEGACY") == nullptr) {
    // the Segformer head: sources exactly 2x / 4x / 8x smaller than the output (staged kernel with static row patterns)
    ResizeStagedParams ps;
    ps.q = p;
    int off = 0;
    const int fs[3] = {2, 4, 8};
    for (int k = 0; k < 3; ++k) {
      ps.nc[k] = (RT_X - 1) / fs[k] + 3;        // columns a 32-wide tile can touch
      ps.off[k] = off;
      off += (RT_Y / fs[k] + 2) * ps.nc[k] * 128;
    }
    dim3 grid((unsigned)(((Wo + RT_X - 1) / RT_X) * (C / RT_C)), (unsigned)((Ho + RT_Y - 1) / RT_Y), (unsigned)B);
    resize_sum_staged_kernel<2, 4, 8><<<grid, 256, off, (cudaStream_t)stream>>>(ps);
    MMSAM_LAUNCH_CHECK();
    return MMSAM_OK;
  }
  dim3 grid((unsigned)(((long long)Wo * (C / 8) + 255) / 256), (unsigned)((Ho + RS_YB - 1) / RS_YB), (unsigned)B);
  resize_sum_affine_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p);
  MMSAM_LAUNCH_CHECK();
  return MMSAM_OK;
}

MMSAM_API int mmsam_upsample_argmax_f32(const float* logits, void* labels_u8, int B, int hs, int ws, int ldl,
                                        int ncls, int Ho, int Wo, int Hc, int Wc, void* stream) {
  using namespace mmsam;
  if (B < 0 || hs <= 0 || ws <= 0 || ncls <= 0 || ncls > 256 || (ldl & 3) || ldl < ((ncls + 3) & ~3) || Ho <= 0 ||
      Wo <= 0 || Hc <= 0 || Wc <= 0 || Hc > Ho || Wc > Wo)
    return MMSAM_ERR_BAD_ARG;
  if (B == 0) return MMSAM_OK;
  if (!logits || !labels_u8 || (((uintptr_t)logits) & 15)) return MMSAM_ERR_BAD_ARG;
  const long long total = (long long)B * Hc * Wc;
  const float rh = (float)hs / (float)Ho, rw = (float)ws / (float)Wo;
  // staged variant: the source window of a (UA_RY rows x 256 columns) output tile must fit 48 KB of shared memory
  const int cols_s = (int)ceilf(256.f * rw) + 2, rows_s = (int)ceilf((float)UA_RY * rh) + 2;
  const size_t smem = (size_t)cols_s * rows_s * UA_PITCH * sizeof(float);
  if (ldl <= 32 && smem <= 48 * 1024 && B <= 65535 && (Hc + UA_RY - 1) / UA_RY <= 65535 && !getenv("MMSAM_ARGMAX_FLAT")) {
    dim3 grid((unsigned)((Wc + 255) / 256), (unsigned)((Hc + UA_RY - 1) / UA_RY), (unsigned)B);
    upsample_argmax_staged_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(logits, (uint8_t*)labels_u8, hs, ws, ldl, ncls, Hc,
                                                                            Wc, rh, rw, cols_s, rows_s);
  } else {
    upsample_argmax_kernel<<<grid_for(total, 256, 16), 256, 0, (cudaStream_t)stream>>>(
        logits, (uint8_t*)labels_u8, B, hs, ws, ldl, ncls, Ho, Wo, Hc, Wc, rh, rw);
  }
  MMSAM_LAUNCH_CHECK();
  return MMSAM_OK;
}

MMSAM_API int mmsam_confusion_u8(const void* pred_u8, const void* gt_u8, void* conf_u64, long long n, int ncls,
                                 int ignore_index, void* stream) {
  using namespace mmsam;
  if (n < 0 || ncls <= 0 || ncls > 64) return MMSAM_ERR_BAD_ARG;
  if (n == 0) return MMSAM_OK;
  if (!pred_u8 || !gt_u8 || !conf_u64) return MMSAM_ERR_BAD_ARG;
  confusion_kernel<<<grid_for(n, 256 * 16, 4), 256, ncls * ncls * 4, (cudaStream_t)stream>>>(
      (const uint8_t*)pred_u8, (const uint8_t*)gt_u8, (unsigned long long*)conf_u64, n, ncls, ignore_index);
  MMSAM_LAUNCH_CHECK();
  return MMSAM_OK;
}
