// The callers on either side of the encoder: the uint8 input pipeline fused into the stem / patch-embed
// operand, and the segmentor's post-processing (chained logits resizes, sliding-window overlap-add + argmax). HBM-bound
// streaming kernels.
#include "common.cuh"

namespace mmsam {

// ------------------------------------------------------------------------------------------------
// Decoded image, HWC uint8 [B, Hs, Ws, C]  ->  patch-major bf16 rows [(b, py, px), (c, ky, kx)] of the p x p / stride p
// conv that consumes it (ConvNeXt stem 4x4, ViT patch embed 16x16), normalised on the fly:
//     pad to H x W with pad_val (Pad_multimodal, pipelines/transform.py:2934-..., BEFORE the normalisation, as in the
//     test pipelines)  ->  (v * prescale - mean[c]) / std[c]  (Normalize_multimodal, :2717-2815; norm_by_max: 1/255;
//     to_rgb: channel order reversed, mmcv.imnormalize)  ->  ImageToTensor's HWC -> CHW is the (c, ky, kx) column order.
// The fp32 NCHW image never exists: the host ships 1 byte per value instead of 4.
// Thread = (b, py, px, ky): p pixels x C bytes in (contiguous), C runs of p bf16 out.
// ------------------------------------------------------------------------------------------------
struct U8Norm {
  float mean[4], rstd[4];
  float prescale, pad_val;
  int swap_rb;
};

template <int P>
__global__ void __launch_bounds__(256)
patchify_u8_kernel(const uint8_t* __restrict__ img, __nv_bfloat16* __restrict__ out, int B, int Hs, int Ws, int C, int H,
                   int W, const U8Norm nm) {
  const int PW = W / P, PH = H / P;
  const long long total = (long long)B * PH * PW * P;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int ky = (int)(idx % P);
    long long t = idx / P;
    const int px = (int)(t % PW); t /= PW;
    const int py = (int)(t % PH);
    const int b = (int)(t / PH);
    const int y = py * P + ky, x0 = px * P;
    __nv_bfloat16* dst = out + (((long long)b * PH + py) * PW + px) * (C * P * P) + ky * P;
    const uint8_t* src = img + (((long long)b * Hs + y) * Ws + x0) * C;
    for (int c = 0; c < C; ++c) {
      const int cs = nm.swap_rb ? C - 1 - c : c;          // source channel of output channel c
      float v[P];
#pragma unroll
      for (int kx = 0; kx < P; ++kx) {
        const float raw = (y < Hs && x0 + kx < Ws) ? (float)src[kx * C + cs] : nm.pad_val;
        v[kx] = (raw * nm.prescale - nm.mean[c]) * nm.rstd[c];
      }
      __nv_bfloat16* d = dst + c * P * P;
#pragma unroll
      for (int kx = 0; kx < P; kx += 2) *reinterpret_cast<uint32_t*>(d + kx) = pack_bf16(v[kx], v[kx + 1]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Bilinear resize (align_corners=False) of channels-last fp32 logits [B, hs, ws, ld] -> [B, Ho, Wo, ld]: the extra
// resize steps of EncoderDecoder (encode_decode -> image size, then whole_inference -> ori_shape or
// whole_inference_dim -> test_cfg.dim, encoder_decoder.py:103-107, 317-325, 341-346) when they are not identities.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
resize_f32_kernel(const float* __restrict__ src, float* __restrict__ dst, int B, int hs, int ws, int ld, int Ho, int Wo,
                  float rh, float rw) {
  const int nv = ld >> 2;
  const long long total = (long long)B * Ho * Wo * nv;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(idx % nv);
    long long t = idx / nv;
    const int x = (int)(t % Wo); t /= Wo;
    const int y = (int)(t % Ho);
    const int b = (int)(t / Ho);
    float sy = (y + 0.5f) * rh - 0.5f, sx = (x + 0.5f) * rw - 0.5f;
    sy = sy < 0.f ? 0.f : sy;
    sx = sx < 0.f ? 0.f : sx;
    int y0 = (int)sy, x0 = (int)sx;
    y0 = y0 > hs - 1 ? hs - 1 : y0;
    x0 = x0 > ws - 1 ? ws - 1 : x0;
    const int y1 = y0 < hs - 1 ? y0 + 1 : y0, x1 = x0 < ws - 1 ? x0 + 1 : x0;
    const float ly = sy - y0, lx = sx - x0;
    const float w00 = (1.f - ly) * (1.f - lx), w01 = (1.f - ly) * lx, w10 = ly * (1.f - lx), w11 = ly * lx;
    const float* sb = src + (long long)b * hs * ws * ld + v * 4;
    const float4 a = __ldg(reinterpret_cast<const float4*>(sb + ((long long)y0 * ws + x0) * ld));
    const float4 c = __ldg(reinterpret_cast<const float4*>(sb + ((long long)y0 * ws + x1) * ld));
    const float4 d = __ldg(reinterpret_cast<const float4*>(sb + ((long long)y1 * ws + x0) * ld));
    const float4 e = __ldg(reinterpret_cast<const float4*>(sb + ((long long)y1 * ws + x1) * ld));
    float4 o;
    o.x = w00 * a.x + w01 * c.x + w10 * d.x + w11 * e.x;
    o.y = w00 * a.y + w01 * c.y + w10 * d.y + w11 * e.y;
    o.z = w00 * a.z + w01 * c.z + w10 * d.z + w11 * e.z;
    o.w = w00 * a.w + w01 * c.w + w10 * d.w + w11 * e.w;
    *reinterpret_cast<float4*>(dst + idx * 4) = o;
  }
}

// ------------------------------------------------------------------------------------------------
// slide_inference (encoder_decoder.py:191-234) after the head, in one pass over the frame: for every frame pixel the
// logits of each crop covering it are bilinearly up-sampled from the crop's head resolution (resize to the crop size,
// :103-107), summed (preds += F.pad(crop_seg_logit)), divided by the overlap count (:222) and arg-maxed (:477) — or, when a
// rescale to ori_shape follows (:223-229), written out as fp32 [B, H, W, ld] for the resize kernel above.
// Crop j of image b is logits[(j * B + b)] (the crops of a frame are batched position-major through the network).
// Thread = frame pixel; <= 32 classes (accumulated in registers); <= 16 crop positions.
// ------------------------------------------------------------------------------------------------
struct SlideParams {
  int ncrops;
  int y1[16], x1[16], ch[16], cw[16];   // crop box origin and size in frame pixels
};

__global__ void __launch_bounds__(256)
slide_merge_kernel(const float* __restrict__ logits, uint8_t* __restrict__ labels, float* __restrict__ preds, int B, int hs,
                   int ws, int ld, int ncls, int H, int W, const SlideParams sp) {
  const long long total = (long long)B * H * W;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(idx % W);
    long long t = idx / W;
    const int y = (int)(t % H);
    const int b = (int)(t / H);
    float acc[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) acc[c] = 0.f;
    int count = 0;
    for (int j = 0; j < sp.ncrops; ++j) {
      const int yy = y - sp.y1[j], xx = x - sp.x1[j];
      if (yy < 0 || xx < 0 || yy >= sp.ch[j] || xx >= sp.cw[j]) continue;
      ++count;
      float sy = (yy + 0.5f) * ((float)hs / (float)sp.ch[j]) - 0.5f, sx = (xx + 0.5f) * ((float)ws / (float)sp.cw[j]) - 0.5f;
      sy = sy < 0.f ? 0.f : sy;
      sx = sx < 0.f ? 0.f : sx;
      int y0 = (int)sy, x0 = (int)sx;
      y0 = y0 > hs - 1 ? hs - 1 : y0;
      x0 = x0 > ws - 1 ? ws - 1 : x0;
      const int y1 = y0 < hs - 1 ? y0 + 1 : y0, x1 = x0 < ws - 1 ? x0 + 1 : x0;
      const float ly = sy - y0, lx = sx - x0;
      const float w00 = (1.f - ly) * (1.f - lx), w01 = (1.f - ly) * lx, w10 = ly * (1.f - lx), w11 = ly * lx;
      const float* lb = logits + (long long)(j * B + b) * hs * ws * ld;
      const float* p00 = lb + ((long long)y0 * ws + x0) * ld;
      const float* p01 = lb + ((long long)y0 * ws + x1) * ld;
      const float* p10 = lb + ((long long)y1 * ws + x0) * ld;
      const float* p11 = lb + ((long long)y1 * ws + x1) * ld;
#pragma unroll
      for (int c = 0; c < 32; c += 4) {
        if (c >= ncls) break;
        const float4 a = __ldg(reinterpret_cast<const float4*>(p00 + c));
        const float4 bq = __ldg(reinterpret_cast<const float4*>(p01 + c));
        const float4 cq = __ldg(reinterpret_cast<const float4*>(p10 + c));
        const float4 d = __ldg(reinterpret_cast<const float4*>(p11 + c));
        // the crop's resized logit first (one value, as F.interpolate produces it), then the sum over crops
        acc[c] += w00 * a.x + w01 * bq.x + w10 * cq.x + w11 * d.x;
        acc[c + 1] += w00 * a.y + w01 * bq.y + w10 * cq.y + w11 * d.y;
        acc[c + 2] += w00 * a.z + w01 * bq.z + w10 * cq.z + w11 * d.z;
        acc[c + 3] += w00 * a.w + w01 * bq.w + w10 * cq.w + w11 * d.w;
      }
    }
    // preds / count_mat is a true division in the reference (:222; every pixel is covered: count >= 1, :219)
    const float cnt = (float)(count > 0 ? count : 1);
#pragma unroll
    for (int c = 0; c < 32; ++c) acc[c] = acc[c] / cnt;
    if (preds) {
      float* o = preds + idx * ld;
#pragma unroll
      for (int c = 0; c < 32; c += 4) {
        if (c >= ld) break;
        *reinterpret_cast<float4*>(o + c) = make_float4(acc[c], acc[c + 1], acc[c + 2], acc[c + 3]);
      }
    }
    if (labels) {
      float best = -INFINITY;
      int arg = 0;
#pragma unroll
      for (int c = 0; c < 32; ++c)
        if (c < ncls && acc[c] > best) { best = acc[c]; arg = c; }
      labels[idx] = (uint8_t)arg;
    }
  }
}

static inline unsigned pl_grid(long long total, int waves = 16) {
  long long b = (total + 255) / 256;
  const long long cap = (long long)kNumSMs * waves;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (unsigned)b;
}

}  // namespace mmsam

// See include/mmsam_b200.h for the contracts.
MMSAM_API int mmsam_patchify_u8(const void* img_hwc_u8, void* out, int B, int Hs, int Ws, int C, int H, int W, int p,
                                const float* mean_host, const float* std_host, float prescale, float pad_val, int swap_rb,
                                void* stream) {
  using namespace mmsam;
  if (B < 0 || C <= 0 || C > 4 || p <= 0 || H % p || W % p || Hs <= 0 || Ws <= 0 || Hs > H || Ws > W) return MMSAM_ERR_BAD_ARG;
  if (p != 4 && p != 16) return MMSAM_ERR_UNSUPPORTED;
  if (B == 0) return MMSAM_OK;
  if (!img_hwc_u8 || !out || !mean_host || !std_host || (((uintptr_t)out) & 3)) return MMSAM_ERR_BAD_ARG;
  U8Norm nm;
  for (int c = 0; c < 4; ++c) {
    nm.mean[c] = c < C ? mean_host[c] : 0.f;
    nm.rstd[c] = c < C ? 1.f / std_host[c] : 1.f;
  }
  nm.prescale = prescale; nm.pad_val = pad_val; nm.swap_rb = swap_rb;
  const long long total = (long long)B * (H / p) * (W / p) * p;
  cudaStream_t st = (cudaStream_t)stream;
  if (p == 4) patchify_u8_kernel<4><<<pl_grid(total), 256, 0, st>>>((const uint8_t*)img_hwc_u8, (__nv_bfloat16*)out, B, Hs, Ws, C, H, W, nm);
  else patchify_u8_kernel<16><<<pl_grid(total), 256, 0, st>>>((const uint8_t*)img_hwc_u8, (__nv_bfloat16*)out, B, Hs, Ws, C, H, W, nm);
  MMSAM_LAUNCH_CHECK();
  return MMSAM_OK;
}

MMSAM_API int mmsam_resize_logits_f32(const float* src, float* dst, int B, int hs, int ws, int ld, int Ho, int Wo, void* stream) {
  using namespace mmsam;
  if (B < 0 || hs <= 0 || ws <= 0 || ld <= 0 || (ld & 3) || Ho <= 0 || Wo <= 0) return MMSAM_ERR_BAD_ARG;
  if (B == 0) return MMSAM_OK;
  if (!src || !dst || (((uintptr_t)src | (uintptr_t)dst) & 15)) return MMSAM_ERR_BAD_ARG;
  resize_f32_kernel<<<pl_grid((long long)B * Ho * Wo * (ld / 4)), 256, 0, (cudaStream_t)stream>>>(
      src, dst, B, hs, ws, ld, Ho, Wo, (float)hs / (float)Ho, (float)ws / (float)Wo);
  MMSAM_LAUNCH_CHECK();
  return MMSAM_OK;
}

MMSAM_API int mmsam_slide_merge_f32(const float* crop_logits, void* labels_u8, float* preds, int B, int hs, int ws, int ld,
                                    int ncls, int H, int W, int ncrops, const int* crop_boxes_host, void* stream) {
  using namespace mmsam;
  if (B < 0 || hs <= 0 || ws <= 0 || ld <= 0 || (ld & 3) || ld > 32 || ncls <= 0 || ncls > ld || H <= 0 || W <= 0) return MMSAM_ERR_BAD_ARG;
  if (ncrops <= 0 || ncrops > 16) return MMSAM_ERR_UNSUPPORTED;
  if (B == 0) return MMSAM_OK;
  if (!crop_logits || (!labels_u8 && !preds) || !crop_boxes_host) return MMSAM_ERR_BAD_ARG;
  if ((((uintptr_t)crop_logits | (uintptr_t)preds) & 15)) return MMSAM_ERR_BAD_ARG;
  SlideParams sp;
  sp.ncrops = ncrops;
  for (int j = 0; j < 16; ++j) {
    const bool on = j < ncrops;
    sp.y1[j] = on ? crop_boxes_host[4 * j] : 0;
    sp.x1[j] = on ? crop_boxes_host[4 * j + 1] : 0;
    sp.ch[j] = on ? crop_boxes_host[4 * j + 2] - crop_boxes_host[4 * j] : 0;
    sp.cw[j] = on ? crop_boxes_host[4 * j + 3] - crop_boxes_host[4 * j + 1] : 0;
    if (on && (sp.ch[j] <= 0 || sp.cw[j] <= 0)) return MMSAM_ERR_BAD_ARG;
  }
  slide_merge_kernel<<<pl_grid((long long)B * H * W), 256, 0, (cudaStream_t)stream>>>(
      crop_logits, (uint8_t*)labels_u8, preds, B, hs, ws, ld, ncls, H, W, sp);
  MMSAM_LAUNCH_CHECK();
  return MMSAM_OK;
}
