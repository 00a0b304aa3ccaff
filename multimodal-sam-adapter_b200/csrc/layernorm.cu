// Channel LayerNorm over the last dimension of a [rows, C] bf16 matrix (HBM-bound).
//
// Serves every nn.LayerNorm / LayerNorm2d on the path: ViT Block.norm1/norm2
// (segmentation/mmseg_custom/models/backbones/base/image_encoder.py:398,421), Injector/Extractor
// query_norm/feat_norm/ffn_norm (adapter_modules_multimodal_mix_mod_new_in_twin_convnext_new.py:
// 494-501, 527-532) and ConvNeXt LN2d (mmpretrain_custom/models/utils/norm.py:52-87), which in a
// channels-last layout is the same row-wise operation.
//
// One lane group (8 / 16 / 32 lanes) per row, the row lives in registers (<= 8 x 16 B per lane), two-pass mean/variance in
// fp32 (biased variance, like F.layer_norm). An optional int32 row map scatters the output rows
// (dst = map[src], -1 = drop): that is how norm1 writes straight into SAM's zero-padded 14x14
// window layout (image_encoder.py:504-526) without a separate partition pass. A second scatter mode
// writes each row into its slot of the 2x2-patchified matrix consumed by ConvNeXt's stride-2
// downsample conv (LN2d -> Conv2d(k=2,s=2), twin_convnext.py:313-336), which is then a plain GEMM.
#include "common.cuh"
#include <cstdlib>

namespace mmsam {

// G lanes share a row (G = 32: the classic warp-per-row; 16 / 8 for narrow rows such as ConvNeXt stage 0's C = 96,
// where a full warp per row would leave 20 of 32 lanes idle: 25 % of HBM peak measured, two rows per warp ~2x), and
// every lane group keeps R rows in flight (all loads issued before the first reduction): the per-row chain
// load -> reduce -> reduce -> store is ~2-3 us of latency, and the mid-sized maps only give a warp 1-2 rows.
// XF / YF: the input / output rows are fp32 instead of bf16 (the fp32 residual streams: the ConvNeXt towers' feature
// map and the ViT token stream are kept in fp32 in HBM, their LayerNorms read them directly).
__device__ __forceinline__ void ld8(const void* row, int v, bool f32, float* f) {
  if (f32) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(row) + 2 * v), b = __ldg(reinterpret_cast<const float4*>(row) + 2 * v + 1);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  } else {
    unpack8(__ldg(reinterpret_cast<const uint4*>(row) + v), f);
  }
}
__device__ __forceinline__ void st8(void* row, int v, bool f32, const float* o) {
  if (f32) {
    reinterpret_cast<float4*>(row)[2 * v] = make_float4(o[0], o[1], o[2], o[3]);
    reinterpret_cast<float4*>(row)[2 * v + 1] = make_float4(o[4], o[5], o[6], o[7]);
  } else {
    reinterpret_cast<uint4*>(row)[v] = pack8(o);
  }
}

template <int NV, int G, int R, bool XF, bool YF>
__global__ void __launch_bounds__(256)
layernorm_kernel(const void* __restrict__ x, const float* __restrict__ gamma,
                 const float* __restrict__ beta, void* __restrict__ y, void* __restrict__ y2,
                 const int* __restrict__ row_map, long long rows, int C, long long ldx,
                 long long ldy, float eps, int ps_h, int ps_w) {
  pdl_wait();
  constexpr int XB = XF ? 4 : 2, YB = YF ? 4 : 2;              // bytes per element
  constexpr int RPW = 32 / G;                                   // rows per warp (side by side)
  const int lane = threadIdx.x & (G - 1);                        // lane within the row group
  const int sub = (threadIdx.x & 31) / G;                        // row group within the warp
  const unsigned gmask = G == 32 ? 0xffffffffu : (((1u << G) - 1u) << (sub * G));
  const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  const long long step = nwarps * RPW;
  const int nvec = C >> 3;
  auto group_sum = [&](float v) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(gmask, v, o);
    return v;
  };
  // gamma / beta are staged once per CTA in shared memory, permuted so that the float4 a lane needs for its 8 channels
  // sits next to its neighbours' (conflict-free LDS.128). Reading them from global per row costs 32 L1 wavefronts per
  // 512-byte row slice (fp32 parameters are 2x the bf16 data, at a 32-byte lane stride) against 8 for the data itself:
  // the tag stage, not HBM, capped the kernel at 3.6-4.9 TB/s.
  __shared__ float4 s_g[NV * 2 * G], s_b[NV * 2 * G];
  for (int idx = threadIdx.x; idx < NV * 2 * G; idx += blockDim.x) {
    const int l = idx % G, iq = idx / G, i = iq >> 1, q = iq & 1;
    const int v = l + G * i;
    float4 g4 = make_float4(0.f, 0.f, 0.f, 0.f), b4 = g4;
    if (v < nvec) {
      g4 = __ldg(reinterpret_cast<const float4*>(gamma) + 2 * v + q);
      b4 = __ldg(reinterpret_cast<const float4*>(beta) + 2 * v + q);
    }
    s_g[idx] = g4;
    s_b[idx] = b4;
  }
  __syncthreads();
  for (long long row0 = warp * RPW + sub; row0 < rows; row0 += step * R) {
    long long dst[R], dcol[R];
    bool on[R];
    float f[R][NV][8];
    float s[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const long long row = row0 + r * step;
      on[r] = row < rows;
      dst[r] = row;
      dcol[r] = 0;
      if (on[r]) {
        if (row_map) {
          dst[r] = row_map[row];
          on[r] = dst[r] >= 0;
        } else if (ps_h > 0) {
          // 2x2 patchify scatter for the stride-2 downsample conv: row (b,y,x) -> row (b,y/2,x/2),
          // column block (y&1)*2 + (x&1)
          const long long hw = (long long)ps_h * ps_w;
          const long long b = row / hw;
          const int rr = (int)(row - b * hw);
          const int yy = rr / ps_w, xx = rr - yy * ps_w;
          dst[r] = (b * (ps_h / 2) + yy / 2) * (ps_w / 2) + xx / 2;
          dcol[r] = ((yy & 1) * 2 + (xx & 1)) * C;
        }
      }
      s[r] = 0.f;
      if (on[r]) {
        const char* xr = reinterpret_cast<const char*>(x) + row * ldx * XB;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          const int v = lane + G * i;
          if (v < nvec) {
            ld8(xr, v, XF, f[r][i]);
#pragma unroll
            for (int j = 0; j < 8; ++j) s[r] += f[r][i][j];
          }
        }
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      if (!on[r]) continue;                                      // uniform over the lane group
      const float mean = group_sum(s[r]) / (float)C;
      float s2 = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int v = lane + G * i;
        if (v < nvec) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float d = f[r][i][j] - mean;
            s2 += d * d;
          }
        }
      }
      const float rstd = rsqrtf(group_sum(s2) / (float)C + eps);
      char* yr = reinterpret_cast<char*>(y) + (dst[r] * ldy + dcol[r]) * YB;
      char* yr2 = y2 ? reinterpret_cast<char*>(y2) + (dst[r] * ldy + dcol[r]) * YB : nullptr;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int v = lane + G * i;
        if (v < nvec) {
          const float4 g0 = s_g[(2 * i) * G + lane], g1 = s_g[(2 * i + 1) * G + lane];
          const float4 b0 = s_b[(2 * i) * G + lane], b1 = s_b[(2 * i + 1) * G + lane];
          float o[8];
          o[0] = (f[r][i][0] - mean) * rstd * g0.x + b0.x;
          o[1] = (f[r][i][1] - mean) * rstd * g0.y + b0.y;
          o[2] = (f[r][i][2] - mean) * rstd * g0.z + b0.z;
          o[3] = (f[r][i][3] - mean) * rstd * g0.w + b0.w;
          o[4] = (f[r][i][4] - mean) * rstd * g1.x + b1.x;
          o[5] = (f[r][i][5] - mean) * rstd * g1.y + b1.y;
          o[6] = (f[r][i][6] - mean) * rstd * g1.z + b1.z;
          o[7] = (f[r][i][7] - mean) * rstd * g1.w + b1.w;
          st8(yr, v, YF, o);
          if (yr2) {  // second output: x + LN(x)  (GFE: x + attn(norm1(x)) keeps norm1(x) as a residual)
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] += f[r][i][j];
            st8(yr2, v, YF, o);
          }
        }
      }
    }
  }
}


// Row statistics only: stats[row] = (mean, 1 / sqrt(var + eps)) in fp32, same two-pass arithmetic as layernorm_kernel.
// For LayerNorm -> Linear pairs folded into the GEMM (mmsam_gemm_ln_bf16): LN(x) W^T + b
//   = rstd * (x (g (.) W)^T - mean * s) + c,  s_n = sum_k g_k W_nk,  c_n = sum_k beta_k W_nk + b_n,
// so the normalised copy of x is never written or re-read: this pass costs half of the LayerNorm's HBM traffic.
// Warp per row, C % 8 == 0, C <= 2048.
// NV = 16-byte chunks per lane (C <= 256 * NV): the row lives in 8 * NV registers; every warp keeps TWO rows in flight
// (both rows' loads are issued before the first reduction) and the register budget leaves 5+ CTAs per SM, which is what
// a read-only streaming pass needs to reach HBM speed (a first version with a fixed 64-register row ran at 3.3 TB/s).
template <int NV>
__global__ void __launch_bounds__(256)
rowstats_kernel(const __nv_bfloat16* __restrict__ x, float2* __restrict__ stats, long long rows, int C, long long ldx,
                float eps) {
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  const int nvec = C >> 3;
  const float invC = 1.f / (float)C;
  for (long long row = warp * 2; row < rows; row += nwarps * 2) {
    uint4 raw[2][NV];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const long long rr = row + r < rows ? row + r : row;
      const uint4* xr = reinterpret_cast<const uint4*>(x + rr * ldx);
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int v = lane + 32 * i;
        raw[r][i] = v < nvec ? __ldg(xr + v) : make_uint4(0, 0, 0, 0);
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      float f[NV][8];
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        unpack8(raw[r][i], f[i]);
#pragma unroll
        for (int j = 0; j < 8; ++j) sum += f[i][j];
      }
      const float mean = warp_sum(sum) * invC;      // lanes past the row end contributed zeros
      float var = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        if (lane + 32 * i < nvec) {
#pragma unroll
          for (int j = 0; j < 8; ++j) { const float d = f[i][j] - mean; var = fmaf(d, d, var); }
        }
      }
      var = warp_sum(var) * invC;
      if (lane == 0 && row + r < rows) stats[row + r] = make_float2(mean, rsqrtf(var + eps));
    }
  }
}

}  // namespace mmsam

MMSAM_API int mmsam_layernorm(const void* x, int x_dtype, const float* gamma, const float* beta, void* y, int y_dtype, void* y2,
                              const int* row_map_dev, long long rows, int C, long long ldx,
                              long long ldy, float eps, int ps_h, int ps_w, void* stream) {
  using namespace mmsam;
  if (rows < 0 || C <= 0 || (C & 7) || C > 2048 || (ldx & 7) || (ldy & 7)) return MMSAM_ERR_BAD_ARG;
  if ((x_dtype != MMSAM_BF16 && x_dtype != MMSAM_F32) || (y_dtype != MMSAM_BF16 && y_dtype != MMSAM_F32)) return MMSAM_ERR_BAD_DTYPE;
  if (x_dtype == MMSAM_BF16 && y_dtype == MMSAM_F32) return MMSAM_ERR_UNSUPPORTED;     // bf16 -> fp32 is not instantiated
  const int mode = x_dtype == MMSAM_BF16 ? 0 : (y_dtype == MMSAM_BF16 ? 1 : 2);
  if (ps_h < 0 || ps_w < 0 || (ps_h > 0 && ((ps_h | ps_w) & 1))) return MMSAM_ERR_BAD_ARG;
  if (rows == 0) return MMSAM_OK;
  if (!x || !gamma || !beta || !y) return MMSAM_ERR_BAD_ARG;
  if ((((uintptr_t)x | (uintptr_t)y | (uintptr_t)gamma | (uintptr_t)beta) & 15)) return MMSAM_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const int wpb = 8;
  const int nvec = C / 8;
  // C = 96 (12 chunks, ConvNeXt stage 0 at 1/4 resolution): 16 lanes per row leave 4 idle; 4 lanes x 3 chunks use them all
  // likewise C = 192 (24 chunks) = 8 lanes x 3 and C = 384 (48 chunks) = 16 lanes x 3 instead of 24 / 48 of 32 / 64 slots
  static const bool narrow = !getenv("MMSAM_LN_NO_G4");
  const int G = (narrow && nvec == 12) ? 4 : (narrow && nvec == 24) ? 8 : (narrow && nvec == 48) ? 16
                : (nvec > 16 ? 32 : (nvec > 8 ? 16 : 8));
  const int rpw = 32 / G;
  const int nv = (nvec + 31) / 32;
  const int R = 1;   // rows in flight per lane group: 2 was measured SLOWER everywhere (registers -> occupancy)
  long long blocks = (rows + (long long)wpb * rpw * R - 1) / ((long long)wpb * rpw * R);
  static const int cap_waves = getenv("MMSAM_LN_CAP") ? atoi(getenv("MMSAM_LN_CAP")) : 16;
  const long long cap = (long long)kNumSMs * cap_waves;
  if (blocks > cap) blocks = cap;
#define LN_LAUNCH(NV, GG, RR, XFF, YFF) \
  mmsam_host::launch_pdl(layernorm_kernel<NV, GG, RR, XFF, YFF>, dim3((unsigned)blocks), dim3(wpb * 32), 0, st, x, gamma, beta, y, y2, row_map_dev, rows, C, ldx, ldy, eps, ps_h, ps_w)
#define LN_CASE(NV, GG, RR)                                    \
  do {                                                         \
    if (mode == 0) LN_LAUNCH(NV, GG, RR, false, false);        \
    else if (mode == 1) LN_LAUNCH(NV, GG, RR, true, false);    \
    else LN_LAUNCH(NV, GG, RR, true, true);                    \
  } while (0)
  if (G == 4) LN_CASE(3, 4, 1);
  else if (G == 8 && nvec == 24) LN_CASE(3, 8, 1);
  else if (G == 8) LN_CASE(1, 8, 1);
  else if (G == 16 && nvec == 48) LN_CASE(3, 16, 1);
  else if (G == 16) LN_CASE(1, 16, 1);
  else switch (nv) {
    case 1: LN_CASE(1, 32, 1); break;
    case 2: LN_CASE(2, 32, 1); break;
    case 3: LN_CASE(3, 32, 1); break;
    case 4: LN_CASE(4, 32, 1); break;
    case 5: case 6: LN_CASE(6, 32, 1); break;
    default: LN_CASE(8, 32, 1); break;
  }
#undef LN_CASE
#undef LN_LAUNCH
  MMSAM_LAUNCH_CHECK();
  return MMSAM_OK;
}

MMSAM_API int mmsam_rowstats_bf16(const void* x, float* stats, long long rows, int C, long long ldx, float eps, void* stream) {
  using namespace mmsam;
  if (rows < 0 || C <= 0 || (C & 7) || C > 2048 || ldx < C || (ldx & 7)) return MMSAM_ERR_BAD_ARG;
  if (rows == 0) return MMSAM_OK;
  if (!x || !stats || (((uintptr_t)x) & 15) || (((uintptr_t)stats) & 7)) return MMSAM_ERR_BAD_ARG;
  long long blocks = (rows + 15) / 16;
  if (blocks > (long long)kNumSMs * 16) blocks = (long long)kNumSMs * 16;
  const int nv = (C / 8 + 31) / 32;
  cudaStream_t st = (cudaStream_t)stream;
#define ROWSTATS(NVV) mmsam_host::launch_pdl(rowstats_kernel<NVV>, dim3((unsigned)blocks), dim3(256), 0, st, (const __nv_bfloat16*)x, (float2*)stats, rows, C, ldx, eps)
  if (nv <= 1) ROWSTATS(1);
  else if (nv <= 2) ROWSTATS(2);
  else if (nv <= 4) ROWSTATS(4);
  else ROWSTATS(8);
#undef ROWSTATS
  MMSAM_LAUNCH_CHECK();
  return MMSAM_OK;
}
