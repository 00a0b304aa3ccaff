// Pixel-sized kernels of the RoadFormer2Neck fusion (segmentation/mmseg_custom/models/backbones/
// adapter_modules_multimodal_mix_mod_new_in_twin_convnext_new.py:75-394) on channels-last bf16 maps.
// All of them stream the maps once (HBM-bound); the dense contractions of the neck run on gemm.cu /
// conv3x3.cu. Numerically sensitive reductions (HW-long Gram sums, LayerNorm-over-HW statistics) are
// accumulated in fp32 per chunk and combined in fp64 / fp32 atomics.
#include "common.cuh"
#include <cstdlib>

namespace mmsam {

// ------------------------------------------------------------------------------------------------
// Gram matrix over pixels:  S[b, i, j] += sum_pix X[b, pix, qoff + i] * X[b, pix, koff + j]
// (AttentionBase q @ k^T over HW, :98-103, and GFFM's cross-modal energy, :250-254), plus the squared
// norms of the q / k rows (F.normalize, :100-101). 64x64 output tile per CTA, 4x4 per thread, pixels
// staged 32 at a time through shared memory as fp32. blk > 0: only the block-diagonal (per-head)
// elements are produced.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gram_kernel(const __nv_bfloat16* __restrict__ X, long long ld, int qoff, int koff, int n, int HW, int chunk,
            int blk, float* __restrict__ S, float* __restrict__ nq, float* __restrict__ nk) {
  __shared__ __align__(16) float sQ[32][64];
  __shared__ __align__(16) float sK[32][64];
  const int nt = (n + 63) / 64;
  const int ti = blockIdx.x / nt, tj = blockIdx.x % nt;
  if (blk > 0) {  // skip tiles that do not touch the block diagonal
    const int i0 = ti * 64, i1 = min(i0 + 63, n - 1), j0 = tj * 64, j1 = min(j0 + 63, n - 1);
    if (i1 / blk < j0 / blk || j1 / blk < i0 / blk) return;
  }
  const int b = blockIdx.z;
  const int p0 = blockIdx.y * chunk, p1 = min(p0 + chunk, HW);
  const __nv_bfloat16* Xb = X + (long long)b * HW * ld;
  const int t = threadIdx.x;
  const int ty = t >> 4, tx = t & 15;
  const int lpx = t >> 3, lcv = t & 7;  // loader: pixel within step, 8-channel vector
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float nacc = 0.f;
  const bool do_norm = nq != nullptr && ti == tj && t < 128;
  for (int ps = p0; ps < p1; ps += 32) {
    {
      const int pix = ps + lpx;
      float fq[8], fk[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) fq[j] = fk[j] = 0.f;
      if (pix < p1) {
        const int cq = ti * 64 + lcv * 8, ck = tj * 64 + lcv * 8;
        if (cq < n) unpack8(__ldg(reinterpret_cast<const uint4*>(Xb + (long long)pix * ld + qoff + cq)), fq);
        if (ck < n) unpack8(__ldg(reinterpret_cast<const uint4*>(Xb + (long long)pix * ld + koff + ck)), fk);
      }
      *reinterpret_cast<float4*>(&sQ[lpx][lcv * 8]) = make_float4(fq[0], fq[1], fq[2], fq[3]);
      *reinterpret_cast<float4*>(&sQ[lpx][lcv * 8 + 4]) = make_float4(fq[4], fq[5], fq[6], fq[7]);
      *reinterpret_cast<float4*>(&sK[lpx][lcv * 8]) = make_float4(fk[0], fk[1], fk[2], fk[3]);
      *reinterpret_cast<float4*>(&sK[lpx][lcv * 8 + 4]) = make_float4(fk[4], fk[5], fk[6], fk[7]);
    }
    __syncthreads();
#pragma unroll 8
    for (int px = 0; px < 32; ++px) {
      const float4 q = *reinterpret_cast<const float4*>(&sQ[px][ty * 4]);
      const float4 k = *reinterpret_cast<const float4*>(&sK[px][tx * 4]);
      const float qa[4] = {q.x, q.y, q.z, q.w}, ka[4] = {k.x, k.y, k.z, k.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(qa[i], ka[j], acc[i][j]);
    }
    if (do_norm) {
      const float* src = t < 64 ? &sQ[0][t] : &sK[0][t - 64];
#pragma unroll 8
      for (int px = 0; px < 32; ++px) nacc = fmaf(src[px * 64], src[px * 64], nacc);
    }
    __syncthreads();
  }
  float* Sb = S + (long long)b * n * n;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gi = ti * 64 + ty * 4 + i;
    if (gi >= n) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gj = tj * 64 + tx * 4 + j;
      if (gj >= n || (blk > 0 && gi / blk != gj / blk)) continue;
      atomicAdd(Sb + (long long)gi * n + gj, acc[i][j]);
    }
  }
  if (do_norm) {
    const int c = ti * 64 + (t & 63);
    if (c < n) atomicAdd((t < 64 ? nq : nk) + (long long)b * n + c, nacc);
  }
}

// ------------------------------------------------------------------------------------------------
// Per-chunk column statistics of o [B, HW, C] for GFFM's LayerNorm over the spatial axis (:262-264):
// part[chunk, b, c, {sum o, sum o^2, sum o*w[pix]}] in fp32 (chunks of 64 or 512 px; combined in fp64 later).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
colstats_kernel(const __nv_bfloat16* __restrict__ o, const float* __restrict__ wpix, float* __restrict__ part,
                int B, int HW, int C, int chunk) {
  const int CV = C >> 3;
  const int b = blockIdx.y, ch = blockIdx.x;
  const int p0 = ch * chunk, p1 = min(p0 + chunk, HW);
  __shared__ float sred[3][2048];
  const int PL = blockDim.x / CV;  // pixel lanes (blockDim.x is a multiple of CV)
  const int v = threadIdx.x % CV, pl = threadIdx.x / CV;
  float s[8], q[8], w[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = q[j] = w[j] = 0.f;
  if (pl < PL) {
    for (int pix = p0 + pl; pix < p1; pix += PL) {
      float f[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(o + ((long long)b * HW + pix) * C + v * 8)), f);
      const float wp = __ldg(wpix + pix);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s[j] += f[j];
        q[j] = fmaf(f[j], f[j], q[j]);
        w[j] = fmaf(f[j], wp, w[j]);
      }
    }
  }
  // reduce the PL pixel lanes through shared memory (C <= 2048)
  for (int j = threadIdx.x; j < C; j += blockDim.x) { sred[0][j] = 0.f; sred[1][j] = 0.f; sred[2][j] = 0.f; }
  __syncthreads();
  if (pl < PL) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      atomicAdd(&sred[0][v * 8 + j], s[j]);
      atomicAdd(&sred[1][v * 8 + j], q[j]);
      atomicAdd(&sred[2][v * 8 + j], w[j]);
    }
  }
  __syncthreads();
  float* dst = part + (((long long)ch * B + b) * C) * 3;
  for (int j = threadIdx.x; j < C; j += blockDim.x) {
    dst[j * 3 + 0] = sred[0][j];
    dst[j * 3 + 1] = sred[1][j];
    dst[j * 3 + 2] = sred[2][j];
  }
}

// u[pix, c] = gelu(a[pix, c]) * a[pix, C + c]   (Mlp gate, :129-130)
__global__ void __launch_bounds__(256)
gate_kernel(const __nv_bfloat16* __restrict__ a, __nv_bfloat16* __restrict__ u, long long rows, int C) {
  const int CV = C >> 3;
  const long long total = rows * CV;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long r = idx / CV;
    const int v = (int)(idx - r * CV);
    float x1[8], x2[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(a + r * 2 * C + v * 8)), x1);
    unpack8(__ldg(reinterpret_cast<const uint4*>(a + r * 2 * C + C + v * 8)), x2);
#pragma unroll
    for (int j = 0; j < 8; ++j) x1[j] = gelu_erf(x1[j]) * x2[j];
    *reinterpret_cast<uint4*>(u + r * C + v * 8) = pack8(x1);
  }
}

// ------------------------------------------------------------------------------------------------
// f = s1 * LN_HW(o) * (1 + ffrm_gate) + s2 * lo      (GFFM norm :262-264, FFRM :158-162, Scale2 :279-280)
//   LN_HW(o)[pix, c] = (o - mu[b,c]) * rstd[b,c] * w[pix] + bias[pix]
// and the coordinate-attention pools (:190-196): ph[b, y, c] = sum_x f, pw_part[strip, b, x, c] = sum_{y in strip} f.
// CTA = (strip of RS rows) x (64 channels) x image; thread = (8-channel vector, one of 32 columns).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
combine_pool_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ lo,
                    const float* __restrict__ mu, const float* __restrict__ rstd, const float* __restrict__ gate,
                    const float* __restrict__ wpix, const float* __restrict__ bpix, float s1, float s2,
                    __nv_bfloat16* __restrict__ f, float* __restrict__ ph, float* __restrict__ pw_part, int B, int H,
                    int W, int C, int RS) {
  extern __shared__ float s_ph[];  // [RS][64]
  const int strip = blockIdx.x, cb = blockIdx.y, b = blockIdx.z;
  const int cv = threadIdx.x & 7, xl = threadIdx.x >> 3;
  const int c = cb * 64 + cv * 8;
  const bool cok = c < C;
  const int y0 = strip * RS, y1 = min(y0 + RS, H);
  for (int i = threadIdx.x; i < RS * 64; i += blockDim.x) s_ph[i] = 0.f;
  __syncthreads();
  float m[8], r[8], g[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    m[j] = cok ? mu[(long long)b * C + c + j] : 0.f;
    r[j] = cok ? rstd[(long long)b * C + c + j] : 0.f;
    g[j] = cok ? gate[(long long)b * C + c + j] * s1 : 0.f;
  }
  for (int xs = 0; xs < W; xs += 32) {
    const int x = xs + xl;
    float col[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) col[j] = 0.f;
    for (int y = y0; y < y1; ++y) {
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = 0.f;
      if (x < W && cok) {
        const long long pix = (long long)y * W + x;
        const long long off = ((long long)b * H * W + pix) * C + c;
        float a[8], l[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(o + off)), a);
        unpack8(__ldg(reinterpret_cast<const uint4*>(lo + off)), l);
        const float wp = __ldg(wpix + pix), bp = __ldg(bpix + pix);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          v[j] = fmaf(fmaf((a[j] - m[j]) * r[j], wp, bp), g[j], s2 * l[j]);
          col[j] += v[j];
        }
        *reinterpret_cast<uint4*>(f + off) = pack8(v);
      }
      // row sums: reduce the 4 columns held by this warp, then shared-memory atomics
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float t = v[j];
        t += __shfl_xor_sync(0xffffffffu, t, 8);
        t += __shfl_xor_sync(0xffffffffu, t, 16);
        if ((threadIdx.x & 31) < 8) atomicAdd(&s_ph[(y - y0) * 64 + cv * 8 + j], t);
      }
    }
    if (x < W && cok) {
      float* dst = pw_part + (((long long)strip * B + b) * W + x) * C + c;
      *reinterpret_cast<float4*>(dst) = make_float4(col[0], col[1], col[2], col[3]);
      *reinterpret_cast<float4*>(dst + 4) = make_float4(col[4], col[5], col[6], col[7]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < (y1 - y0) * 64; i += blockDim.x) {
    const int yy = i >> 6, cc = cb * 64 + (i & 63);
    if (cc < C) ph[((long long)b * H + y0 + yy) * C + cc] = s_ph[i];
  }
}

// out[pix, c] = f[pix, c] * (1 + aw[b, x, c] * ah[b, y, c])     (CoordinateAttention + CA residual, :187-221)
__global__ void __launch_bounds__(256)
ca_apply_kernel(const __nv_bfloat16* __restrict__ f, const float* __restrict__ ah, const float* __restrict__ aw,
                __nv_bfloat16* __restrict__ out, int B, int H, int W, int C) {
  const int CV = C >> 3;
  const long long total = (long long)B * H * W * CV;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(idx % CV);
    long long t = idx / CV;
    const int x = (int)(t % W); t /= W;
    const int y = (int)(t % H);
    const int b = (int)(t / H);
    float a[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(f + idx * 8)), a);
    const float* hp = ah + ((long long)b * H + y) * C + v * 8;
    const float* wp = aw + ((long long)b * W + x) * C + v * 8;
    const float4 h0 = __ldg(reinterpret_cast<const float4*>(hp)), h1 = __ldg(reinterpret_cast<const float4*>(hp + 4));
    const float4 w0 = __ldg(reinterpret_cast<const float4*>(wp)), w1 = __ldg(reinterpret_cast<const float4*>(wp + 4));
    a[0] *= fmaf(h0.x, w0.x, 1.f); a[1] *= fmaf(h0.y, w0.y, 1.f); a[2] *= fmaf(h0.z, w0.z, 1.f); a[3] *= fmaf(h0.w, w0.w, 1.f);
    a[4] *= fmaf(h1.x, w1.x, 1.f); a[5] *= fmaf(h1.y, w1.y, 1.f); a[6] *= fmaf(h1.z, w1.z, 1.f); a[7] *= fmaf(h1.w, w1.w, 1.f);
    *reinterpret_cast<uint4*>(out + idx * 8) = pack8(a);
  }
}

static inline unsigned nk_grid(long long total, int per_block = 256, int waves = 16) {
  long long b = (total + per_block - 1) / per_block;
  const long long cap = (long long)kNumSMs * waves;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (unsigned)b;
}

}  // namespace mmsam

int mmsam_gram_tc(const void* X, long long ld, int qoff, int koff, int n, int B, int HW, int blk, float* S, float* nq,
                  float* nk, cudaStream_t st);   // gram_tc.cu

MMSAM_API int mmsam_gram_bf16(const void* X, long long ld, int qoff, int koff, int n, int B, int HW, int blk,
                              float* S, float* nq, float* nk, void* stream) {
  using namespace mmsam;
  if (B < 0 || n <= 0 || HW <= 0 || (ld & 7) || (qoff & 7) || (koff & 7) || (n & 7) || blk < 0) return MMSAM_ERR_BAD_ARG;
  if ((nq == nullptr) != (nk == nullptr)) return MMSAM_ERR_BAD_ARG;
  if (B == 0) return MMSAM_OK;
  if (!X || !S || (((uintptr_t)X) & 15)) return MMSAM_ERR_BAD_ARG;
  if (!getenv("MMSAM_GRAM_SIMT")) {   // tensor-core path; shapes it does not take fall through to the SIMT kernel
    const int rc = mmsam_gram_tc(X, ld, qoff, koff, n, B, HW, blk, S, nq, nk, (cudaStream_t)stream);
    if (rc != MMSAM_ERR_UNSUPPORTED) return rc;
  }
  const int nt = (n + 63) / 64;
  int nchunks = HW / 2048;
  if (nchunks < 1) nchunks = 1;
  if (nchunks > 16) nchunks = 16;
  int chunk = (HW + nchunks - 1) / nchunks;
  chunk = (chunk + 31) / 32 * 32;
  nchunks = (HW + chunk - 1) / chunk;
  dim3 grid(nt * nt, nchunks, B);
  gram_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)X, ld, qoff, koff, n, HW, chunk, blk, S, nq, nk);
  MMSAM_LAUNCH_CHECK();
  return MMSAM_OK;
}

// part must hold mmsam_colstats_chunks(HW) * B * C * 3 floats
// pixels per chunk: small maps (the 32^2 / 64^2 levels carry 768 - 1536 channels, i.e. 1 - 2 pixel lanes per CTA) get short
// chunks so that the launch still fills the GPU (16 CTAs walking 512 pixels each took 205 us for 25 MB)
static inline int colstats_chunk_px(int HW) { return HW >= 16384 ? 512 : 64; }
MMSAM_API int mmsam_colstats_chunks(int HW) { const int c = colstats_chunk_px(HW); return (HW + c - 1) / c; }

MMSAM_API int mmsam_colstats_bf16(const void* o, const float* wpix, float* part, int B, int HW, int C, void* stream) {
  using namespace mmsam;
  if (B < 0 || HW <= 0 || C <= 0 || (C & 7) || C > 2048) return MMSAM_ERR_BAD_ARG;
  if (B == 0) return MMSAM_OK;
  if (!o || !wpix || !part || (((uintptr_t)o) & 15)) return MMSAM_ERR_BAD_ARG;
  const int CV = C / 8;
  if (CV > 256) return MMSAM_ERR_UNSUPPORTED;
  const int PL = 256 / CV;
  dim3 grid(mmsam_colstats_chunks(HW), B);
  colstats_kernel<<<grid, CV * PL, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)o, wpix, part, B, HW, C, colstats_chunk_px(HW));
  MMSAM_LAUNCH_CHECK();
  return MMSAM_OK;
}

MMSAM_API int mmsam_gate_bf16(const void* a, void* u, long long rows, int C, void* stream) {
  using namespace mmsam;
  if (rows < 0 || C <= 0 || (C & 7)) return MMSAM_ERR_BAD_ARG;
  if (rows == 0) return MMSAM_OK;
  if (!a || !u || (((uintptr_t)a | (uintptr_t)u) & 15)) return MMSAM_ERR_BAD_ARG;
  gate_kernel<<<nk_grid(rows * (C / 8)), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)a, (__nv_bfloat16*)u, rows, C);
  MMSAM_LAUNCH_CHECK();
  return MMSAM_OK;
}

MMSAM_API int mmsam_combine_pool_rows(int H) { return H >= 64 ? 16 : (H >= 16 ? 8 : (H >= 4 ? 4 : 1)); }

// ph [B,H,C] fp32 (row sums), pw_part [ceil(H/RS), B, W, C] fp32 (per-strip column sums), RS = mmsam_combine_pool_rows(H)
MMSAM_API int mmsam_combine_pool_bf16(const void* o, const void* lo, const float* mu, const float* rstd,
                                      const float* gate, const float* wpix, const float* bpix, float s1, float s2,
                                      void* f, float* ph, float* pw_part, int B, int H, int W, int C, void* stream) {
  using namespace mmsam;
  if (B < 0 || H <= 0 || W <= 0 || C <= 0 || (C & 7)) return MMSAM_ERR_BAD_ARG;
  if (B == 0) return MMSAM_OK;
  if (!o || !lo || !mu || !rstd || !gate || !wpix || !bpix || !f || !ph || !pw_part) return MMSAM_ERR_BAD_ARG;
  if ((((uintptr_t)o | (uintptr_t)lo | (uintptr_t)f | (uintptr_t)pw_part) & 15)) return MMSAM_ERR_BAD_ARG;
  const int RS = mmsam_combine_pool_rows(H);
  dim3 grid((H + RS - 1) / RS, (C + 63) / 64, B);
  combine_pool_kernel<<<grid, 256, RS * 64 * sizeof(float), (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)o, (const __nv_bfloat16*)lo, mu, rstd, gate, wpix, bpix, s1, s2, (__nv_bfloat16*)f, ph,
      pw_part, B, H, W, C, RS);
  MMSAM_LAUNCH_CHECK();
  return MMSAM_OK;
}

MMSAM_API int mmsam_ca_apply_bf16(const void* f, const float* ah, const float* aw, void* out, int B, int H, int W,
                                  int C, void* stream) {
  using namespace mmsam;
  if (B < 0 || H <= 0 || W <= 0 || C <= 0 || (C & 7)) return MMSAM_ERR_BAD_ARG;
  if (B == 0) return MMSAM_OK;
  if (!f || !ah || !aw || !out || (((uintptr_t)f | (uintptr_t)out | (uintptr_t)ah | (uintptr_t)aw) & 15)) return MMSAM_ERR_BAD_ARG;
  ca_apply_kernel<<<nk_grid((long long)B * H * W * (C / 8)), 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)f, ah, aw, (__nv_bfloat16*)out, B, H, W, C);
  MMSAM_LAUNCH_CHECK();
  return MMSAM_OK;
}
