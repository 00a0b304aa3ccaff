// Pixel-sized kernels of the RoadFormer2Neck fusion (segmentation/mmseg_custom/models/backbones/
// adapter_modules_multimodal_mix_mod_new_in_twin_convnext_new.py:75-394) on channels-last bf16 maps.
// All of them stream the maps once (HBM-bound); the dense contractions of the neck run on gemm.cu /
// conv3x3.cu. Numerically sensitive reductions (HW-long Gram sums, LayerNorm-over-HW statistics) are accumulated in
// fp32 per pixel chunk and the chunks are combined in a FIXED order by the consumer kernel: there is no floating-point
// atomic on the path, results are bit-identical from run to run, eager or graph, whatever the batch an image sits in.
// The O(B*C^2) steps between the pixel passes (softmax of the channel-attention matrices with proj folded in, GFFM's
// two softmaxes, LayerNorm-over-HW statistics -> FFRM gate, coordinate-attention vectors) are small kernels of this
// file as well: the neck launches no library kernel.
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
  const long long slot = (long long)blockIdx.y * gridDim.z + b;   // (chunk, image): this CTA's partial sums
  float* Sb = S + slot * n * n;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gi = ti * 64 + ty * 4 + i;
    if (gi >= n) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gj = tj * 64 + tx * 4 + j;
      if (gj >= n || (blk > 0 && gi / blk != gj / blk)) continue;
      Sb[(long long)gi * n + gj] = acc[i][j];
    }
  }
  if (do_norm) {
    const int c = ti * 64 + (t & 63);
    if (c < n) (t < 64 ? nq : nk)[slot * n + c] = nacc;
  }
}

// ------------------------------------------------------------------------------------------------
// AttentionBase after the Gram pass (:100-107), per (head, image): add the chunk partials in order,
//   att = softmax_j( S[a, j] / (max(|q_a|, 1e-12) max(|k_j|, 1e-12)) * temperature[h] )
// and fold `proj` (and scale2) into one [ci, ci] matrix per image so that attn @ v followed by proj is ONE GEMM over
// the pixels:   Weff[b, i, h*ch + j] = scale2 * sum_a Wproj[i, h*ch + a] * att[b, h, a, j].
// grid (heads, B, row splits of Weff); every CTA recomputes its head's att (ch x ch, <= 96 x 96) in shared memory.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gfe_weff_kernel(const float* __restrict__ S_part, const float* __restrict__ nq_part, const float* __restrict__ nk_part,
                int nchunks, int B, int ci, int heads, const float* __restrict__ temp, const float* __restrict__ wproj,
                const float* __restrict__ scale2, __nv_bfloat16* __restrict__ weff) {
  extern __shared__ float s_gf[];
  const int ch = ci / heads;
  const bool vec = (ch & 3) == 0;      // 4 x 4 register tiles with 16-byte loads in the proj fold
  const int cs = vec ? ch + 4 : ch + 1;
  float* att = s_gf;                 // [ch][cs]
  float* rq = att + ch * cs;         // [ch] 1 / max(|q|, eps)
  float* rk = rq + ch;
  const int h = blockIdx.x, b = blockIdx.y;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31, nwarps = blockDim.x >> 5;
  const long long nn = (long long)ci * ci;
  // four elements per trip: their chunk-0 loads are requested together (one load in flight per thread made this phase a
  // chain of ~36 L2 latencies per CTA at ch = 96)
  for (int e0 = t; e0 < ch * ch; e0 += 4 * blockDim.x) {
    float acc[4];
    const float* src[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int e = min(e0 + u * (int)blockDim.x, ch * ch - 1);
      const int a = e / ch, j = e - a * ch;
      src[u] = S_part + (long long)b * nn + (long long)(h * ch + a) * ci + h * ch + j;
      acc[u] = src[u][0];
    }
    for (int c = 1; c < nchunks; ++c) {
#pragma unroll
      for (int u = 0; u < 4; ++u) acc[u] += src[u][(long long)c * B * nn];
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int e = e0 + u * (int)blockDim.x;
      if (e < ch * ch) { const int a = e / ch, j = e - a * ch; att[a * cs + j] = acc[u]; }
    }
  }
  for (int e = t; e < 2 * ch; e += blockDim.x) {
    const float* src = (e < ch ? nq_part : nk_part) + (long long)b * ci + h * ch + (e < ch ? e : e - ch);
    float acc = 0.f;
    for (int c = 0; c < nchunks; ++c) acc += src[(long long)c * B * ci];
    (e < ch ? rq : rk)[e < ch ? e : e - ch] = 1.f / fmaxf(sqrtf(acc), 1e-12f);
  }
  __syncthreads();
  const float tp = temp[h];
  for (int a = warp; a < ch; a += nwarps) {          // warp = one softmax row
    float* row = att + a * cs;
    float mx = -INFINITY;
    for (int j = lane; j < ch; j += 32) {
      const float v = row[j] * rq[a] * rk[j] * tp;
      row[j] = v;
      mx = fmaxf(mx, v);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < ch; j += 32) {
      const float e = __expf(row[j] - mx);
      row[j] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
    for (int j = lane; j < ch; j += 32) row[j] *= inv;
  }
  __syncthreads();
  // proj fold: Weff[i, h ch + j] = scale2 * sum_a wproj[i, h ch + a] * att[a, j]  for this CTA's rows [i0, i1).
  // ci x ch x ch FMAs per (head, image): with one load pair per FMA the kernel sat on the LSU (238 us at ci = 768);
  // a thread now owns a 4 x 4 tile and reads wproj / att 16 bytes at a time (8 loads per 64 FMAs).
  const float s2 = scale2[0];
  const int rows_per = (ci + gridDim.z - 1) / gridDim.z;
  const int i0 = blockIdx.z * rows_per, i1 = min(i0 + rows_per, ci);
  if (vec) {
    const int jq_n = ch >> 2, iq_n = (i1 - i0 + 3) >> 2;
    for (int tile = t; tile < iq_n * jq_n; tile += blockDim.x) {
      const int iq = tile / jq_n, jq = tile - iq * jq_n;
      const int ib = i0 + iq * 4;
      const float* wp[4];
#pragma unroll
      for (int ii = 0; ii < 4; ++ii) wp[ii] = wproj + (long long)min(ib + ii, i1 - 1) * ci + h * ch;
      float acc[4][4];
#pragma unroll
      for (int ii = 0; ii < 4; ++ii)
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) acc[ii][jj] = 0.f;
#pragma unroll 2
      for (int a4 = 0; a4 < ch; a4 += 4) {
        float4 w[4], at[4];
#pragma unroll
        for (int ii = 0; ii < 4; ++ii) w[ii] = __ldg(reinterpret_cast<const float4*>(wp[ii] + a4));
#pragma unroll
        for (int aa = 0; aa < 4; ++aa) at[aa] = *reinterpret_cast<const float4*>(att + (a4 + aa) * cs + jq * 4);
#pragma unroll
        for (int ii = 0; ii < 4; ++ii) {
          const float wv[4] = {w[ii].x, w[ii].y, w[ii].z, w[ii].w};
#pragma unroll
          for (int aa = 0; aa < 4; ++aa) {      // a ascending: the summation order of the scalar path
            acc[ii][0] = fmaf(wv[aa], at[aa].x, acc[ii][0]);
            acc[ii][1] = fmaf(wv[aa], at[aa].y, acc[ii][1]);
            acc[ii][2] = fmaf(wv[aa], at[aa].z, acc[ii][2]);
            acc[ii][3] = fmaf(wv[aa], at[aa].w, acc[ii][3]);
          }
        }
      }
#pragma unroll
      for (int ii = 0; ii < 4; ++ii) {
        if (ib + ii < i1) {
          __nv_bfloat162 lo = __floats2bfloat162_rn(acc[ii][0] * s2, acc[ii][1] * s2);
          __nv_bfloat162 hi = __floats2bfloat162_rn(acc[ii][2] * s2, acc[ii][3] * s2);
          uint2 v;
          v.x = *reinterpret_cast<uint32_t*>(&lo);
          v.y = *reinterpret_cast<uint32_t*>(&hi);
          *reinterpret_cast<uint2*>(weff + (long long)b * nn + (long long)(ib + ii) * ci + h * ch + jq * 4) = v;
        }
      }
    }
  } else {
    for (int e = t; e < (i1 - i0) * ch; e += blockDim.x) {
      const int i = i0 + e / ch, j = e % ch;
      const float* wp = wproj + (long long)i * ci + h * ch;
      float acc = 0.f;
      for (int a = 0; a < ch; ++a) acc = fmaf(__ldg(wp + a), att[a * cs + j], acc);
      weff[(long long)b * nn + (long long)i * ci + h * ch + j] = __float2bfloat16_rn(acc * s2);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// GFFM attention maps (:250-259) from the chunk partials of the cross-modal energy E = fx^T fy  [B, ci, ci]:
//   ax[b, i, :] = softmax_j E[b, i, j]      ay[b, j, :] = softmax_i E[b, i, j]     (bf16, the GEMM weights)
// grid (ci, B, 2); block = one row (or column) of E.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gffm_softmax_kernel(const float* __restrict__ E_part, int nchunks, int B, int ci, __nv_bfloat16* __restrict__ ax,
                    __nv_bfloat16* __restrict__ ay) {
  __shared__ float s_red[8];
  __shared__ float s_bc;
  const int r = blockIdx.x, b = blockIdx.y, tr = blockIdx.z;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const long long nn = (long long)ci * ci;
  constexpr int PER = 4;                                   // ci <= 1024
  float v[PER];
  float mx = -INFINITY;
#pragma unroll
  for (int k = 0; k < PER; ++k) {
    const int j = t + k * 256;
    v[k] = -INFINITY;
    if (j < ci) {
      const float* src = E_part + (long long)b * nn + (tr == 0 ? (long long)r * ci + j : (long long)j * ci + r);
      float acc = 0.f;
      for (int c = 0; c < nchunks; ++c) acc += src[(long long)c * B * nn];
      v[k] = acc;
      mx = fmaxf(mx, acc);
    }
  }
  mx = warp_max(mx);
  if (lane == 0) s_red[warp] = mx;
  __syncthreads();
  if (t == 0) {
    float m = s_red[0];
    for (int w = 1; w < 8; ++w) m = fmaxf(m, s_red[w]);
    s_bc = m;
  }
  __syncthreads();
  mx = s_bc;
  float sum = 0.f;
#pragma unroll
  for (int k = 0; k < PER; ++k) {
    v[k] = (t + k * 256 < ci) ? __expf(v[k] - mx) : 0.f;
    sum += v[k];
  }
  sum = warp_sum(sum);
  __syncthreads();
  if (lane == 0) s_red[warp] = sum;
  __syncthreads();
  if (t == 0) {
    float a = 0.f;
    for (int w = 0; w < 8; ++w) a += s_red[w];           // fixed order
    s_bc = 1.f / a;
  }
  __syncthreads();
  const float inv = s_bc;
  __nv_bfloat16* dst = (tr == 0 ? ax : ay) + (long long)b * nn + (long long)r * ci;
#pragma unroll
  for (int k = 0; k < PER; ++k) {
    const int j = t + k * 256;
    if (j < ci) dst[j] = __float2bfloat16_rn(v[k] * inv);
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
  __shared__ float sred[256 * 24];   // [pixel lane][C][3], PL * C <= 2048
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
  // reduce the PL pixel lanes through shared memory in a fixed order (no atomics: bit-reproducible)
  if (pl < PL) {
    float* mine = sred + ((long long)pl * C + v * 8) * 3;
#pragma unroll
    for (int j = 0; j < 8; ++j) { mine[j * 3 + 0] = s[j]; mine[j * 3 + 1] = q[j]; mine[j * 3 + 2] = w[j]; }
  }
  __syncthreads();
  float* dst = part + (((long long)ch * B + b) * C) * 3;
  for (int j = threadIdx.x; j < C * 3; j += blockDim.x) {
    float a = 0.f;
    for (int l = 0; l < PL; ++l) a += sred[(long long)l * C * 3 + j];
    dst[j] = a;
  }
}

// ------------------------------------------------------------------------------------------------
// GFFM's LayerNorm over the spatial axis (:262-264) + FFRM gate (:148-162) from the colstats partials, per image:
//   mu, rstd per channel (fp64 combination of the chunk partials, in order), GAP of LN_HW(o) analytically
//   (gap = rstd * (sum o*w - mu * sum w) / HW + mean b), a = Wffrm gap (1x1 conv, no bias), GroupNorm(32), ReLU,
//   gate = 1 + sigmoid(.)  (the "+1" is FFRM's residual: x * atten + x).
// grid (B, splits), block 1024: a CTA owns whole GroupNorm groups (their rows of the 1x1 conv), so the splits are
// independent; every CTA rebuilds the pooled vector (cheap) and CTA (b, 0) writes mu / rstd. With one CTA per image
// (round 1) eight SMs streamed the C x C fp32 weight alone: 172 us at C = 1536.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
ffrm_gate_kernel(const float* __restrict__ part, int nch, int B, int HW, int C, double sum_w, double mean_b, float ln_eps,
                 const float* __restrict__ wffrm, const float* __restrict__ gn_w, const float* __restrict__ gn_b,
                 int groups, float gn_eps, float* __restrict__ mu_out, float* __restrict__ rstd_out,
                 float* __restrict__ gate_out) {
  extern __shared__ float s_ff[];
  float* gap = s_ff;        // [C]
  float* a = gap + C;       // [C]
  const int b = blockIdx.x, t = threadIdx.x, warp = t >> 5, lane = t & 31, nwarps = blockDim.x >> 5;
  for (int c = t; c < C; c += blockDim.x) {
    double s0 = 0, s1 = 0, s2 = 0;
    for (int k = 0; k < nch; ++k) {
      const float* p = part + (((long long)k * B + b) * C + c) * 3;
      s0 += (double)p[0]; s1 += (double)p[1]; s2 += (double)p[2];
    }
    const double mu = s0 / HW;
    double var = s1 / HW - mu * mu;
    var = var > 0 ? var : 0;
    const double rstd = 1.0 / sqrt(var + (double)ln_eps);
    if (blockIdx.y == 0) {
      mu_out[(long long)b * C + c] = (float)mu;
      rstd_out[(long long)b * C + c] = (float)rstd;
    }
    gap[c] = (float)(rstd * (s2 - mu * sum_w) / HW + mean_b);
  }
  __syncthreads();
  const int cg = C / groups;
  const int gper = (groups + gridDim.y - 1) / gridDim.y;
  const int g_lo = blockIdx.y * gper, g_hi = min(g_lo + gper, groups);
  for (int c = g_lo * cg + warp; c < g_hi * cg; c += nwarps) {   // warp = one output channel of the 1x1 conv
    const float* w = wffrm + (long long)c * C;
    float acc = 0.f;
    for (int k = lane; k < C; k += 32) acc = fmaf(__ldg(w + k), gap[k], acc);
    acc = warp_sum(acc);
    if (lane == 0) a[c] = acc;
  }
  __syncthreads();
  for (int g = g_lo + warp; g < g_hi; g += nwarps) {     // warp = one GroupNorm group (biased variance)
    float s = 0.f;
    for (int k = lane; k < cg; k += 32) s += a[g * cg + k];
    const float m = warp_sum(s) / cg;
    float q = 0.f;
    for (int k = lane; k < cg; k += 32) { const float d = a[g * cg + k] - m; q = fmaf(d, d, q); }
    const float rs = rsqrtf(warp_sum(q) / cg + gn_eps);
    for (int k = lane; k < cg; k += 32) {
      const int c = g * cg + k;
      const float y = fmaxf((a[c] - m) * rs * gn_w[c] + gn_b[c], 0.f);
      gate_out[(long long)b * C + c] = 1.f + 1.f / (1.f + __expf(-y));
    }
  }
}

// ------------------------------------------------------------------------------------------------
// CoordinateAttention vectors (:187-215) from the pooled sums of combine_pool_kernel, per (position, image):
//   y = mean over the other axis [C] -> conv1 (C -> mip, bias) -> eval BN -> h_swish -> conv_h | conv_w (mip -> C) -> sigmoid
// position < H: a row (ah[b, y, :]), else a column (aw[b, x, :]). grid (H + W, B), block 256.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
ca_vectors_kernel(const float* __restrict__ ph, const float* __restrict__ pw_part, int nstrips, int B, int H, int W, int C,
                  int mip, const float* __restrict__ w1, const float* __restrict__ b1, const float* __restrict__ bn_s,
                  const float* __restrict__ bn_t, const float* __restrict__ wh, const float* __restrict__ bh,
                  const float* __restrict__ ww, const float* __restrict__ bw, float* __restrict__ ah,
                  float* __restrict__ aw) {
  extern __shared__ float s_ca[];
  float* y = s_ca;          // [C]
  float* tm = y + C;        // [mip]
  const int pos = blockIdx.x, b = blockIdx.y, t = threadIdx.x, warp = t >> 5, lane = t & 31, nwarps = blockDim.x >> 5;
  const bool is_h = pos < H;
  for (int c = t; c < C; c += blockDim.x) {
    float v;
    if (is_h) v = ph[((long long)b * H + pos) * C + c] / W;
    else {
      v = 0.f;
      // strips in order; the loads of four strips are requested together (the kernel is a chain of short latency-bound
      // phases: 126 us at the 256^2 level with one load in flight per thread)
      int s = 0;
      for (; s + 4 <= nstrips; s += 4) {
        const float p0 = pw_part[(((long long)s * B + b) * W + (pos - H)) * C + c];
        const float p1 = pw_part[(((long long)(s + 1) * B + b) * W + (pos - H)) * C + c];
        const float p2 = pw_part[(((long long)(s + 2) * B + b) * W + (pos - H)) * C + c];
        const float p3 = pw_part[(((long long)(s + 3) * B + b) * W + (pos - H)) * C + c];
        v += p0; v += p1; v += p2; v += p3;
      }
      for (; s < nstrips; ++s) v += pw_part[(((long long)s * B + b) * W + (pos - H)) * C + c];
      v /= H;
    }
    y[c] = v;
  }
  __syncthreads();
  for (int m = warp; m < mip; m += nwarps) {
    const float* w = w1 + (long long)m * C;
    float acc = 0.f;
#pragma unroll 8
    for (int k = lane; k < C; k += 32) acc = fmaf(__ldg(w + k), y[k], acc);
    acc = warp_sum(acc);
    if (lane == 0) {
      const float z = (acc + b1[m]) * bn_s[m] + bn_t[m];
      tm[m] = z * fminf(fmaxf(z + 3.f, 0.f), 6.f) * (1.f / 6.f);
    }
  }
  __syncthreads();
  const float* wo = is_h ? wh : ww;
  const float* bo = is_h ? bh : bw;
  float* dst = is_h ? ah + ((long long)b * H + pos) * C : aw + ((long long)b * W + (pos - H)) * C;
  for (int c = t; c < C; c += blockDim.x) {
    float acc = bo[c];
#pragma unroll 8
    for (int m = 0; m < mip; ++m) acc = fmaf(__ldg(wo + (long long)c * mip + m), tm[m], acc);
    dst[c] = 1.f / (1.f + __expf(-acc));
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
  extern __shared__ float s_ph[];  // [8 warps][RS][64]: every warp accumulates its own columns, combined in order at the end
  const int strip = blockIdx.x, cb = blockIdx.y, b = blockIdx.z;
  const int warp_id = threadIdx.x >> 5;
  const int cv = threadIdx.x & 7, xl = threadIdx.x >> 3;
  const int c = cb * 64 + cv * 8;
  const bool cok = c < C;
  const int y0 = strip * RS, y1 = min(y0 + RS, H);
  for (int i = threadIdx.x; i < 8 * RS * 64; i += blockDim.x) s_ph[i] = 0.f;
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
    // four rows of loads are requested before the first is consumed: with one (o, lo) pair in flight per thread and
    // ~2.6 CTAs per SM the kernel ran at 45 % of its HBM floor (208 us for 604 MB at the 256^2 level)
    for (int yb = y0; yb < y1; yb += 4) {
    uint4 ua[4], ul[4];
    float wp4[4], bp4[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      ua[q] = ul[q] = make_uint4(0, 0, 0, 0); wp4[q] = bp4[q] = 0.f;
      if (yb + q < y1 && x < W && cok) {
        const long long pix = (long long)(yb + q) * W + x;
        const long long off = ((long long)b * H * W + pix) * C + c;
        ua[q] = __ldg(reinterpret_cast<const uint4*>(o + off));
        ul[q] = __ldg(reinterpret_cast<const uint4*>(lo + off));
        wp4[q] = __ldg(wpix + pix); bp4[q] = __ldg(bpix + pix);
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int y = yb + q;
      if (y >= y1) break;                 // uniform over the CTA
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = 0.f;
      if (x < W && cok) {
        const long long pix = (long long)y * W + x;
        const long long off = ((long long)b * H * W + pix) * C + c;
        float a[8], l[8];
        unpack8(ua[q], a);
        unpack8(ul[q], l);
        const float wp = wp4[q], bp = bp4[q];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          v[j] = fmaf(fmaf((a[j] - m[j]) * r[j], wp, bp), g[j], s2 * l[j]);
          col[j] += v[j];
        }
        *reinterpret_cast<uint4*>(f + off) = pack8(v);
      }
      // row sums: reduce the 4 columns held by this warp, accumulate into the warp's private shared-memory row
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float t = v[j];
        t += __shfl_xor_sync(0xffffffffu, t, 8);
        t += __shfl_xor_sync(0xffffffffu, t, 16);
        if ((threadIdx.x & 31) < 8) s_ph[(warp_id * RS + (y - y0)) * 64 + cv * 8 + j] += t;
      }
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
    float a = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) a += s_ph[w * RS * 64 + i];
    if (cc < C) ph[((long long)b * H + y0 + yy) * C + cc] = a;
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
int mmsam_gram_tc_plan(int n, int B, int HW, int norms, int* chunk_out);   // gram_tc.cu

static int gram_simt_forced() {
  static const int v = getenv("MMSAM_GRAM_SIMT") != nullptr;
  return v;
}
// pixel chunking of the SIMT fallback (maps whose pixel count is not a multiple of 64)
static int gram_simt_plan(int HW, int* chunk_out) {
  int nchunks = HW / 2048;
  if (nchunks < 1) nchunks = 1;
  if (nchunks > 16) nchunks = 16;
  int chunk = (HW + nchunks - 1) / nchunks;
  chunk = (chunk + 31) / 32 * 32;
  if (chunk_out) *chunk_out = chunk;
  return (HW + chunk - 1) / chunk;
}

MMSAM_API int mmsam_gram_chunks(int n, int B, int HW, int norms) {
  if (n <= 0 || B <= 0 || HW <= 0) return 0;
  if (!gram_simt_forced()) {
    const int c = mmsam_gram_tc_plan(n, B, HW, norms, nullptr);
    if (c > 0) return c;
  }
  return gram_simt_plan(HW, nullptr);
}

MMSAM_API int mmsam_gram_bf16(const void* X, long long ld, int qoff, int koff, int n, int B, int HW, int blk,
                              float* S_part, float* nq_part, float* nk_part, void* stream) {
  using namespace mmsam;
  if (B < 0 || n <= 0 || HW <= 0 || (ld & 7) || (qoff & 7) || (koff & 7) || (n & 7) || blk < 0) return MMSAM_ERR_BAD_ARG;
  if ((nq_part == nullptr) != (nk_part == nullptr)) return MMSAM_ERR_BAD_ARG;
  if (B == 0) return MMSAM_OK;
  if (!X || !S_part || (((uintptr_t)X) & 15)) return MMSAM_ERR_BAD_ARG;
  if (!gram_simt_forced()) {   // tensor-core path; shapes it does not take fall through to the SIMT kernel
    const int rc = mmsam_gram_tc(X, ld, qoff, koff, n, B, HW, blk, S_part, nq_part, nk_part, (cudaStream_t)stream);
    if (rc != MMSAM_ERR_UNSUPPORTED) return rc;
  }
  const int nt = (n + 63) / 64;
  int chunk = 0;
  const int nchunks = gram_simt_plan(HW, &chunk);
  dim3 grid(nt * nt, nchunks, B);
  gram_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)X, ld, qoff, koff, n, HW, chunk, blk, S_part,
                                                      nq_part, nk_part);
  MMSAM_LAUNCH_CHECK();
  return MMSAM_OK;
}

MMSAM_API int mmsam_gfe_weff_bf16(const float* S_part, const float* nq_part, const float* nk_part, int nchunks, int B, int ci,
                                  int heads, const float* temperature, const float* wproj, const float* scale2, void* weff,
                                  void* stream) {
  using namespace mmsam;
  if (B < 0 || ci <= 0 || heads <= 0 || ci % heads || nchunks <= 0) return MMSAM_ERR_BAD_ARG;
  if (B == 0) return MMSAM_OK;
  if (!S_part || !nq_part || !nk_part || !temperature || !wproj || !scale2 || !weff) return MMSAM_ERR_BAD_ARG;
  const int ch = ci / heads;
  const int smem = (ch * (ch + 4) + 2 * ch) * (int)sizeof(float);
  if (smem > 48 * 1024) return MMSAM_ERR_UNSUPPORTED;
  int zs = (2 * kNumSMs + heads * B - 1) / (heads * B);      // row splits: ~2 CTAs per SM (more splits repeat the partial sums and the softmax: 4 per SM measured slower)
  if (zs > (ci + 15) / 16) zs = (ci + 15) / 16;
  if (zs < 1) zs = 1;
  dim3 grid(heads, B, zs);
  gfe_weff_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(S_part, nq_part, nk_part, nchunks, B, ci, heads, temperature,
                                                             wproj, scale2, (__nv_bfloat16*)weff);
  MMSAM_LAUNCH_CHECK();
  return MMSAM_OK;
}

MMSAM_API int mmsam_gffm_softmax_bf16(const float* E_part, int nchunks, int B, int ci, void* ax, void* ay, void* stream) {
  using namespace mmsam;
  if (B < 0 || ci <= 0 || ci > 1024 || nchunks <= 0) return MMSAM_ERR_BAD_ARG;
  if (B == 0) return MMSAM_OK;
  if (!E_part || !ax || !ay) return MMSAM_ERR_BAD_ARG;
  dim3 grid(ci, B, 2);
  gffm_softmax_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(E_part, nchunks, B, ci, (__nv_bfloat16*)ax, (__nv_bfloat16*)ay);
  MMSAM_LAUNCH_CHECK();
  return MMSAM_OK;
}

MMSAM_API int mmsam_ffrm_gate_f32(const float* colstats_part, int nchunks, int B, int HW, int C, double sum_w, double mean_b,
                                  float ln_eps, const float* wffrm, const float* gn_weight, const float* gn_bias, int groups,
                                  float gn_eps, float* mu, float* rstd, float* gate, void* stream) {
  using namespace mmsam;
  if (B < 0 || HW <= 0 || C <= 0 || groups <= 0 || C % groups || nchunks <= 0 || C > 4096) return MMSAM_ERR_BAD_ARG;
  if (B == 0) return MMSAM_OK;
  if (!colstats_part || !wffrm || !gn_weight || !gn_bias || !mu || !rstd || !gate) return MMSAM_ERR_BAD_ARG;
  int splits = groups < 16 ? groups : 16;          // whole groups per CTA; every split re-reads the colstats partials
  ffrm_gate_kernel<<<dim3(B, splits), 1024, 2 * C * sizeof(float), (cudaStream_t)stream>>>(colstats_part, nchunks, B, HW, C, sum_w, mean_b,
                                                                             ln_eps, wffrm, gn_weight, gn_bias, groups, gn_eps,
                                                                             mu, rstd, gate);
  MMSAM_LAUNCH_CHECK();
  return MMSAM_OK;
}

MMSAM_API int mmsam_ca_vectors_f32(const float* ph, const float* pw_part, int nstrips, int B, int H, int W, int C, int mip,
                                   const float* w1, const float* b1, const float* bn_scale, const float* bn_shift,
                                   const float* wh, const float* bh, const float* ww, const float* bw, float* ah, float* aw,
                                   void* stream) {
  using namespace mmsam;
  if (B < 0 || H <= 0 || W <= 0 || C <= 0 || mip <= 0 || nstrips <= 0 || (C + mip) * 4 > 48 * 1024) return MMSAM_ERR_BAD_ARG;
  if (B == 0) return MMSAM_OK;
  if (!ph || !pw_part || !w1 || !b1 || !bn_scale || !bn_shift || !wh || !bh || !ww || !bw || !ah || !aw) return MMSAM_ERR_BAD_ARG;
  dim3 grid(H + W, B);
  ca_vectors_kernel<<<grid, 256, (C + mip) * sizeof(float), (cudaStream_t)stream>>>(ph, pw_part, nstrips, B, H, W, C, mip, w1, b1,
                                                                                    bn_scale, bn_shift, wh, bh, ww, bw, ah, aw);
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
  combine_pool_kernel<<<grid, 256, 8 * RS * 64 * sizeof(float), (cudaStream_t)stream>>>(
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
