// Multi-scale deformable attention sampling (forward) for sm_100a.
//
// Replaces the reference's ms_deformable_im2col_gpu_kernel
// (segmentation/ops/src/cuda/ms_deform_im2col_cuda.cuh:237-299, bilinear helper :33-84) and its
// host wrapper ms_deform_attn_cuda_forward (segmentation/ops/src/cuda/ms_deform_attn_cuda.cu:20-80).
//
//   out[n,q,m,:] = sum_l sum_p w[n,q,m,l,p] * bilinear(value_l[n,:,m,:], loc[n,q,m,l,p])
//
// with the reference's conventions: h_im = loc_y*H - 0.5, w_im = loc_x*W - 0.5 (align_corners =
// False), zero padding outside the map, loc = (x, y) order.
//
// Two kernels:
//  * msda_vec_kernel  - the hot one. A thread owns VEC contiguous channels of one (query, head)
//    pair (16-byte gathers); a CTA owns QB consecutive queries x HB heads so the gathered footprint
//    (a band of the value map for a few heads) stays L1-resident; no pre-zeroing, one 16 B store
//    per thread.
//  * msda_generic_kernel - any D / dtype (incl. fp64), one thread per output element; used for the
//    reference's own known-answer shapes (D = 2) and as the catch-all.
#include "common.cuh"
#include "cg2.cuh"
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <type_traits>

namespace mmsam {

template <typename T> struct Acc { using type = float; };
template <> struct Acc<double> { using type = double; };

template <typename T> __device__ __forceinline__ typename Acc<T>::type ldv(const T* p) { return (typename Acc<T>::type)(*p); }
template <> __device__ __forceinline__ float ldv<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <> __device__ __forceinline__ float ldv<__half>(const __half* p) { return __half2float(*p); }

template <typename T> __device__ __forceinline__ void stv(T* p, typename Acc<T>::type v) { *p = (T)v; }
template <> __device__ __forceinline__ void stv<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ void stv<__half>(__half* p, float v) { *p = __float2half_rn(v); }

// ------------------------------------------------------------------------------------------------
// generic kernel
// ------------------------------------------------------------------------------------------------
template <typename VT, typename AT>
__global__ void __launch_bounds__(256)
msda_generic_kernel(const VT* __restrict__ value, const int64_t* __restrict__ shapes,
                    const int64_t* __restrict__ lsi, const AT* __restrict__ loc,
                    const AT* __restrict__ attw, VT* __restrict__ out, int N, int S, int M, int D,
                    int Lq, int L, int P) {
  using A = typename Acc<VT>::type;
  const long long total = (long long)N * Lq * M * D;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % D);
    long long t = idx / D;
    const int m = (int)(t % M);
    t /= M;
    const int q = (int)(t % Lq);
    const int n = (int)(t / Lq);
    const long long pair = ((long long)n * Lq + q) * M + m;
    const AT* lp = loc + pair * L * P * 2;
    const AT* wp = attw + pair * L * P;
    const VT* vbase = value + (long long)n * S * M * D + (long long)m * D + c;
    A acc = 0;
    for (int l = 0; l < L; ++l) {
      const int H = (int)shapes[2 * l], W = (int)shapes[2 * l + 1];
      const VT* vl = vbase + (long long)lsi[l] * M * D;
      for (int p = 0; p < P; ++p) {
        const A lx = (A)ldv(lp + (l * P + p) * 2), ly = (A)ldv(lp + (l * P + p) * 2 + 1);
        const A aw = (A)ldv(wp + l * P + p);
        const A h_im = ly * H - (A)0.5, w_im = lx * W - (A)0.5;
        if (h_im > -1 && w_im > -1 && h_im < H && w_im < W) {
          const int h0 = (int)floor(h_im), w0 = (int)floor(w_im);
          const A lh = h_im - h0, lw = w_im - w0, hh = 1 - lh, hw = 1 - lw;
          A v1 = 0, v2 = 0, v3 = 0, v4 = 0;
          const long long rs = (long long)M * D;
          if (h0 >= 0 && w0 >= 0) v1 = ldv(vl + ((long long)h0 * W + w0) * rs);
          if (h0 >= 0 && w0 + 1 <= W - 1) v2 = ldv(vl + ((long long)h0 * W + w0 + 1) * rs);
          if (h0 + 1 <= H - 1 && w0 >= 0) v3 = ldv(vl + ((long long)(h0 + 1) * W + w0) * rs);
          if (h0 + 1 <= H - 1 && w0 + 1 <= W - 1) v4 = ldv(vl + ((long long)(h0 + 1) * W + w0 + 1) * rs);
          acc += aw * (hh * hw * v1 + hh * lw * v2 + lh * hw * v3 + lh * lw * v4);
        }
      }
    }
    stv(out + idx, acc);
  }
}

// ------------------------------------------------------------------------------------------------
// vectorised kernel: 16-byte gathers
// ------------------------------------------------------------------------------------------------
template <typename VT> struct Vec16;  // 16 bytes of VT
template <> struct Vec16<__nv_bfloat16> {
  static constexpr int N = 8;
  __device__ static void load(const __nv_bfloat16* p, float* f) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
    unpack8(v, f);
  }
  __device__ static void store(__nv_bfloat16* p, const float* f) {
    *reinterpret_cast<uint4*>(p) = pack8(f);
  }
};
template <> struct Vec16<__half> {
  static constexpr int N = 8;
  __device__ static void load(const __half* p, float* f) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
    const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) { float2 t = __half22float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
  }
  __device__ static void store(__half* p, const float* f) {
    uint4 v;
    __half2* h = reinterpret_cast<__half2*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = v;
  }
};
template <> struct Vec16<float> {
  static constexpr int N = 4;
  __device__ static void load(const float* p, float* f) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(p));
    f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
  }
  __device__ static void store(float* p, const float* f) {
    *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
  }
};

// One CTA: QB consecutive queries x HB heads; TPP = D / VEC threads per (query, head) pair.
// blockDim.x = QB * HB * TPP. gridDim = (ceil(Lq / QB), M / HB, N).
template <typename VT, typename AT, int P>
__global__ void __launch_bounds__(256)
msda_vec_kernel(const VT* __restrict__ value, const int64_t* __restrict__ shapes,
                const int64_t* __restrict__ lsi, const AT* __restrict__ loc,
                const AT* __restrict__ attw, VT* __restrict__ out, int S, int M, int D, int Lq, int L,
                int HB, int TPP) {
  constexpr int VEC = Vec16<VT>::N;
  const int n = blockIdx.z;
  const int t = threadIdx.x;
  const int chunk = t % TPP;
  const int pr = t / TPP;
  const int hi = pr % HB;
  const int qi = pr / HB;
  const int q = blockIdx.x * (blockDim.x / (HB * TPP)) + qi;
  const int m = blockIdx.y * HB + hi;
  if (q >= Lq) return;
  const long long pair = ((long long)n * Lq + q) * M + m;
  const AT* lp = loc + pair * L * P * 2;
  const AT* wp = attw + pair * L * P;
  const long long rs = (long long)M * D;  // elements between neighbouring spatial positions
  const VT* vbase = value + (long long)n * S * rs + (long long)m * D + chunk * VEC;

  float acc[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) acc[i] = 0.f;

  for (int l = 0; l < L; ++l) {
    const int H = (int)__ldg(shapes + 2 * l), W = (int)__ldg(shapes + 2 * l + 1);
    const VT* vl = vbase + (long long)__ldg(lsi + l) * rs;
    float lx[P], ly[P], aw[P];
#pragma unroll
    for (int p = 0; p < P; ++p) {
      lx[p] = (float)ldv(lp + (l * P + p) * 2);
      ly[p] = (float)ldv(lp + (l * P + p) * 2 + 1);
      aw[p] = (float)ldv(wp + l * P + p);
    }
#pragma unroll
    for (int p = 0; p < P; ++p) {
      const float h_im = ly[p] * H - 0.5f, w_im = lx[p] * W - 0.5f;
      if (h_im > -1.f && w_im > -1.f && h_im < H && w_im < W) {
        const float hf = floorf(h_im), wf = floorf(w_im);
        const int h0 = (int)hf, w0 = (int)wf;
        const float lh = h_im - hf, lw = w_im - wf, hh = 1.f - lh, hw = 1.f - lw;
        const float c1 = aw[p] * hh * hw, c2 = aw[p] * hh * lw, c3 = aw[p] * lh * hw,
                    c4 = aw[p] * lh * lw;
        const bool t0 = h0 >= 0, b0 = h0 + 1 <= H - 1, l0 = w0 >= 0, r0 = w0 + 1 <= W - 1;
        const VT* p00 = vl + ((long long)h0 * W + w0) * rs;
        float v[VEC];
        if (t0 && l0) {
          Vec16<VT>::load(p00, v);
#pragma unroll
          for (int i = 0; i < VEC; ++i) acc[i] = fmaf(c1, v[i], acc[i]);
        }
        if (t0 && r0) {
          Vec16<VT>::load(p00 + rs, v);
#pragma unroll
          for (int i = 0; i < VEC; ++i) acc[i] = fmaf(c2, v[i], acc[i]);
        }
        if (b0 && l0) {
          Vec16<VT>::load(p00 + (long long)W * rs, v);
#pragma unroll
          for (int i = 0; i < VEC; ++i) acc[i] = fmaf(c3, v[i], acc[i]);
        }
        if (b0 && r0) {
          Vec16<VT>::load(p00 + (long long)W * rs + rs, v);
#pragma unroll
          for (int i = 0; i < VEC; ++i) acc[i] = fmaf(c4, v[i], acc[i]);
        }
      }
    }
  }
  Vec16<VT>::store(out + pair * D + chunk * VEC, acc);
}

// Fused front end used by the backbone's Injector / Extractor path: instead of materialised
// sampling_locations / attention_weights it reads the raw output of ONE query projection GEMM
// (ops/modules/ms_deform_attn.py:108-119: sampling_offsets | attention_weights Linear outputs, fp32,
// row = query, columns = [M*L*P*2 offsets | M*L*P logits]) plus the reference point of each query,
// and does the softmax over L*P and  loc = ref + off / (W_l, H_l)  in registers.
template <typename VT, int P>
__global__ void __launch_bounds__(256)
msda_fused_kernel(const VT* __restrict__ value, const int64_t* __restrict__ shapes,
                  const int64_t* __restrict__ lsi, const float* __restrict__ qproj, long long ldq,
                  const float* __restrict__ ref, VT* __restrict__ out, int S, int M, int D, int Lq, int L,
                  int HB, int TPP) {
  constexpr int VEC = Vec16<VT>::N;
  const int n = blockIdx.z;
  const int t = threadIdx.x;
  const int chunk = t % TPP;
  const int pr = t / TPP;
  const int hi = pr % HB;
  const int qi = pr / HB;
  const int q = blockIdx.x * (blockDim.x / (HB * TPP)) + qi;
  const int m = blockIdx.y * HB + hi;
  if (q >= Lq) return;
  const long long row = (long long)n * Lq + q;
  const float* offp = qproj + row * ldq + (long long)m * L * P * 2;
  const float* lgp = qproj + row * ldq + (long long)M * L * P * 2 + (long long)m * L * P;
  const float rx = __ldg(ref + 2 * q), ry = __ldg(ref + 2 * q + 1);
  const long long rs = (long long)M * D;
  const VT* vbase = value + (long long)n * S * rs + (long long)m * D + chunk * VEC;

  // softmax statistics over the L*P logits of this (query, head)
  float mx = -INFINITY;
  for (int i = 0; i < L * P; ++i) mx = fmaxf(mx, __ldg(lgp + i));
  float den = 0.f;
  for (int i = 0; i < L * P; ++i) den += __expf(__ldg(lgp + i) - mx);
  const float inv = 1.f / den;

  float acc[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) acc[i] = 0.f;
  for (int l = 0; l < L; ++l) {
    const int H = (int)__ldg(shapes + 2 * l), W = (int)__ldg(shapes + 2 * l + 1);
    const VT* vl = vbase + (long long)__ldg(lsi + l) * rs;
    float ox[P], oy[P], aw[P];
#pragma unroll
    for (int p = 0; p < P; ++p) {
      const float2 o = __ldg(reinterpret_cast<const float2*>(offp + (l * P + p) * 2));
      ox[p] = o.x; oy[p] = o.y;
      aw[p] = __expf(__ldg(lgp + l * P + p) - mx) * inv;
    }
#pragma unroll
    for (int p = 0; p < P; ++p) {
      const float h_im = (ry + oy[p] / H) * H - 0.5f, w_im = (rx + ox[p] / W) * W - 0.5f;
      if (h_im > -1.f && w_im > -1.f && h_im < H && w_im < W) {
        const float hf = floorf(h_im), wf = floorf(w_im);
        const int h0 = (int)hf, w0 = (int)wf;
        const float lh = h_im - hf, lw = w_im - wf, hh = 1.f - lh, hw = 1.f - lw;
        const float c1 = aw[p] * hh * hw, c2 = aw[p] * hh * lw, c3 = aw[p] * lh * hw, c4 = aw[p] * lh * lw;
        const bool t0 = h0 >= 0, b0 = h0 + 1 <= H - 1, l0 = w0 >= 0, r0 = w0 + 1 <= W - 1;
        const VT* p00 = vl + ((long long)h0 * W + w0) * rs;
        float v[VEC];
        if (t0 && l0) {
          Vec16<VT>::load(p00, v);
#pragma unroll
          for (int i = 0; i < VEC; ++i) acc[i] = fmaf(c1, v[i], acc[i]);
        }
        if (t0 && r0) {
          Vec16<VT>::load(p00 + rs, v);
#pragma unroll
          for (int i = 0; i < VEC; ++i) acc[i] = fmaf(c2, v[i], acc[i]);
        }
        if (b0 && l0) {
          Vec16<VT>::load(p00 + (long long)W * rs, v);
#pragma unroll
          for (int i = 0; i < VEC; ++i) acc[i] = fmaf(c3, v[i], acc[i]);
        }
        if (b0 && r0) {
          Vec16<VT>::load(p00 + (long long)W * rs + rs, v);
#pragma unroll
          for (int i = 0; i < VEC; ++i) acc[i] = fmaf(c4, v[i], acc[i]);
        }
      }
    }
  }
  Vec16<VT>::store(out + (row * M + m) * D + chunk * VEC, acc);
}

// Cooperative fused kernel for the hot shapes (bf16, D = 32, P = 4, L <= 4): a warp owns ONE query x 8 heads,
// the 4 lanes of a head are (a) the 4 sampling points of a level while locations / softmax weights are computed
// (one point per lane, softmax statistics by two width-4 shuffles) and (b) the 4 16-byte channel chunks while
// gathering (each point's base index + 4 corner weights are broadcast inside the 4-lane group). Compared with
// msda_fused_kernel (every lane recomputes all points of its head): the query-projection row is read as
// contiguous 256 B / 128 B pieces (3 L1 wavefronts per level instead of ~70), 4x less address arithmetic, and
// the gathers are branch-free (clamped indices, zero weights outside the map). The kernel then sits on the SM's
// gather rate for 64-byte units (tools/micro/gather_bench.cu: ~1.6 cycles per unit, L1 hit or L2 hit alike).
template <int MAXL>
__global__ void __launch_bounds__(256)
msda_fused_coop_kernel(const __nv_bfloat16* __restrict__ value, const int64_t* __restrict__ shapes,
                       const int64_t* __restrict__ lsi, const float* __restrict__ qproj, long long ldq,
                       const float* __restrict__ ref, __nv_bfloat16* __restrict__ out, int S, int M, int Lq, int L) {
  constexpr int P = 4, D = 32;
  const int lane = threadIdx.x & 31;
  const int g = lane >> 2, pl = lane & 3;
  const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int n = blockIdx.z;
  const int m = blockIdx.y * 8 + g;
  if (q >= Lq) return;                       // warp-uniform
  const bool head_ok = m < M;
  const int mc = head_ok ? m : M - 1;        // inactive head groups shadow the last head, their store is skipped
  const long long row = (long long)n * Lq + q;
  const float* offp = qproj + row * ldq + (long long)mc * L * P * 2;
  const float* lgp = qproj + row * ldq + (long long)M * L * P * 2 + (long long)mc * L * P;
  const float rx = __ldg(ref + 2 * q), ry = __ldg(ref + 2 * q + 1);
  const long long rs = (long long)M * D;
  const __nv_bfloat16* vbase = value + (long long)n * S * rs + (long long)mc * D + pl * 8;

  // this lane's point of every level: logits -> softmax over the L*P points of the head
  float lg[MAXL];
  float2 of[MAXL];
  float mx = -INFINITY;
#pragma unroll
  for (int l = 0; l < MAXL; ++l) {
    lg[l] = -INFINITY;
    of[l] = make_float2(0.f, 0.f);
    if (l < L) {
      lg[l] = __ldg(lgp + l * P + pl);
      of[l] = __ldg(reinterpret_cast<const float2*>(offp + (l * P + pl) * 2));
      mx = fmaxf(mx, lg[l]);
    }
  }
  mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
  mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
  float den = 0.f;
#pragma unroll
  for (int l = 0; l < MAXL; ++l) {
    lg[l] = l < L ? __expf(lg[l] - mx) : 0.f;
    den += lg[l];
  }
  den += __shfl_xor_sync(0xffffffffu, den, 1);
  den += __shfl_xor_sync(0xffffffffu, den, 2);
  const float inv = 1.f / den;

  // packed fp32x2 accumulators (channel pairs): one FFMA2 per corner and pair instead of two FFMA — the kernel issues at
  // 68 - 82 % of its slots (ncu), so instructions, not only L1 wavefronts, are on the critical path
  u64 acc2[4] = {0ull, 0ull, 0ull, 0ull};
  const uint32_t rs32 = (uint32_t)rs;
#pragma unroll
  for (int l = 0; l < MAXL; ++l) {
    if (l < L) {
      const int H = (int)__ldg(shapes + 2 * l), W = (int)__ldg(shapes + 2 * l + 1);
      const __nv_bfloat16* vl = vbase + (long long)__ldg(lsi + l) * rs;
      // my point: same arithmetic as the reference (loc = ref + off / (W, H); im = loc * size - 0.5)
      const float h_im = (ry + of[l].y / H) * H - 0.5f, w_im = (rx + of[l].x / W) * W - 0.5f;
      const bool inside = h_im > -1.f && w_im > -1.f && h_im < H && w_im < W;
      const float hf = floorf(h_im), wf = floorf(w_im);
      const int h0 = (int)hf, w0 = (int)wf;
      const float lh = h_im - hf, lw = w_im - wf, hh = 1.f - lh, hw = 1.f - lw;
      const float aw = inside ? lg[l] * inv : 0.f;
      const bool t0 = h0 >= 0, b0 = h0 + 1 <= H - 1, l0 = w0 >= 0, r0 = w0 + 1 <= W - 1;
      float c1 = (t0 && l0) ? aw * hh * hw : 0.f, c2 = (t0 && r0) ? aw * hh * lw : 0.f;
      float c3 = (b0 && l0) ? aw * lh * hw : 0.f, c4 = (b0 && r0) ? aw * lh * lw : 0.f;
      // clamped corner indices: every gather reads valid memory, invalid corners carry a zero weight
      const int h0c = min(max(h0, 0), H - 1), h1c = min(max(h0 + 1, 0), H - 1);
      const int w0c = min(max(w0, 0), W - 1), w1c = min(max(w0 + 1, 0), W - 1);
      int i00 = inside ? h0c * W + w0c : 0;
      int dflag = inside ? ((w1c - w0c) | ((h1c - h0c) << 1)) : 0;
#pragma unroll
      for (int pp = 0; pp < P; ++pp) {
        const int bi = __shfl_sync(0xffffffffu, i00, pp, 4);
        const int bf = __shfl_sync(0xffffffffu, dflag, pp, 4);
        const float k1 = __shfl_sync(0xffffffffu, c1, pp, 4), k2 = __shfl_sync(0xffffffffu, c2, pp, 4);
        const float k3 = __shfl_sync(0xffffffffu, c3, pp, 4), k4 = __shfl_sync(0xffffffffu, c4, pp, 4);
        // 32-bit element offsets inside the level (a level of one image is far below 2^31 elements)
        const uint32_t o00 = (uint32_t)bi * rs32;
        const uint32_t dx = (bf & 1) ? rs32 : 0u, dy = (bf & 2) ? (uint32_t)W * rs32 : 0u;
        const uint4 u00 = __ldg(reinterpret_cast<const uint4*>(vl + o00));
        const uint4 u01 = __ldg(reinterpret_cast<const uint4*>(vl + (o00 + dx)));
        const uint4 u10 = __ldg(reinterpret_cast<const uint4*>(vl + (o00 + dy)));
        const uint4 u11 = __ldg(reinterpret_cast<const uint4*>(vl + (o00 + dy + dx)));
        const u64 w1 = pack2(k1, k1), w2 = pack2(k2, k2), w3 = pack2(k3, k3), w4 = pack2(k4, k4);
        const uint32_t* c00 = reinterpret_cast<const uint32_t*>(&u00);
        const uint32_t* c01 = reinterpret_cast<const uint32_t*>(&u01);
        const uint32_t* c10 = reinterpret_cast<const uint32_t*>(&u10);
        const uint32_t* c11 = reinterpret_cast<const uint32_t*>(&u11);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          acc2[i] = fma2(pack2(bf16lo(c00[i]), bf16hi(c00[i])), w1, acc2[i]);
          acc2[i] = fma2(pack2(bf16lo(c01[i]), bf16hi(c01[i])), w2, acc2[i]);
          acc2[i] = fma2(pack2(bf16lo(c10[i]), bf16hi(c10[i])), w3, acc2[i]);
          acc2[i] = fma2(pack2(bf16lo(c11[i]), bf16hi(c11[i])), w4, acc2[i]);
        }
      }
    }
  }
  float acc[8];
#pragma unroll
  for (int i = 0; i < 4; ++i) unpack2(acc2[i], acc[2 * i], acc[2 * i + 1]);
  if (head_ok) *reinterpret_cast<uint4*>(out + (row * M + m) * D + pl * 8) = pack8(acc);
}


// ------------------------------------------------------------------------------------------------
// Shared-memory staged fused kernel (bf16, D = 32, P = 4, L <= 4, M <= 16).
//
// The cooperative kernel above is capped by the L1 gather rate (~1.6 cycles per 64-byte unit). Shared memory serves
// the same unit in ~0.54 cycles when the two 4-lane groups of a quarter warp read opposite bank halves
// (tools/micro/gather_bench.cu mode 9). So: a CTA owns ONE head of a REGION of the normalised plane (a tile of an
// anchor grid given by the host) and
//   * stages, with one TMA box per level, the window of that head's value map the region's queries are expected to
//     sample: origin = region corner (in level pixels) + the head's prior offset (min over points of the
//     sampling_offsets bias, ops/modules/ms_deform_attn.py:64-74) - margin. The box is dense [y][x][32 ch] = 64 B per
//     pixel, so horizontally adjacent pixels alternate bank halves, and out-of-map pixels are zero-filled by the
//     TMA unit — exactly the reference's zero padding (ms_deform_im2col_cuda.cuh:33-84);
//   * handles every query of every query grid whose reference point falls inside the region (extractor: the 128^2,
//     64^2 and 32^2 token grids all sample the same 64^2 map, so one staged window serves all three).
// 8 lanes own a query: first lane j computes points j and j+8 (location, softmax weight, bilinear factors, byte offset
// in the box), then for each point lanes 0-3 / 4-7 gather the left / right corner column (16 B channel chunks, top and
// bottom row) — the left and right pixels sit in opposite bank halves, so every LDS.128 phase is conflict-free.
// A sample whose corners leave the staged box (offsets far from the prior) is gathered from global memory with the
// reference's bounds logic: correctness never depends on the prior, only speed does.
struct MsdaStagedParams {
  const __nv_bfloat16* value;
  const float* qproj;
  const float* ref;
  __nv_bfloat16* out;
  long long ldq;
  int S, M, Lq, L, margin;
  int H[4], W[4], lsi[4], BW[4], BH[4], box_off[4];
  float pmin[4][16][2];                  // prior minimum offset (x, y), level pixels, per level and head
  int AH, AW, TAH, TAW, tiles_x, NG;     // anchor grid, region tile (anchor cells), query grids
  int gh[3], gw[3], gstart[3];
  unsigned tx_bytes, bar_off;
  unsigned qoff_off, qlg_off, qref_off, qidx_off;   // per-query staging areas: offsets, logits, ref point, query index
  unsigned rec_off, xy_off, in_off, seg_off, sref_off;   // msda_staged2_kernel: per-warp point records / fall-back
                                                         // coordinates / input ring; per-CTA pass list / reference points
  int dbg;                               // perf-debug switches (MMSAM_MSDA_DBG): 1 = no TMA staging, 2 = no gather
};
struct MsdaMaps { CUtensorMap m[4]; };
struct MsdaMaps2 { CUtensorMap m[4]; CUtensorMap qoff, qlg; };   // + the projection matrix: offsets / logits boxes of 8 rows

__device__ __forceinline__ int cdiv_pos(int num, int den) { return (num + den - 1) / den; }   // num >= 1 - den
__device__ __forceinline__ void cp_async_cg16(void* smem_dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_ca8(void* smem_dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(smem_dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

template <int MAXL>
__global__ void __launch_bounds__(512)
msda_staged_kernel(const __grid_constant__ MsdaMaps maps, const MsdaStagedParams p) {
  constexpr int NK = MAXL > 2 ? 2 : 1;   // points per lane (lane j: points j, j + 8)
  // [box level 0 | box level 1 | ... | per-query offsets | logits | ref points | query indices | mbarrier]
  extern __shared__ __align__(128) uint8_t st_smem[];
  uint64_t& bar = *reinterpret_cast<uint64_t*>(st_smem + p.bar_off);
  const int lane = threadIdx.x & 31, sub = lane & 7, side = sub >> 2, ch = sub & 3;
  const int grp = threadIdx.x >> 3, ngrp = blockDim.x >> 3;
  const int m = blockIdx.y, n = blockIdx.z;
  const int ty = blockIdx.x / p.tiles_x, tx = blockIdx.x - ty * p.tiles_x;
  const int X0 = tx * p.TAW, X1 = min(X0 + p.TAW, p.AW), Y0 = ty * p.TAH, Y1 = min(Y0 + p.TAH, p.AH);
  const float fx0 = (float)X0 / (float)p.AW, fy0 = (float)Y0 / (float)p.AH;

  if (threadIdx.x == 0 && !(p.dbg & 1)) {
    mbar_init(&bar, 1);
    fence_barrier_init();
    mbar_arrive_expect_tx(&bar, p.tx_bytes);
    for (int l = 0; l < p.L; ++l) {
      const int ox = (int)floorf(fx0 * p.W[l] - 0.5f + p.pmin[l][m][0]) - p.margin;
      const int oy = (int)floorf(fy0 * p.H[l] - 0.5f + p.pmin[l][m][1]) - p.margin;
      tma_load_4d(st_smem + p.box_off[l], &maps.m[l], &bar, m * 32, ox, oy, n);
    }
  }

  // the region's rectangle of every query grid: queries whose cell centre (i + 0.5) / g lies in [X0, X1) / AW
  int cnt[3], qx0[3], qy0[3], qw[3], total = 0;
#pragma unroll
  for (int g = 0; g < 3; ++g) {
    cnt[g] = 0; qx0[g] = qy0[g] = 0; qw[g] = 1;
    if (g < p.NG) {
      const int ix0 = cdiv_pos(2 * X0 * p.gw[g] - p.AW, 2 * p.AW), ix1 = cdiv_pos(2 * X1 * p.gw[g] - p.AW, 2 * p.AW);
      const int iy0 = cdiv_pos(2 * Y0 * p.gh[g] - p.AH, 2 * p.AH), iy1 = cdiv_pos(2 * Y1 * p.gh[g] - p.AH, 2 * p.AH);
      qx0[g] = ix0; qy0[g] = iy0; qw[g] = max(ix1 - ix0, 1);
      cnt[g] = (ix1 - ix0) * (iy1 - iy0);
      total += cnt[g];
    }
  }

  const int LP = p.L * 4;
  // stage this head's slice of every query's projection row (LP float2 offsets, LP logits), its reference point and
  // its index with asynchronous copies: all of them are in flight at once, none holds a register
  float2* s_off = reinterpret_cast<float2*>(st_smem + p.qoff_off);
  float* s_lg = reinterpret_cast<float*>(st_smem + p.qlg_off);
  float2* s_ref = reinterpret_cast<float2*>(st_smem + p.qref_off);
  int* s_q = reinterpret_cast<int*>(st_smem + p.qidx_off);
  for (int j = threadIdx.x; j < total; j += blockDim.x) {
    int r = j, q = 0;
    bool found = false;
#pragma unroll
    for (int g = 0; g < 3; ++g) {
      if (!found) {
        if (r < cnt[g]) {
          const int iy = r / qw[g], ix = r - iy * qw[g];
          q = p.gstart[g] + (qy0[g] + iy) * p.gw[g] + qx0[g] + ix;
          found = true;
        } else {
          r -= cnt[g];
        }
      }
    }
    const long long row = (long long)n * p.Lq + q;
    const float* offp = p.qproj + row * p.ldq + (long long)m * LP * 2;
    const float* lgp = p.qproj + row * p.ldq + (long long)p.M * LP * 2 + (long long)m * LP;
    for (int c = 0; c < LP / 2; ++c) cp_async_cg16(s_off + j * LP + 2 * c, offp + 4 * c);
    for (int c = 0; c < LP / 4; ++c) cp_async_cg16(s_lg + j * LP + 4 * c, lgp + 4 * c);
    cp_async_ca8(s_ref + j, p.ref + 2 * q);
    s_q[j] = q;
  }
  cp_async_wait_all();
  __syncthreads();        // staged query data visible to every thread; the mbarrier init too
  if (!(p.dbg & 1)) mbar_wait(&bar, 0);

  const long long vrow = (long long)p.M * 32;
  const __nv_bfloat16* vimg = p.value + (long long)n * p.S * vrow + (long long)m * 32 + ch * 8;

  for (int base = 0; base < total; base += ngrp) {     // warp-uniform trip count: the shuffles below use the full mask
    const bool valid = base + grp < total;
    const int j = min(base + grp, total - 1);
    const int q = s_q[j];
    const long long row = (long long)n * p.Lq + q;
    const float2 rf = s_ref[j];
    const float rx = rf.x, ry = rf.y;
    float lg[NK];
    float2 of[NK];
    float mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < NK; ++k) {
      const int pi = sub + 8 * k;
      lg[k] = -INFINITY;
      of[k] = make_float2(0.f, 0.f);
      if (pi < LP) {
        lg[k] = s_lg[j * LP + pi];
        of[k] = s_off[j * LP + pi];
        mx = fmaxf(mx, lg[k]);
      }
    }
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 4));
    float den = 0.f;
#pragma unroll
    for (int k = 0; k < NK; ++k) {
      lg[k] = (sub + 8 * k < LP) ? __expf(lg[k] - mx) : 0.f;
      den += lg[k];
    }
    den += __shfl_xor_sync(0xffffffffu, den, 1);
    den += __shfl_xor_sync(0xffffffffu, den, 2);
    den += __shfl_xor_sync(0xffffffffu, den, 4);
    const float inv = 1.f / den;

    // my points: bilinear factors and the byte offset of the top-left corner inside the staged box (-1: not staged)
    int soff[NK], xy[NK];
    float plw[NK], ptop[NK], pbot[NK];
#pragma unroll
    for (int k = 0; k < NK; ++k) {
      const int pi = sub + 8 * k;
      const int l = min(pi >> 2, p.L - 1);
      const int H = p.H[l], W = p.W[l];
      // same arithmetic as the reference: loc = ref + off / (W, H); im = loc * size - 0.5
      const float h_im = (ry + of[k].y / H) * H - 0.5f, w_im = (rx + of[k].x / W) * W - 0.5f;
      const bool inside = (pi < LP) && h_im > -1.f && w_im > -1.f && h_im < H && w_im < W;
      const float hf = floorf(h_im), wf = floorf(w_im);
      const int y0 = inside ? (int)hf : 0, x0 = inside ? (int)wf : 0;
      const float lh = h_im - hf;
      const float aw = inside ? lg[k] * inv : 0.f;
      plw[k] = inside ? w_im - wf : 0.f;
      ptop[k] = aw * (1.f - lh);
      pbot[k] = aw * lh;
      const int ox = (int)floorf(fx0 * W - 0.5f + p.pmin[l][m][0]) - p.margin;
      const int oy = (int)floorf(fy0 * H - 0.5f + p.pmin[l][m][1]) - p.margin;
      const int bx = x0 - ox, by = y0 - oy;
      const bool inbox = bx >= 0 && bx <= p.BW[l] - 2 && by >= 0 && by <= p.BH[l] - 2 && !(p.dbg & 4);
      soff[k] = inside ? (inbox ? p.box_off[l] + (by * p.BW[l] + bx) * 64 : -1) : 0;   // outside the map: weight 0
      xy[k] = (y0 * 65536) | (x0 & 0xffff);
    }

    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
#pragma unroll
    for (int pp = 0; pp < MAXL * 4; ++pp) {
      if (pp < LP && !(p.dbg & 2)) {
        const int k = pp >> 3, src = pp & 7, l = pp >> 2;
        const int so = __shfl_sync(0xffffffffu, soff[k], src, 8);
        const int pxy = __shfl_sync(0xffffffffu, xy[k], src, 8);
        const float lw = __shfl_sync(0xffffffffu, plw[k], src, 8);
        const float at = __shfl_sync(0xffffffffu, ptop[k], src, 8);
        const float ab = __shfl_sync(0xffffffffu, pbot[k], src, 8);
        const float wx = side ? lw : 1.f - lw;
        const float kt = at * wx, kb = ab * wx;
        uint4 ut, ub;
        if (so >= 0) {
          const uint8_t* a = st_smem + so + side * 64 + ch * 16;
          ut = *reinterpret_cast<const uint4*>(a);
          ub = *reinterpret_cast<const uint4*>(a + p.BW[l] * 64);
        } else {
          const int H = p.H[l], W = p.W[l];
          const int x = (int)(short)(pxy & 0xffff) + side, y = pxy >> 16;
          const __nv_bfloat16* vl = vimg + (long long)p.lsi[l] * vrow;
          ut = ub = make_uint4(0, 0, 0, 0);
          if (x >= 0 && x < W) {
            if (y >= 0 && y < H) ut = __ldg(reinterpret_cast<const uint4*>(vl + ((long long)y * W + x) * vrow));
            if (y + 1 >= 0 && y + 1 < H) ub = __ldg(reinterpret_cast<const uint4*>(vl + ((long long)(y + 1) * W + x) * vrow));
          }
        }
        float v[8];
        unpack8(ut, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = fmaf(kt, v[i], acc[i]);
        unpack8(ub, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = fmaf(kb, v[i], acc[i]);
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 4);
    if (valid && side == 0) *reinterpret_cast<uint4*>(p.out + (row * p.M + m) * 32 + ch * 8) = pack8(acc);
  }
}

// Second structure for the same staging (default of mmsam_msda_fused_staged_bf16; MMSAM_MSDA_STAGED_V=1 selects the
// kernel above). msda_staged_kernel is issue-bound on its skeleton (8 lanes per query at 50 % lane efficiency for
// L = 1, 5 shuffles per point, scalar FMA); here
//   * the region's queries are cut into PASSES of up to 8 consecutive queries of one grid row; the pass list
//     {first query, count} and the reference points are built once per CTA in shared memory, so the loop has no
//     integer division and no per-lane search;
//   * a warp owns a pass at a time and nothing is shared between warps after the box has landed (no CTA barrier in
//     the loop); the projection slices of the next DEPTH passes (a {2 L P, 8 rows} box of offsets and a {L P, 8 rows}
//     box of logits of the fp32 projection matrix) are in flight in a per-warp TMA ring with one mbarrier per slot:
//     no per-lane address arithmetic, no LSU wavefronts for the copies;
//   * phase A: lane = (query, point of a level): location, softmax weight (two width-4 shuffles), bilinear factors
//     and the byte offset of the top-left corner in the staged box -> one 16-byte record {offset, lw, a_top, a_bot}
//     per point in the warp's shared-memory slab. The box is zero-filled outside the map, so a staged sample needs no
//     bounds logic at all (coordinates are clamped to [-2, size + 1]: everything beyond has four invalid corners);
//   * phase B: lane = (query, 16-byte channel chunk): per point ONE broadcast LDS.128 of the record (no shuffles),
//     four LDS.128 corner reads and packed FFMA2 accumulation. The two queries of a quarter warp read the two x
//     neighbours in opposite order when needed, so that their 64-byte units always sit in opposite halves of the
//     128-byte bank line: every corner read is conflict-free (tools/micro/gather_bench.cu mode 9: 0.54 cycles per
//     unit against 0.78 for unordered units and 1.6 through L1). A pass in which some sample left its box (warp
//     vote) runs the same loop with the global-memory fall-back compiled in; the others run branch-free.
__device__ __forceinline__ uint2 lds64(uint32_t a) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
  return v;
}
__device__ __forceinline__ uint32_t lds32(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts64(uint32_t a, uint32_t x, uint32_t y) {
  asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(a), "r"(x), "r"(y) : "memory");
}

__device__ __forceinline__ float ex2_ftz(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_ftz(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

template <int MAXL, int DEPTH>
__global__ void __launch_bounds__(256, 2)
msda_staged2_kernel(const __grid_constant__ MsdaMaps2 maps, const MsdaStagedParams p) {
  extern __shared__ __align__(128) uint8_t st_smem[];
  uint64_t& bar = *reinterpret_cast<uint64_t*>(st_smem + p.bar_off);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const int ql = lane >> 2, pl = lane & 3;   // query slot of the pass; point (phase A) / channel chunk (phase B)
  const int m = blockIdx.y, n = blockIdx.z;
  const int ty = blockIdx.x / p.tiles_x, tx = blockIdx.x - ty * p.tiles_x;
  const int X0 = tx * p.TAW, X1 = min(X0 + p.TAW, p.AW), Y0 = ty * p.TAH, Y1 = min(Y0 + p.TAH, p.AH);
  const float fx0 = (float)X0 / (float)p.AW, fy0 = (float)Y0 / (float)p.AH;
  const uint32_t sbase = smem_u32(st_smem);
  const uint32_t seg_s = sbase + p.seg_off;          // [pass] {first query, count}
  const uint32_t ref_s = sbase + p.sref_off;         // [pass][8] reference point (x, y)
  const uint32_t hdr_s = sbase + p.bar_off + 16;     // [3 grids] {x0, y0, width in passes, rows} + pass count
  uint64_t* ring_bar = reinterpret_cast<uint64_t*>(st_smem + p.bar_off + 64) + (threadIdx.x >> 5) * DEPTH;   // [warp][slot]

  int ox[MAXL], oy[MAXL];
  float rW[MAXL], rH[MAXL];
#pragma unroll
  for (int l = 0; l < MAXL; ++l) {
    const int ll = l < p.L ? l : 0;
    rW[l] = 1.f / (float)p.W[ll]; rH[l] = 1.f / (float)p.H[ll];
    ox[l] = (int)floorf(fx0 * p.W[ll] - 0.5f + p.pmin[ll][m][0]) - p.margin;
    oy[l] = (int)floorf(fy0 * p.H[ll] - 0.5f + p.pmin[ll][m][1]) - p.margin;
  }
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    for (int i = 0; i < 8 * DEPTH; ++i) mbar_init(reinterpret_cast<uint64_t*>(st_smem + p.bar_off + 64) + i, 1);
    fence_barrier_init();
    mbar_arrive_expect_tx(&bar, p.tx_bytes);
#pragma unroll
    for (int l = 0; l < MAXL; ++l)
      if (l < p.L) tma_load_4d(st_smem + p.box_off[l], &maps.m[l], &bar, m * 32, ox[l], oy[l], n);
  }
  // the region's rectangle of every query grid: queries whose cell centre (i + 0.5) / g lies in [X0, X1) / AW.
  // One thread per grid does the divisions; the pass list is then built by one thread per pass.
  if (threadIdx.x < 3) {
    const int g = threadIdx.x;
    int ix0 = 0, iy0 = 0, w = 0, h = 0;
    if (g < p.NG) {
      ix0 = cdiv_pos(2 * X0 * p.gw[g] - p.AW, 2 * p.AW);
      iy0 = cdiv_pos(2 * Y0 * p.gh[g] - p.AH, 2 * p.AH);
      w = cdiv_pos(2 * X1 * p.gw[g] - p.AW, 2 * p.AW) - ix0;
      h = cdiv_pos(2 * Y1 * p.gh[g] - p.AH, 2 * p.AH) - iy0;
    }
    sts128(hdr_s + g * 16, make_uint4((uint32_t)ix0, (uint32_t)iy0, (uint32_t)w, (uint32_t)h));
  }
  __syncthreads();
  int npass = 0;
  {
    int pbase[4];
    uint4 hd[3];
    pbase[0] = 0;
#pragma unroll
    for (int g = 0; g < 3; ++g) {
      hd[g] = lds128(hdr_s + g * 16);
      pbase[g + 1] = pbase[g] + (int)hd[g].w * (((int)hd[g].z + 7) >> 3);
    }
    npass = pbase[3];
    for (int k = threadIdx.x; k < npass; k += blockDim.x) {
      const int g = k >= pbase[2] ? 2 : (k >= pbase[1] ? 1 : 0);
      const uint4 h4 = g == 2 ? hd[2] : (g == 1 ? hd[1] : hd[0]);
      const int r = k - (g == 2 ? pbase[2] : (g == 1 ? pbase[1] : 0));
      const int spr = ((int)h4.z + 7) >> 3;                     // passes per row
      const int iy = r / spr, sx = (r - iy * spr) * 8;
      const int gw = g == 2 ? p.gw[2] : (g == 1 ? p.gw[1] : p.gw[0]);
      const int gs = g == 2 ? p.gstart[2] : (g == 1 ? p.gstart[1] : p.gstart[0]);
      const int q0 = gs + ((int)h4.y + iy) * gw + (int)h4.x + sx;
      const int c = min(8, (int)h4.z - sx);
      sts64(seg_s + k * 8, (uint32_t)q0, (uint32_t)c);
      for (int s8 = 0; s8 < 8; ++s8) {
        const float2 rf = __ldg(reinterpret_cast<const float2*>(p.ref + 2 * (q0 + min(s8, c - 1))));
        sts64(ref_s + (k * 8 + s8) * 8, __float_as_uint(rf.x), __float_as_uint(rf.y));
      }
    }
  }
  __syncthreads();        // pass list, reference points, mbarrier init visible to every thread

  const uint32_t rec_w = sbase + p.rec_off + (uint32_t)warp * (MAXL * 4 * 8 * 16);   // [point][query slot ^ 2 point] 16 B
  const uint32_t xy_w = sbase + p.xy_off + (uint32_t)warp * (MAXL * 4 * 8 * 4);
  const int LP = p.L * 4;
  const long long vrow = (long long)p.M * 32;
  const __nv_bfloat16* vimg = p.value + (long long)n * p.S * vrow + (long long)m * 32 + pl * 8;
  const uint32_t orow32 = (uint32_t)p.M * 32u;    // in-image offsets fit 32 bits (host check)
  __nv_bfloat16* olane = p.out + (long long)n * p.Lq * vrow + m * 32 + pl * 8;

  // input ring: slot = [8 queries][L P float2 offsets] | [8 queries][L P logits], one TMA box each
  const int STAGE = LP * 96;
  const uint32_t in_w = sbase + p.in_off + (uint32_t)(warp * DEPTH * STAGE);
  auto issue_pass = [&](int k, int slot) {
    if (k < npass) {                        // warp-uniform
      if (elect_one()) {
        const int row = n * p.Lq + (int)lds32(seg_s + k * 8);
        uint8_t* st = st_smem + p.in_off + (warp * DEPTH + slot) * STAGE;
        mbar_arrive_expect_tx(&ring_bar[slot], (uint32_t)STAGE);
        tma_load_2d(st, &maps.qoff, &ring_bar[slot], m * LP * 2, row);
        tma_load_2d(st + LP * 64, &maps.qlg, &ring_bar[slot], p.M * LP * 2 + m * LP, row);
      }
      __syncwarp();
    }
  };
#pragma unroll
  for (int d = 0; d < DEPTH; ++d) issue_pass(warp + d * nwarp, d);
  mbar_wait(&bar, 0);

  int slot = 0;
  uint32_t ring_ph = 0;
  for (int k = warp; k < npass; k += nwarp) {     // warp-uniform
    const uint2 sg = lds64(seg_s + k * 8);
    const bool valid = ql < (int)sg.y;
    const int q = (int)sg.x + min(ql, (int)sg.y - 1);
    const uint2 rfu = lds64(ref_s + (k * 8 + ql) * 8);
    const float rx = __uint_as_float(rfu.x), ry = __uint_as_float(rfu.y);
    mbar_wait(&ring_bar[slot], ring_ph);
    float2 of_c[MAXL];
    float lg_c[MAXL];
    {
      const uint32_t st = in_w + (uint32_t)(slot * STAGE);
#pragma unroll
      for (int l = 0; l < MAXL; ++l) {
        lg_c[l] = -INFINITY;
        of_c[l] = make_float2(0.f, 0.f);
        if (l < p.L) {
          const uint2 o = lds64(st + (uint32_t)(ql * (LP * 8) + l * 32 + pl * 8));
          of_c[l] = make_float2(__uint_as_float(o.x), __uint_as_float(o.y));
          lg_c[l] = __uint_as_float(lds32(st + (uint32_t)(LP * 64 + ql * (LP * 4) + l * 16 + pl * 4)));
        }
      }
    }
    const int slot_read = slot;
    if (++slot == DEPTH) { slot = 0; ring_ph ^= 1u; }

    // ---- phase A: lane = (query ql, point pl of every level) ----
    float e[MAXL];
    float mx = -INFINITY;
#pragma unroll
    for (int l = 0; l < MAXL; ++l) mx = fmaxf(mx, lg_c[l]);
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
    float den = 0.f;
#pragma unroll
    for (int l = 0; l < MAXL; ++l) {
      e[l] = l < p.L ? ex2_ftz((lg_c[l] - mx) * 1.4426950408889634f) : 0.f;
      den += e[l];
    }
    den += __shfl_xor_sync(0xffffffffu, den, 1);
    den += __shfl_xor_sync(0xffffffffu, den, 2);
    const float inv = rcp_ftz(den);              // den in [1, 4 L]
    bool spill = false;                          // some sample of this lane is not in its box
#pragma unroll
    for (int l = 0; l < MAXL; ++l) {
      if (l < p.L) {
        const float Hf = (float)p.H[l], Wf = (float)p.W[l];
        // the reference's loc = ref + off / (W, H); im = loc * size - 0.5, with the division as a multiplication by
        // the reciprocal (exact for power-of-two maps, else an ulp of the location: the interpolation is continuous).
        // Clamped to [-2, size + 1]: beyond, all four corners are outside the map (the reference skips the sample),
        // and a NaN location becomes -2 (skipped by the reference's comparisons as well).
        const float h_im = fminf(fmaxf((ry + of_c[l].y * rH[l]) * Hf - 0.5f, -2.f), Hf + 1.f);
        const float w_im = fminf(fmaxf((rx + of_c[l].x * rW[l]) * Wf - 0.5f, -2.f), Wf + 1.f);
        const float hf = floorf(h_im), wf = floorf(w_im);
        const int y0 = (int)hf, x0 = (int)wf;
        const float lh = h_im - hf;
        const float aw = e[l] * inv;
        const int bx = x0 - ox[l], by = y0 - oy[l];
        const bool inbox = (unsigned)bx <= (unsigned)(p.BW[l] - 2) && (unsigned)by <= (unsigned)(p.BH[l] - 2) && !(p.dbg & 4);
        uint4 rec;
        // staged: byte offset of the top-left corner; not staged: -1 - level, the corners come from global memory
        rec.x = (uint32_t)(inbox ? p.box_off[l] + (by * p.BW[l] + bx) * 64 : -1 - l);
        rec.y = __float_as_uint(w_im - wf);
        rec.z = __float_as_uint(aw * (1.f - lh));
        rec.w = __float_as_uint(aw * lh);
        sts128(rec_w + (uint32_t)(((l * 4 + pl) * 8 + (ql ^ (2 * pl))) * 16), rec);
        if (!inbox) {
          sts32(xy_w + (uint32_t)(((l * 4 + pl) * 8 + ql) * 4), (uint32_t)((y0 * 65536) | (x0 & 0xffff)));
          spill = true;
        }
      }
    }
    const bool any_spill = __any_sync(0xffffffffu, spill);
    __syncwarp();          // the records are visible to the whole warp; every lane has consumed its ring slot
    issue_pass(k + DEPTH * nwarp, slot_read);      // refill the slot just read

    // ---- phase B: lane = (query ql, channel chunk pl) ----
    u64 acc2[4] = {0ull, 0ull, 0ull, 0ull};
    auto gather = [&](auto with_fallback) {
#pragma unroll
      for (int l = 0; l < MAXL; ++l) {
        if (l < p.L && !(p.dbg & 2)) {
          const uint32_t rowb = (uint32_t)p.BW[l] * 64u;
#pragma unroll
          for (int pp = 0; pp < 4; ++pp) {
            const uint4 rec = lds128(rec_w + (uint32_t)(((l * 4 + pp) * 8 + (ql ^ (2 * pp))) * 16));
            const int so = (int)rec.x;
            const float lw = __uint_as_float(rec.y), at = __uint_as_float(rec.z), ab = __uint_as_float(rec.w);
            uint4 u0, u1, u2, u3;
            float w0;        // x weight of (u0, u2); (u1, u3) carry 1 - w0
            if (!decltype(with_fallback)::value || so >= 0) {
              // d = 1: this lane reads the right neighbour first, so that the pair (ql even, ql odd) of a quarter
              // warp always hits opposite halves of the 128-byte bank line
              const uint32_t d = (((uint32_t)so >> 6) ^ (uint32_t)ql) & 1u;
              const uint32_t a0 = sbase + (uint32_t)so + (uint32_t)pl * 16u;
              const uint32_t af = a0 + d * 64u, as = a0 + 64u - d * 64u;
              u0 = lds128(af); u1 = lds128(as); u2 = lds128(af + rowb); u3 = lds128(as + rowb);
              w0 = d ? lw : 1.f - lw;
            } else {
              const int lv = -1 - so;
              const int pxy = (int)lds32(xy_w + (uint32_t)(((l * 4 + pp) * 8 + ql) * 4));
              const int H = p.H[lv], W = p.W[lv];
              const int x = (int)(short)(pxy & 0xffff), y = pxy >> 16;
              const __nv_bfloat16* vl = vimg + (long long)p.lsi[lv] * vrow;
              u0 = u1 = u2 = u3 = make_uint4(0, 0, 0, 0);
              const bool yt = y >= 0 && y < H, yb = y + 1 >= 0 && y + 1 < H;
              if (x >= 0 && x < W) {
                if (yt) u0 = __ldg(reinterpret_cast<const uint4*>(vl + ((long long)y * W + x) * vrow));
                if (yb) u2 = __ldg(reinterpret_cast<const uint4*>(vl + ((long long)(y + 1) * W + x) * vrow));
              }
              if (x + 1 >= 0 && x + 1 < W) {
                if (yt) u1 = __ldg(reinterpret_cast<const uint4*>(vl + ((long long)y * W + x + 1) * vrow));
                if (yb) u3 = __ldg(reinterpret_cast<const uint4*>(vl + ((long long)(y + 1) * W + x + 1) * vrow));
              }
              w0 = 1.f - lw;
            }
            const float w1 = 1.f - w0;
            const float k0 = at * w0, k1 = at * w1, k2 = ab * w0, k3 = ab * w1;
            const u64 kk0 = pack2(k0, k0), kk1 = pack2(k1, k1), kk2 = pack2(k2, k2), kk3 = pack2(k3, k3);
            const uint32_t* c0 = reinterpret_cast<const uint32_t*>(&u0);
            const uint32_t* c1 = reinterpret_cast<const uint32_t*>(&u1);
            const uint32_t* c2 = reinterpret_cast<const uint32_t*>(&u2);
            const uint32_t* c3 = reinterpret_cast<const uint32_t*>(&u3);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              acc2[i] = fma2(pack2(bf16lo(c0[i]), bf16hi(c0[i])), kk0, acc2[i]);
              acc2[i] = fma2(pack2(bf16lo(c1[i]), bf16hi(c1[i])), kk1, acc2[i]);
              acc2[i] = fma2(pack2(bf16lo(c2[i]), bf16hi(c2[i])), kk2, acc2[i]);
              acc2[i] = fma2(pack2(bf16lo(c3[i]), bf16hi(c3[i])), kk3, acc2[i]);
            }
          }
        }
      }
    };
    if (any_spill) gather(std::true_type{});
    else gather(std::false_type{});
    float acc[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) unpack2(acc2[i], acc[2 * i], acc[2 * i + 1]);
    if (valid) *reinterpret_cast<uint4*>(olane + (uint32_t)q * orow32) = pack8(acc);
    __syncwarp();          // the records are rewritten by the next pass
  }
}

template <typename VT, typename AT>
static int launch_msda(const void* value, const int64_t* shapes, const int64_t* lsi, const void* loc,
                       const void* attw, void* out, int N, int S, int M, int D, int Lq, int L, int P,
                       cudaStream_t st) {
  const VT* v = (const VT*)value;
  const AT* lc = (const AT*)loc;
  const AT* aw = (const AT*)attw;
  VT* o = (VT*)out;
  if (N == 0 || Lq == 0 || M == 0 || D == 0) return MMSAM_OK;
  if constexpr (!std::is_same<VT, double>::value) {
    constexpr int VEC = Vec16<VT>::N;
    const bool aligned = (((uintptr_t)value | (uintptr_t)out) & 15) == 0;
    if (aligned && D % VEC == 0 && D / VEC <= 32 && (P == 4 || P == 8 || P == 2 || P == 1) && N <= 65535 && M <= 65535) {
      const int TPP = D / VEC;
      int HB = 1;
      for (int h = 2; h >= 1; --h)
        if (M % h == 0) { HB = h; break; }
      int QB = 256 / (HB * TPP);
      if (QB > 32) QB = 32;
      if (QB < 1) QB = 1;
      dim3 grid((Lq + QB - 1) / QB, M / HB, N), block(QB * HB * TPP);
      switch (P) {
        case 1: msda_vec_kernel<VT, AT, 1><<<grid, block, 0, st>>>(v, shapes, lsi, lc, aw, o, S, M, D, Lq, L, HB, TPP); break;
        case 2: msda_vec_kernel<VT, AT, 2><<<grid, block, 0, st>>>(v, shapes, lsi, lc, aw, o, S, M, D, Lq, L, HB, TPP); break;
        case 4: msda_vec_kernel<VT, AT, 4><<<grid, block, 0, st>>>(v, shapes, lsi, lc, aw, o, S, M, D, Lq, L, HB, TPP); break;
        default: msda_vec_kernel<VT, AT, 8><<<grid, block, 0, st>>>(v, shapes, lsi, lc, aw, o, S, M, D, Lq, L, HB, TPP); break;
      }
      MMSAM_LAUNCH_CHECK();
      return MMSAM_OK;
    }
  }
  const long long total = (long long)N * Lq * M * D;
  long long blocks = (total + 255) / 256;
  if (blocks > (long long)kNumSMs * 32) blocks = (long long)kNumSMs * 32;
  msda_generic_kernel<VT, AT><<<(unsigned)blocks, 256, 0, st>>>(v, shapes, lsi, lc, aw, o, N, S, M, D, Lq, L, P);
  MMSAM_LAUNCH_CHECK();
  return MMSAM_OK;
}

}  // namespace mmsam

// See include/mmsam_b200.h for the contract.
MMSAM_API int mmsam_msda_forward(const void* value, const int64_t* spatial_shapes_dev,
                                 const int64_t* level_start_index_dev, const void* sampling_loc,
                                 const void* attn_weight, void* out, int N, int S, int M, int D,
                                 int Lq, int L, int P, int value_dtype, int aux_dtype,
                                 void* stream) {
  using namespace mmsam;
  if (N < 0 || S < 0 || M < 0 || D < 0 || Lq < 0 || L < 0 || P < 0) return MMSAM_ERR_BAD_ARG;
  if (N > 0 && Lq > 0 && M > 0 && D > 0 &&
      (!value || !spatial_shapes_dev || !level_start_index_dev || !sampling_loc || !attn_weight || !out))
    return MMSAM_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
#define MSDA_CASE(VD, AD, VT, AT) \
  if (value_dtype == VD && aux_dtype == AD) \
    return launch_msda<VT, AT>(value, spatial_shapes_dev, level_start_index_dev, sampling_loc, attn_weight, out, N, S, M, D, Lq, L, P, st);
  MSDA_CASE(MMSAM_BF16, MMSAM_F32, __nv_bfloat16, float)
  MSDA_CASE(MMSAM_BF16, MMSAM_BF16, __nv_bfloat16, __nv_bfloat16)
  MSDA_CASE(MMSAM_F16, MMSAM_F32, __half, float)
  MSDA_CASE(MMSAM_F16, MMSAM_F16, __half, __half)
  MSDA_CASE(MMSAM_F32, MMSAM_F32, float, float)
  MSDA_CASE(MMSAM_F64, MMSAM_F64, double, double)
#undef MSDA_CASE
  return MMSAM_ERR_BAD_DTYPE;
}

// Fused variant (bf16 value/out): see msda_fused_kernel. qproj fp32 [N*Lq, ldq], ref fp32 [Lq,2] (x,y).
MMSAM_API int mmsam_msda_fused_bf16(const void* value, const int64_t* spatial_shapes_dev,
                                    const int64_t* level_start_index_dev, const float* qproj, long long ldq,
                                    const float* ref_xy, void* out, int N, int S, int M, int D, int Lq, int L,
                                    int P, void* stream) {
  using namespace mmsam;
  if (N < 0 || S < 0 || M <= 0 || D <= 0 || Lq < 0 || L <= 0) return MMSAM_ERR_BAD_ARG;
  if (N == 0 || Lq == 0) return MMSAM_OK;
  if (!value || !spatial_shapes_dev || !level_start_index_dev || !qproj || !ref_xy || !out) return MMSAM_ERR_BAD_ARG;
  if ((D & 7) || D / 8 > 32 || P != 4 || (((uintptr_t)value | (uintptr_t)out) & 15) || (((uintptr_t)qproj) & 7) || (ldq & 1))
    return MMSAM_ERR_UNSUPPORTED;
  if (D == 32 && L <= 4 && !getenv("MMSAM_MSDA_LEGACY")) {
    dim3 grid((Lq + 7) / 8, (M + 7) / 8, N);
    if (L <= 1)
      msda_fused_coop_kernel<1><<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)value, spatial_shapes_dev,
          level_start_index_dev, qproj, ldq, ref_xy, (__nv_bfloat16*)out, S, M, Lq, L);
    else
      msda_fused_coop_kernel<4><<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)value, spatial_shapes_dev,
          level_start_index_dev, qproj, ldq, ref_xy, (__nv_bfloat16*)out, S, M, Lq, L);
    MMSAM_LAUNCH_CHECK();
    return MMSAM_OK;
  }
  const int TPP = D / 8;
  int HB = (M % 2 == 0) ? 2 : 1;
  int QB = 256 / (HB * TPP);
  if (QB > 32) QB = 32;
  if (QB < 1) QB = 1;
  dim3 grid((Lq + QB - 1) / QB, M / HB, N), block(QB * HB * TPP);
  msda_fused_kernel<__nv_bfloat16, 4><<<grid, block, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)value, spatial_shapes_dev, level_start_index_dev, qproj, ldq, ref_xy,
      (__nv_bfloat16*)out, S, M, D, Lq, L, HB, TPP);
  MMSAM_LAUNCH_CHECK();
  return MMSAM_OK;
}

// Staged variant of mmsam_msda_fused_bf16 (see msda_staged_kernel and include/mmsam_b200.h). All *_host arrays are
// HOST pointers read before the call returns.
MMSAM_API int mmsam_msda_fused_staged_bf16(const void* value, const int* level_hw_host, const float* qproj,
                                           long long ldq, const float* ref_xy, void* out, int N, int S, int M, int D,
                                           int Lq, int L, int P, int n_qgrids, const int* qgrid_hw_host, int anchor_h,
                                           int anchor_w, int tile_h, int tile_w, const float* prior_min_host,
                                           const float* prior_ext_host, int margin, void* stream) {
  using namespace mmsam;
  if (N < 0 || S < 0 || M <= 0 || D <= 0 || Lq < 0 || L <= 0 || margin < 0) return MMSAM_ERR_BAD_ARG;
  if (N == 0 || Lq == 0) return MMSAM_OK;
  if (!value || !level_hw_host || !qproj || !ref_xy || !out || !qgrid_hw_host || !prior_min_host || !prior_ext_host)
    return MMSAM_ERR_BAD_ARG;
  if (D != 32 || P != 4 || L > 4 || M > 16 || n_qgrids < 1 || n_qgrids > 3 || N > 65535 || anchor_h <= 0 ||
      anchor_w <= 0 || tile_h <= 0 || tile_w <= 0 || (((uintptr_t)value | (uintptr_t)out) & 15) ||
      (((uintptr_t)qproj) & 15) || (ldq & 3) || (((uintptr_t)ref_xy) & 7))
    return MMSAM_ERR_UNSUPPORTED;
  MsdaStagedParams p;
  memset(&p, 0, sizeof(p));
  p.value = (const __nv_bfloat16*)value; p.qproj = qproj; p.ref = ref_xy; p.out = (__nv_bfloat16*)out;
  p.ldq = ldq; p.S = S; p.M = M; p.Lq = Lq; p.L = L; p.margin = margin;
  p.AH = anchor_h; p.AW = anchor_w; p.TAH = tile_h; p.TAW = tile_w; p.NG = n_qgrids;
  long long s_sum = 0, q_sum = 0;
  unsigned off = 0;
  for (int l = 0; l < L; ++l) {
    p.H[l] = level_hw_host[2 * l]; p.W[l] = level_hw_host[2 * l + 1];
    if (p.H[l] <= 0 || p.W[l] <= 0 || p.H[l] > 32767 || p.W[l] > 32767) return MMSAM_ERR_BAD_ARG;
    p.lsi[l] = (int)s_sum;
    s_sum += (long long)p.H[l] * p.W[l];
    const int span_x = (tile_w * p.W[l] + anchor_w - 1) / anchor_w, span_y = (tile_h * p.H[l] + anchor_h - 1) / anchor_h;
    p.BW[l] = span_x + (int)ceilf(prior_ext_host[2 * l]) + 2 * margin + 2;
    p.BH[l] = span_y + (int)ceilf(prior_ext_host[2 * l + 1]) + 2 * margin + 2;
    if (p.BW[l] > 256 || p.BH[l] > 256) return MMSAM_ERR_UNSUPPORTED;
    p.box_off[l] = (int)off;
    off += (unsigned)(p.BW[l] * p.BH[l] * 64);
    off = (off + 127u) & ~127u;
    for (int m = 0; m < M; ++m) {
      p.pmin[l][m][0] = prior_min_host[(l * M + m) * 2];
      p.pmin[l][m][1] = prior_min_host[(l * M + m) * 2 + 1];
    }
  }
  p.tx_bytes = 0;
  for (int l = 0; l < L; ++l) p.tx_bytes += (unsigned)(p.BW[l] * p.BH[l] * 64);
  for (int g = 0; g < n_qgrids; ++g) {
    p.gh[g] = qgrid_hw_host[2 * g]; p.gw[g] = qgrid_hw_host[2 * g + 1];
    if (p.gh[g] <= 0 || p.gw[g] <= 0 || p.gh[g] > 16384 || p.gw[g] > 16384) return MMSAM_ERR_BAD_ARG;
    p.gstart[g] = (int)q_sum;
    q_sum += (long long)p.gh[g] * p.gw[g];
  }
  if (s_sum != S || q_sum != Lq) return MMSAM_ERR_BAD_ARG;
  // the most queries a region serves: same cell-centre assignment as the kernel, per axis
  auto max_cells = [](int anchor, int tile, int g) {
    int best = 0;
    for (int a0 = 0; a0 < anchor; a0 += tile) {
      const int a1 = a0 + tile < anchor ? a0 + tile : anchor;
      const int i0 = (2 * a0 * g - anchor + 2 * anchor - 1) / (2 * anchor), i1 = (2 * a1 * g - anchor + 2 * anchor - 1) / (2 * anchor);
      if (i1 - i0 > best) best = i1 - i0;
    }
    return best;
  };
  unsigned qmax = 0;
  for (int g = 0; g < n_qgrids; ++g)
    qmax += (unsigned)(max_cells(anchor_w, tile_w, p.gw[g]) * max_cells(anchor_h, tile_h, p.gh[g]));
  const unsigned LP = (unsigned)L * 4;
  p.qoff_off = off; off += qmax * LP * 8;
  p.qlg_off = off; off += qmax * LP * 4;
  p.qref_off = off; off += qmax * 8;
  p.qidx_off = off; off += qmax * 4;
  off = (off + 15u) & ~15u;
  static const int version = [] { const char* e = getenv("MMSAM_MSDA_STAGED_V"); return e ? atoi(e) : 2; }();
  if (version != 1 && ((long long)N * Lq >= (1ll << 31) || (long long)Lq * M * 32 >= (1ll << 32) || ldq < (long long)M * LP * 3)) return MMSAM_ERR_UNSUPPORTED;
  if (version != 1) {      // msda_staged2_kernel keeps the query inputs in registers: the areas above are not used
    off = p.qoff_off;
    const unsigned maxl = L <= 1 ? 1u : 4u;
    p.rec_off = off; off += 8u * maxl * 4u * 8u * 16u;      // 8 warps x points x 8 query slots x 16 B
    p.xy_off = off; off += 8u * maxl * 4u * 8u * 4u;
    off = (off + 127u) & ~127u;
    p.in_off = off; off += 8u * (L <= 1 ? 4u : 2u) * (LP * 96u);   // 8 warps x DEPTH x STAGE (TMA destinations: 128-byte multiples)
    unsigned pmax = 0;     // passes (<= 8 consecutive queries of a grid row) of the largest region
    for (int g = 0; g < n_qgrids; ++g)
      pmax += (unsigned)(((max_cells(anchor_w, tile_w, p.gw[g]) + 7) / 8) * max_cells(anchor_h, tile_h, p.gh[g]));
    p.seg_off = off; off += pmax * 8u;
    p.sref_off = off; off += pmax * 64u;
    off = (off + 15u) & ~15u;
  }
  p.bar_off = off;
  { const char* e = getenv("MMSAM_MSDA_DBG"); p.dbg = e ? atoi(e) : 0; }
  const unsigned smem = off + 64 + 8 * 4 * 8;      // mbarrier + the rectangle header + the ring barriers of msda_staged2_kernel
  if (smem > 226u * 1024u) return MMSAM_ERR_UNSUPPORTED;
  p.tiles_x = (anchor_w + tile_w - 1) / tile_w;
  const int tiles_y = (anchor_h + tile_h - 1) / tile_h;

  MsdaMaps maps;
  mmsam_host::EncodeTiledFn enc = mmsam_host::get_encode_tiled();
  if (!enc) return MMSAM_ERR_DRIVER;
  for (int l = 0; l < 4; ++l) {
    const int ll = l < L ? l : 0;
    cuuint64_t dims[4] = {(cuuint64_t)M * 32, (cuuint64_t)p.W[ll], (cuuint64_t)p.H[ll], (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)M * 64, (cuuint64_t)p.W[ll] * M * 64, (cuuint64_t)S * M * 64};
    cuuint32_t box[4] = {32, (cuuint32_t)p.BW[ll], (cuuint32_t)p.BH[ll], 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    void* base = (void*)((const __nv_bfloat16*)value + (long long)p.lsi[ll] * M * 32);
    if (enc(&maps.m[l], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return MMSAM_ERR_DRIVER;
  }
  dim3 grid((unsigned)(p.tiles_x * tiles_y), (unsigned)M, (unsigned)N);
  cudaStream_t st = (cudaStream_t)stream;
  MMSAM_SET_SMEM_ONCE(msda_staged_kernel<2>, 226 * 1024);
  MMSAM_SET_SMEM_ONCE(msda_staged_kernel<4>, 226 * 1024);
  if (version != 1) {
    MsdaMaps2 maps2;
    for (int l = 0; l < 4; ++l) maps2.m[l] = maps.m[l];
    {
      cuuint64_t dims[2] = {(cuuint64_t)ldq, (cuuint64_t)N * Lq};
      cuuint64_t strides[1] = {(cuuint64_t)ldq * 4};
      cuuint32_t estr[2] = {1, 1};
      cuuint32_t box_o[2] = {LP * 2, 8}, box_l[2] = {LP, 8};
      if (enc(&maps2.qoff, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)qproj, dims, strides, box_o, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS ||
          enc(&maps2.qlg, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)qproj, dims, strides, box_l, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return MMSAM_ERR_DRIVER;
    }
    MMSAM_SET_SMEM_ONCE((msda_staged2_kernel<1, 4>), 226 * 1024);
    MMSAM_SET_SMEM_ONCE((msda_staged2_kernel<4, 2>), 226 * 1024);
    if (L <= 1) msda_staged2_kernel<1, 4><<<grid, 256, smem, st>>>(maps2, p);
    else msda_staged2_kernel<4, 2><<<grid, 256, smem, st>>>(maps2, p);
    MMSAM_LAUNCH_CHECK();
    return MMSAM_OK;
  }
  // 8 lanes per query: one pass over the region's queries when there are few (injector), 256 threads otherwise
  const unsigned threads = qmax <= 64 ? 512 : 256;
  if (L <= 2) msda_staged_kernel<2><<<grid, threads, smem, st>>>(maps, p);
  else msda_staged_kernel<4><<<grid, threads, smem, st>>>(maps, p);
  MMSAM_LAUNCH_CHECK();
  return MMSAM_OK;
}
