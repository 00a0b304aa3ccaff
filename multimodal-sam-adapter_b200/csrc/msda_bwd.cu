// Multi-scale deformable attention sampling, BACKWARD (the training half of the reference's only native op).
//
// Replaces ms_deformable_col2im_cuda / ms_deform_attn_cuda_backward of the reference extension
// (segmentation/ops/src/cuda/ms_deform_im2col_cuda.cuh:301-921, ms_deform_attn_cuda.cu:83-153; bound as
// MSDA.ms_deform_attn_backward, segmentation/ops/src/vision.cpp:13-16, and called from
// MSDeformAttnFunction.backward, segmentation/ops/functions/ms_deform_attn_func.py:39-50).
//
// With  out[n,q,m,:] = sum_{l,p} a * B(value_l[n,:,m,:], loc)   (a = attention weight, B = bilinear sample with the
// reference's conventions: w_im = x * W - 0.5, h_im = y * H - 0.5, zero padding, sample skipped unless
// -1 < h_im < H and -1 < w_im < W) and g = grad_output[n,q,m,:]:
//   grad_value[corner]      += a * w_corner * g                        (atomic: many samples hit one pixel)
//   grad_attn_weight         = <g, B>
//   grad_sampling_loc.x      = W * a * <g, dB/dw_im>,   dB/dw_im = hh (v01 - v00) + lh (v11 - v10)
//   grad_sampling_loc.y      = H * a * <g, dB/dh_im>,   dB/dh_im = hw (v10 - v00) + lw (v11 - v01)
// (v of a corner outside the map = 0). The reference runs one thread per (n, q, m, l, p, channel) and reduces the three
// scalars over the channels through shared memory, in one of seven kernel variants chosen by the channel count. Here a
// WARP owns a (n, q, m) triple: lane = channel (strided for D > 32), the L * P samples are walked in order, the three
// scalars are reduced with shuffles, lane 0 stores them (each (n, q, m, l, p) has exactly one owner: plain stores, no
// atomics, deterministic); only grad_value uses atomics, like the reference. Works for any D / L / P, fp32 and fp64
// (gradcheck). The inference path never calls this; it exists so that the drop-in extension module is complete.
#include "common.cuh"

namespace mmsam {

template <typename T>
__global__ void __launch_bounds__(256)
msda_backward_kernel(const T* __restrict__ value, const int64_t* __restrict__ shapes, const int64_t* __restrict__ lsi,
                     const T* __restrict__ loc, const T* __restrict__ attw, const T* __restrict__ gout,
                     T* __restrict__ gvalue, T* __restrict__ gloc, T* __restrict__ gattw, long long items, int S, int M,
                     int D, int Lq, int L, int P) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long item = warp0; item < items; item += nwarps) {       // item = (n, q, m)
    const int m = (int)(item % M);
    const long long nq = item / M;
    const int n = (int)(nq / Lq);
    const T* g = gout + item * D;
    const T* vb = value + ((long long)n * S * M + m) * D;
    T* gvb = gvalue + ((long long)n * S * M + m) * D;
    const long long rs = (long long)M * D;                            // elements between two pixels of a head
    for (int l = 0; l < L; ++l) {
      const int H = (int)shapes[2 * l], W = (int)shapes[2 * l + 1];
      const long long base = lsi[l] * rs;
      for (int p = 0; p < P; ++p) {
        const long long si = (item * L + l) * P + p;
        const T x = loc[2 * si], y = loc[2 * si + 1], a = attw[si];
        const T h_im = y * H - (T)0.5, w_im = x * W - (T)0.5;
        T ga = 0, gx = 0, gy = 0;
        if (h_im > (T)-1 && w_im > (T)-1 && h_im < (T)H && w_im < (T)W) {          // warp-uniform
          const T hf = floor(h_im), wf = floor(w_im);
          const int h0 = (int)hf, w0 = (int)wf;
          const T lh = h_im - hf, lw = w_im - wf, hh = (T)1 - lh, hw = (T)1 - lw;
          const bool t_ok = h0 >= 0, b_ok = h0 + 1 <= H - 1, l_ok = w0 >= 0, r_ok = w0 + 1 <= W - 1;
          const long long o00 = base + ((long long)h0 * W + w0) * rs, o01 = o00 + rs, o10 = o00 + (long long)W * rs, o11 = o10 + rs;
          for (int c = lane; c < D; c += 32) {
            const T gc = g[c];
            const T v00 = (t_ok && l_ok) ? vb[o00 + c] : (T)0, v01 = (t_ok && r_ok) ? vb[o01 + c] : (T)0;
            const T v10 = (b_ok && l_ok) ? vb[o10 + c] : (T)0, v11 = (b_ok && r_ok) ? vb[o11 + c] : (T)0;
            const T ag = a * gc;
            if (t_ok && l_ok) atomicAdd(gvb + o00 + c, hh * hw * ag);
            if (t_ok && r_ok) atomicAdd(gvb + o01 + c, hh * lw * ag);
            if (b_ok && l_ok) atomicAdd(gvb + o10 + c, lh * hw * ag);
            if (b_ok && r_ok) atomicAdd(gvb + o11 + c, lh * lw * ag);
            ga += gc * (hh * hw * v00 + hh * lw * v01 + lh * hw * v10 + lh * lw * v11);
            gx += ag * (hh * (v01 - v00) + lh * (v11 - v10));
            gy += ag * (hw * (v10 - v00) + lw * (v11 - v01));
          }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          ga += __shfl_xor_sync(0xffffffffu, ga, o);
          gx += __shfl_xor_sync(0xffffffffu, gx, o);
          gy += __shfl_xor_sync(0xffffffffu, gy, o);
        }
        if (lane == 0) {
          gattw[si] = ga;
          gloc[2 * si] = gx * (T)W;
          gloc[2 * si + 1] = gy * (T)H;
        }
      }
    }
  }
}

template <typename T>
static int launch_msda_bwd(const void* value, const int64_t* shapes, const int64_t* lsi, const void* loc, const void* attw,
                           const void* gout, void* gvalue, void* gloc, void* gattw, int N, int S, int M, int D, int Lq,
                           int L, int P, cudaStream_t st) {
  if (cudaMemsetAsync(gvalue, 0, (size_t)N * S * M * D * sizeof(T), st) != cudaSuccess) return MMSAM_ERR_DRIVER;
  const long long items = (long long)N * Lq * M;
  if (items == 0 || D == 0) return MMSAM_OK;
  long long blocks = (items + 7) / 8;                  // 8 warps per CTA
  if (blocks > (long long)kNumSMs * 16) blocks = (long long)kNumSMs * 16;
  msda_backward_kernel<T><<<(unsigned)blocks, 256, 0, st>>>((const T*)value, shapes, lsi, (const T*)loc, (const T*)attw,
                                                            (const T*)gout, (T*)gvalue, (T*)gloc, (T*)gattw, items, S, M, D,
                                                            Lq, L, P);
  MMSAM_LAUNCH_CHECK();
  return MMSAM_OK;
}

}  // namespace mmsam

// See include/mmsam_b200.h for the contract.
MMSAM_API int mmsam_msda_backward(const void* value, const int64_t* spatial_shapes_dev, const int64_t* level_start_index_dev,
                                  const void* sampling_loc, const void* attn_weight, const void* grad_output,
                                  void* grad_value, void* grad_sampling_loc, void* grad_attn_weight, int N, int S, int M,
                                  int D, int Lq, int L, int P, int dtype, void* stream) {
  using namespace mmsam;
  if (N < 0 || S < 0 || M < 0 || D < 0 || Lq < 0 || L < 0 || P < 0) return MMSAM_ERR_BAD_ARG;
  if (dtype != MMSAM_F32 && dtype != MMSAM_F64) return MMSAM_ERR_BAD_DTYPE;
  const bool any = N > 0 && M > 0 && D > 0;
  if (any && S > 0 && (!value || !grad_value)) return MMSAM_ERR_BAD_ARG;
  if (any && Lq > 0 && L > 0 && P > 0 &&
      (!spatial_shapes_dev || !level_start_index_dev || !sampling_loc || !attn_weight || !grad_output || !grad_sampling_loc ||
       !grad_attn_weight))
    return MMSAM_ERR_BAD_ARG;
  if (!any || S == 0) return MMSAM_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == MMSAM_F32)
    return launch_msda_bwd<float>(value, spatial_shapes_dev, level_start_index_dev, sampling_loc, attn_weight, grad_output, grad_value,
                                  grad_sampling_loc, grad_attn_weight, N, S, M, D, Lq, L, P, st);
  return launch_msda_bwd<double>(value, spatial_shapes_dev, level_start_index_dev, sampling_loc, attn_weight, grad_output, grad_value,
                                 grad_sampling_loc, grad_attn_weight, N, S, M, D, Lq, L, P, st);
}
