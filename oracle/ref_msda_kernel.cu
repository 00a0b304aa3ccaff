// TEST / BENCH INFRASTRUCTURE, not product code.
// Compiles the REFERENCE's own CUDA forward kernel, unmodified, for sm_100a so that bench.py can time it on the same B200
// next to mmsam_msda_fused_bf16 ("reference_kernel_gbs"): ms_deformable_im2col_cuda<scalar_t> /
// ms_deformable_im2col_gpu_kernel (segmentation/ops/src/cuda/ms_deform_im2col_cuda.cuh:237-299, 923-954) is included
// from where it lies under /root/reference (oracle/Makefile passes the -I); nothing of it is copied here. The host
// dispatcher of the reference (ms_deform_attn_cuda.cu:20-80) is ATen code that no longer compiles with torch 2.11
// (value.type()), so this file calls the launcher template directly with plain pointers — the same launch the
// dispatcher makes per im2col_step slice (:61-75), here with the whole batch as one slice.
// Output: oracle/_ref/libref_msda.so (git-ignored, travels to the GPU box).
#include <cuda_fp16.h>
#include "ms_deform_im2col_cuda.cuh"

extern "C" __attribute__((visibility("default")))
int ref_msda_im2col(const void* value, const int64_t* shapes, const int64_t* lsi, const void* loc, const void* attn,
                    void* out, int N, int S, int M, int D, int Lq, int L, int P, int is_half, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (is_half)
    ms_deformable_im2col_cuda<at::Half>(st, (const at::Half*)value, shapes, lsi, (const at::Half*)loc, (const at::Half*)attn,
                                        N, S, M, D, L, Lq, P, (at::Half*)out);
  else
    ms_deformable_im2col_cuda<float>(st, (const float*)value, shapes, lsi, (const float*)loc, (const float*)attn, N, S, M,
                                     D, L, Lq, P, (float*)out);
  return (int)cudaGetLastError();
}
