/* mmsam_b200 — C ABI of the B200-native MM SAM-Adapter hot path.
 *
 * Plain pointers and sizes only (no torch types). All pointers are DEVICE pointers unless a name
 * ends in _host; every call is asynchronous on `stream` (a cudaStream_t passed as void*), never
 * allocates, never synchronises, and the caller owns every buffer. Return value: 0 = ok,
 * > 0 = cudaError_t of the failed launch, < 0 = argument error (see MMSAM_ERR_*).
 *
 * Citations are into the reference tree (segmentation/...), naming the interface each entry
 * point replaces.
 */
#ifndef MMSAM_B200_H_
#define MMSAM_B200_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define MMSAM_OK 0
#define MMSAM_ERR_BAD_ARG (-1)
#define MMSAM_ERR_BAD_DTYPE (-2)
#define MMSAM_ERR_UNSUPPORTED (-3)
#define MMSAM_ERR_DRIVER (-4)

/* dtype codes */
#define MMSAM_F32 0
#define MMSAM_F16 1
#define MMSAM_BF16 2
#define MMSAM_F64 3

/* Library / build identification: returns the sm arch the kernels were compiled for (100). */
int mmsam_arch(void);

/* Multi-scale deformable attention, forward.
 * Replaces MSDA.ms_deform_attn_forward = ms_deform_attn_cuda_forward
 * (ops/src/vision.cpp:13-16, ops/src/cuda/ms_deform_attn_cuda.cu:20-80) and the kernel
 * ms_deformable_im2col_gpu_kernel (ops/src/cuda/ms_deform_im2col_cuda.cuh:237-299).
 *   value [N,S,M,D] (value_dtype), spatial_shapes [L,2] int64 (H,W), level_start_index [L] int64,
 *   sampling_loc [N,Lq,M,L,P,2] (aux_dtype; (x,y) in [0,1]), attn_weight [N,Lq,M,L,P] (aux_dtype),
 *   out [N,Lq,M*D] (value_dtype). All contiguous. No im2col_step: the whole batch is one launch. */
int mmsam_msda_forward(const void* value, const int64_t* spatial_shapes_dev,
                       const int64_t* level_start_index_dev, const void* sampling_loc,
                       const void* attn_weight, void* out, int N, int S, int M, int D, int Lq, int L,
                       int P, int value_dtype, int aux_dtype, void* stream);

/* Row LayerNorm over the last dim of a bf16 [rows, C] matrix, fp32 affine, biased variance.
 * Replaces nn.LayerNorm / LayerNorm2d calls on the path (base/image_encoder.py:398,421;
 * adapter_modules_...new.py:494-501,527-532; mmpretrain_custom/models/utils/norm.py:52-87).
 * row_map_dev (optional int32[rows]): destination row of each source row, -1 = drop. */
int mmsam_layernorm_bf16(const void* x, const float* gamma, const float* beta, void* y,
                         const int* row_map_dev, long long rows, int C, long long ldx, long long ldy,
                         float eps, void* stream);

/* out = epilogue(A[M,K] . W[N,K]^T): bf16 operands, fp32 accumulation on tcgen05 tensor cores.
 * Replaces every nn.Linear / 1x1 conv / patchified conv on the path (see csrc/gemm.cu header).
 *   epilogue: (+bias[N]) -> act (0 none, 1 exact GELU, 2 ReLU, 3 ReLU6) -> (*scale[N]) ->
 *             (+residual[dst_row, col]) -> store bf16 (out_f32=0) or fp32 (out_f32=1)
 *   row_mode 0: dst_row = row; 1: dst_row = row_map_dev[row] (-1 drops the row);
 *            2: 2x2 pixel shuffle of a ConvTranspose2d(k=2,s=2): rows are (b, y<ps_h, x<ps_w),
 *               columns (dy, dx, c<ps_c) -> dst_row (b, 2y+dy, 2x+dx), dst col c.
 *   block_n: 64/128/256 tile width, 0 = auto. max_ctas: 0 = all SMs. */
int mmsam_gemm_bf16(const void* A, long long lda, const void* W, long long ldw, const float* bias,
                    const float* scale, const void* residual, long long ldr, void* out,
                    long long ldo, int M, int N, int K, int act, int out_f32, int row_mode,
                    const int* row_map_dev, int ps_h, int ps_w, int ps_c, int block_n, int max_ctas,
                    void* stream);

/* Fused multi-head attention (head_dim 64) with SAM's decomposed relative-position bias.
 * Replaces Attention.forward + add_decomposed_rel_pos (base/image_encoder.py:483-501, 587-623).
 *   qkv [Bp, T, 3, nh, 64] bf16 (the qkv Linear output), out [Bp, T, nh*64] bf16,
 *   tab_h / tab_w: bf16 [pad16(2*Kh-1), 64] / [pad16(2*Kw-1), 64] relative-position tables
 *   (row r = q - k + K - 1, zero padded; both NULL = no bias), T == Kh*Kw, scale = 64^-0.5.
 *   All T keys take part in the softmax (SAM does not mask its zero-padded window tokens). */
int mmsam_attention_bf16(const void* qkv, void* out, const void* tab_h, const void* tab_w, int Bp, int T,
                         int nh, int Kh, int Kw, float scale, int max_ctas, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MMSAM_B200_H_ */
