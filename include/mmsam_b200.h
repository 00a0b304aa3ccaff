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

/* Multi-scale deformable attention, backward (training; the inference path never calls it).
 * Replaces MSDA.ms_deform_attn_backward = ms_deform_attn_cuda_backward (ops/src/vision.cpp:13-16,
 * ops/src/cuda/ms_deform_attn_cuda.cu:83-153) and the ms_deformable_col2im_* kernels
 * (ops/src/cuda/ms_deform_im2col_cuda.cuh:301-921), as called by MSDeformAttnFunction.backward
 * (ops/functions/ms_deform_attn_func.py:39-50). Inputs as in mmsam_msda_forward plus grad_output [N,Lq,M*D]; every
 * floating tensor has the same dtype (MMSAM_F32 or MMSAM_F64, the two the reference's gradient checks use).
 * Outputs: grad_value [N,S,M,D] (zeroed by this call, accumulated with atomics like the reference's),
 * grad_sampling_loc [N,Lq,M,L,P,2], grad_attn_weight [N,Lq,M,L,P] (plain stores: deterministic). */
int mmsam_msda_backward(const void* value, const int64_t* spatial_shapes_dev, const int64_t* level_start_index_dev,
                        const void* sampling_loc, const void* attn_weight, const void* grad_output, void* grad_value,
                        void* grad_sampling_loc, void* grad_attn_weight, int N, int S, int M, int D, int Lq, int L, int P,
                        int dtype, void* stream);

/* Row LayerNorm over the last dim of a [rows, C] matrix, fp32 affine, biased variance. x_dtype / y_dtype: MMSAM_BF16 or
 * MMSAM_F32 (bf16 -> bf16, fp32 -> bf16, fp32 -> fp32; the fp32 forms serve the fp32 residual streams).
 * Replaces nn.LayerNorm / LayerNorm2d calls on the path (base/image_encoder.py:398,421;
 * adapter_modules_...new.py:494-501,527-532; mmpretrain_custom/models/utils/norm.py:52-87).
 * row_map_dev (optional int32[rows]): destination row of each source row, -1 = drop.
 * ps_h, ps_w > 0 (and no row map): rows are (b, y<ps_h, x<ps_w); row goes to row (b, y/2, x/2), column
 * block (y&1)*2+(x&1) of a [rows/4, 4C] matrix = the 2x2 patchify of ConvNeXt's LN2d -> Conv2d(k2,s2)
 * downsample (base/twin_convnext.py:313-336).
 * y2 (optional): second output x + LN(x) with the same row mapping (GFE residual, adapter_modules_...new.py:143). */
int mmsam_layernorm(const void* x, int x_dtype, const float* gamma, const float* beta, void* y, int y_dtype, void* y2,
                    const int* row_map_dev, long long rows, int C, long long ldx, long long ldy,
                    float eps, int ps_h, int ps_w, void* stream);

/* out = epilogue(A[M,K] . W[N,K]^T): bf16 operands, fp32 accumulation on tcgen05 tensor cores.
 * Replaces every nn.Linear / 1x1 conv / patchified conv on the path (see csrc/gemm.cu header).
 *   epilogue: (+bias[N]) -> act (0 none, 1 exact GELU, 2 ReLU, 3 ReLU6) -> (*scale[N]) ->
 *             (+residual[dst_row, col]) -> store bf16 (out_f32=0) or fp32 (out_f32=1)
 *             residual is bf16, or fp32 when res_f32 != 0 (then out_f32 must be set: the fp32 residual streams)
 *   row_mode 0: dst_row = row; 1: dst_row = row_map_dev[row] (-1 drops the row);
 *            2: 2x2 pixel shuffle of a ConvTranspose2d(k=2,s=2): rows are (b, y<ps_h, x<ps_w),
 *               columns (dy, dx, c<ps_c) -> dst_row (b, 2y+dy, 2x+dx), dst col c.
 *   block_n: 64/128/256 tile width, 0 = auto. max_ctas: 0 = all SMs. */
int mmsam_gemm_bf16(const void* A, long long lda, const void* W, long long ldw, const float* bias,
                    const float* scale, const void* residual, int res_f32, long long ldr, void* out,
                    long long ldo, int M, int N, int K, int act, int out_f32, int row_mode,
                    const int* row_map_dev, int ps_h, int ps_w, int ps_c, int block_n, int max_ctas,
                    void* stream);

/* Grouped form of mmsam_gemm_bf16 for per-image weights: rows [g * rows_per_group, (g+1) * rows_per_group) of A use the
 * weight matrix W[g] of W [G * N, K], G = M / rows_per_group; bias / scale [N] are shared; bf16 output, identity rows.
 * rows_per_group % 256 == 0 (a 256-row tile never straddles two groups). The fusion neck's per-image channel-attention
 * and cross-modal attention products (adapter_modules_...new.py:104-107, 255-259) in one launch per batch. */
int mmsam_gemm_grouped_bf16(const void* A, long long lda, const void* W, long long ldw, const float* bias,
                            const float* scale, const void* residual, long long ldr, void* out, long long ldo, int M,
                            int N, int K, int rows_per_group, int act, int block_n, int max_ctas, void* stream);

/* Fused multi-head attention (head_dim 64) with SAM's decomposed relative-position bias.
 * Replaces Attention.forward + add_decomposed_rel_pos (base/image_encoder.py:483-501, 587-623).
 *   qkv [Bp, T, 3, nh, 64] bf16 (the qkv Linear output), out [Bp, T, nh*64] bf16,
 *   tab_h / tab_w: bf16 [pad16(2*Kh-1), 64] / [pad16(2*Kw-1), 64] relative-position tables
 *   (row r = q - k + K - 1, zero padded; both NULL = no bias), T == Kh*Kw, scale = 64^-0.5.
 *   All T keys take part in the softmax (SAM does not mask its zero-padded window tokens).
 *   out_row_map_dev (optional int32[Bp*T]): destination row of each (b', t) output row, -1 = drop: fuses
 *   window_unpartition (base/image_encoder.py:529-551) into the store. */
int mmsam_attention_bf16(const void* qkv, void* out, const int* out_row_map_dev, const void* tab_h,
                         const void* tab_w, int Bp, int T, int nh, int Kh, int Kw, float scale, int max_ctas,
                         void* stream);

/* SAM window attention with the window un-partition fused into the store (base/image_encoder.py:399-416, 529-551):
 * qkv = bf16 [B * nwh * nww, 196, 3, nh, 64] in window order (window (b, wy, wx) = (b * nwh + wy) * nww + wx, 14 x 14 tokens
 * each, zero-padded windows at the right / bottom edge included: nwh = ceil(H / 14), nww = ceil(W / 14)), out = bf16
 * [B, H, W, nh * 64], the token map proj consumes. Same arithmetic as mmsam_attention_bf16 with Kh = Kw = 14 (padded tokens
 * are ordinary keys); each 126- / 70-row tile leaves through one 4-D TMA store that the copy engine clips at the image
 * edge. tab_h / tab_w as in mmsam_attention_bf16 (both or neither). */
int mmsam_attention_window_bf16(const void* qkv, void* out, const void* tab_h, const void* tab_w, int B, int H, int W, int nh,
                                float scale, int max_ctas, void* stream);

/* stats[row] = (mean, 1 / sqrt(var + eps)) fp32 pairs of the rows of a bf16 [rows, C] matrix (row stride ldx), the
 * statistics nn.LayerNorm would use (biased variance). C % 8 == 0, C <= 2048. */
int mmsam_rowstats_bf16(const void* x, float* stats, long long rows, int C, long long ldx, float eps, void* stream);

/* LayerNorm -> Linear folded into one GEMM (Injector / Extractor query_norm / feat_norm / ffn_norm feeding
 * sampling_offsets | attention_weights, value_proj and ffn.fc1, adapter_modules_...new.py:490-542):
 *   out = act(rowstat[m].rstd * (A W^T - rowstat[m].mean * colsum[n]) + bias[n]) (+ residual)
 * with W = bf16(gamma (.) W_linear), colsum[n] = sum_k W[n,k], bias[n] = sum_k beta[k] W_linear[n,k] + b_linear[n],
 * rowstat from mmsam_rowstats_bf16 on A. Same layouts / restrictions as mmsam_gemm_bf16 (identity row mode). */
int mmsam_gemm_ln_bf16(const void* A, long long lda, const void* W, long long ldw, const float* bias,
                       const float* colsum, const float* rowstat, const void* residual, long long ldr, void* out,
                       long long ldo, int M, int N, int K, int act, int out_f32, int block_n, int max_ctas,
                       void* stream);

/* ConvNeXt block tail (twin_convnext.py:98-132: norm -> pwconv1 -> GELU -> pwconv2 -> gamma -> residual) in one launch:
 *   t[m, :] += gamma (.) ( GELU( LN(y[m, :]) W1^T + b1 ) W2^T + b2 ),   LN over the C channels (biased variance, eps)
 * y = bf16 [M, C] (row stride ldy: the depthwise conv output), t = fp32 [M, C] (row stride ldt), updated IN PLACE (each
 * element receives exactly one fp32 add, by a TMA reduction: deterministic). The LayerNorm AFFINE is pre-folded by the
 * caller: W1 = bf16(ln_weight (.) W_pwconv1) [4C, C] contiguous, bias1 = W_pwconv1 ln_bias + b_pwconv1; the normalisation
 * itself ((y - mean) * rstd, rounded to bf16) is applied by the kernel to the tile it holds in shared memory. colsum1 is
 * unused (kept from the first version of this entry point, which folded the statistics into the epilogue) and may be
 * null. W2 = bf16 [C, 4C] contiguous; gamma may be null (= 1). C in {96, 192, 384} (the output tile [128 x C] fp32 stays
 * in tensor memory; C = 768 does not fit -> MMSAM_ERR_UNSUPPORTED, use LayerNorm + two GEMMs).
 * Pointers 16-byte aligned, ldy % 8 == 0, ldt % 4 == 0. */
int mmsam_convnext_mlp_bf16(const void* y, long long ldy, const void* W1, const float* colsum1, const float* bias1,
                            const void* W2, const float* bias2, const float* gamma, float* t, long long ldt, int M, int C,
                            float eps, int max_ctas, void* stream);

/* MSDeformAttn core fused with its front end (bf16 value/out): reads the raw fp32 output of the
 * query projection (columns [M*L*P*2 sampling offsets | M*L*P attention logits], row stride ldq;
 * ops/modules/ms_deform_attn.py:108-113) and the per-query reference point ref_xy [Lq,2] (broadcast
 * over batch and levels, adapter_modules_...new.py:397-431), doing softmax over L*P and
 * loc = ref + off / (W_l, H_l) (ms_deform_attn.py:115-119) in registers. P must be 4. */
int mmsam_msda_fused_bf16(const void* value, const int64_t* spatial_shapes_dev,
                          const int64_t* level_start_index_dev, const float* qproj, long long ldq,
                          const float* ref_xy, void* out, int N, int S, int M, int D, int Lq, int L, int P,
                          void* stream);

/* Same contract as mmsam_msda_fused_bf16, for callers that know the geometry on the host (the backbone does:
 * deform_inputs, adapter_modules_...new.py:397-431): the value windows are staged in shared memory by TMA and
 * gathered from there (~3x the L1 gather rate; off / (W, H) is computed as off * (1 / W, 1 / H): exact for power-of-two
 * maps, else an ulp of the location). level_hw_host [L][2] = (H_l, W_l), levels packed back to back in
 * value (sum = S); qgrid_hw_host [n_qgrids][2]: the queries are n_qgrids <= 3 row-major grids back to back
 * (sum = Lq) whose reference points are the cell centres; a CTA owns one head of a tile_h x tile_w tile of the
 * anchor_h x anchor_w anchor grid (a partition of the normalised plane) and serves every query whose cell centre
 * falls inside. prior_min_host [L][M][2] = per level/head minimum over the P points of the sampling_offsets bias
 * (x, y; ops/modules/ms_deform_attn.py:64-74), prior_ext_host [L][2] = max over heads of (max - min): they only
 * place and size the staged boxes; a sample that leaves its box is gathered from global memory, so results do not
 * depend on them. Returns MMSAM_ERR_UNSUPPORTED (-3) when the boxes do not fit shared memory or M > 16, D != 32,
 * P != 4, L > 4: call mmsam_msda_fused_bf16 instead. *_host arrays are HOST pointers read before the call returns. */
int mmsam_msda_fused_staged_bf16(const void* value, const int* level_hw_host, const float* qproj, long long ldq,
                                 const float* ref_xy, void* out, int N, int S, int M, int D, int Lq, int L, int P,
                                 int n_qgrids, const int* qgrid_hw_host, int anchor_h, int anchor_w, int tile_h,
                                 int tile_w, const float* prior_min_host, const float* prior_ext_host, int margin,
                                 void* stream);

/* Depthwise k x k conv (k = 3 or 7, stride 1, zero "same" padding) on channels-last maps (input bf16, or fp32 for k = 7:
 * x_dtype; output bf16), fp32
 * weights given tap-major [k*k][C], optional bias[C], act 0 none / 1 exact GELU / 3 ReLU6.
 * Up to 3 grids per batch item share the weights (ConvFFN DWConv over the 128^2|64^2|32^2 token
 * grids, adapter_modules_...new.py:456-471); grid i starts grid_*_off_host[i] elements into a batch
 * item of in_bstride / out_bstride elements. Also ConvNeXt's 7x7 (base/twin_convnext.py:98-101).
 * The *_host arrays are HOST pointers read before the call returns. */
int mmsam_dwconv(const void* x, int x_dtype, void* y, const float* w_tap_major, const float* bias, int B, int C,
                      int ksize, int ngrids, const int* grid_hw_host, const long long* grid_in_off_host,
                      const long long* grid_out_off_host, long long in_bstride, long long out_bstride, int act,
                      void* stream);

/* One modality's decoded image, HWC uint8 [B, H, W, C] (C <= 4, W % 4 == 0) -> channels [c_off, c_off + C) of the fp32
 * NCHW network input [B, Ctot, H, W]: out = (v * prescale - mean[c]) / std[c] (Normalize_multimodal with norm_by_max:
 * prescale = 1/255, pipelines/transform.py:2796-2806; + ImageToTensor's HWC -> CHW). mean / std are HOST pointers. */
int mmsam_normalize_u8(const void* img_hwc_u8, float* out_nchw, int B, int H, int W, int C, int Ctot, int c_off,
                       const float* mean_host, const float* std_host, float prescale, void* stream);

/* One modality's decoded image, HWC uint8 [B, Hs, Ws, C] (C <= 4), straight to the bf16 patch rows
 * [(b,py,px), (c,ky,kx)] of the p x p / stride p conv that consumes it (p = 4: ConvNeXt stem, 16: ViT patch embed):
 * zero-extended to H x W with pad_val (Pad_multimodal, pipelines/transform.py:2934, applied before the normalisation),
 * (v * prescale - mean[c]) / std[c] (Normalize_multimodal, :2717-2815; norm_by_max: prescale = 1/255), swap_rb != 0
 * reverses the channel order first (to_rgb of mmcv.imnormalize). Replaces normalise + ImageToTensor + patchify: the
 * fp32 NCHW network input is never materialised. mean / std are HOST pointers. */
int mmsam_patchify_u8(const void* img_hwc_u8, void* out, int B, int Hs, int Ws, int C, int H, int W, int p,
                      const float* mean_host, const float* std_host, float prescale, float pad_val, int swap_rb,
                      void* stream);

/* NCHW fp32 image channels [c_off, c_off+C) -> bf16 rows [(b,py,px), (c,ky,kx)] for p x p / stride p
 * convs as GEMMs (patch embed base/image_encoder.py:662-671; ConvNeXt stem twin_convnext.py:295-312). */
int mmsam_patchify_f32(const float* img, void* out, int B, int Ctot, int c_off, int C, int H, int W, int p,
                       void* stream);

/* out[b,y,x,:] = (base[b,y,x,:] + bilinear(src[b])[y,x,:]) * scale + shift on channels-last maps (src bf16 or fp32:
 * src_dtype; base / out bf16
 * (align_corners=False; base and scale/shift optional; ld* = elements between pixels, *_bstride =
 * elements between batch items). Backbone tail (..._new.py:326-337) and head resize-into-concat
 * (decode_heads/segformer_head.py:55-61). */
int mmsam_resize_add_affine(const void* src, int src_dtype, const void* base, const float* scale, const float* shift,
                                 void* out, int B, int Hs, int Ws, int Ho, int Wo, int C, long long src_bstride,
                                 long long lds, long long base_bstride, long long ldb, long long out_bstride,
                                 long long ldo, void* stream);

/* out[b,y,x,:] = act((base[b,y,x,:] + sum_k bilinear(src_k[b] (Hs_k x Ws_k) -> Ho x Wo)[y,x,:]) * scale + shift),
 * dense channels-last bf16, nsrc <= 3 sources (src_hw_host [nsrc][2], HOST pointer), base / scale / shift optional,
 * relu != 0 applies ReLU. The Segformer head's concat + fusion conv + BN + ReLU (decode_heads/segformer_head.py:55-64)
 * with the fusion weight applied per level BEFORE the (linear) resize. */
int mmsam_resize_sum_affine_bf16(const void* base, int nsrc, const void* src0, const void* src1, const void* src2,
                                 const int* src_hw_host, const float* scale, const float* shift, int relu, void* out,
                                 int B, int Ho, int Wo, int C, void* stream);

/* labels_u8[b, y<Hc, x<Wc] = argmax_c bilinear(logits[b] (hs x ws x ldl fp32, ncls valid) -> Ho x Wo):
 * resize + softmax + argmax (+ crop) of segmentors/encoder_decoder.py:96-117, 329-414, 449, 477. */
int mmsam_upsample_argmax_f32(const float* logits, void* labels_u8, int B, int hs, int ws, int ldl, int ncls,
                              int Ho, int Wo, int Hc, int Wc, void* stream);

/* Bilinear resize (align_corners=False) of channels-last fp32 logits [B, hs, ws, ld] -> [B, Ho, Wo, ld] (ld % 4 == 0): the
 * chained resizes of EncoderDecoder when they are not identities (encode_decode -> image size, then whole_inference ->
 * ori_shape / whole_inference_dim -> test_cfg.dim; encoder_decoder.py:103-107, 317-325, 341-346). */
int mmsam_resize_logits_f32(const float* src, float* dst, int B, int hs, int ws, int ld, int Ho, int Wo, void* stream);

/* slide_inference after the head (encoder_decoder.py:198-226) in one pass over the frame: crop j of image b is
 * crop_logits[j * B + b] (fp32 [hs, ws, ld] at the head's resolution); crop_boxes_host [ncrops][4] = (y1, x1, y2, x2) in
 * frame pixels (HOST pointer, ncrops <= 16). Per frame pixel: sum over the covering crops of the crop logits resized to
 * the crop size, divided by the overlap count; labels_u8 [B, H, W] = argmax (optional), preds fp32 [B, H, W, ld] (optional:
 * when a rescale to ori_shape follows, :223-229). ld <= 32. */
int mmsam_slide_merge_f32(const float* crop_logits, void* labels_u8, float* preds, int B, int hs, int ws, int ld, int ncls,
                          int H, int W, int ncrops, const int* crop_boxes_host, void* stream);

/* conf_u64[gt*ncls + pred] += 1 for every pixel with gt != ignore_index: device-side
 * intersect_and_union (mmseg_custom/apis/evaluation/metrics_micro.py:26-86). */
int mmsam_confusion_u8(const void* pred_u8, const void* gt_u8, void* conf_u64, long long n, int ncls,
                       int ignore_index, void* stream);

/* Grouped 3x3 conv (stride 1, pad 1, no bias) on channels-last bf16 as a tcgen05 implicit GEMM.
 * Replaces AttentionBase.qkv2 and Mlp.dwconv of the fusion neck (adapter_modules_...new.py:84, 118-119).
 * w_packed: bf16 [(ceil(Cout/NS) * 9 * KC) * 64, 64], NS = mmsam_conv3x3_nstride(Cin, Cout, groups) output channels
 * per n-tile, KC = mmsam_conv3x3_kblocks(Cin, Cout, groups); block (n-tile, tap, k-block) = W[co, ci, tap] for the
 * tile's co (rows >= NS zero), ci in the 64-channel window starting at
 * (((n-tile*NS) / (Cout/groups)) * (Cin/groups) & ~7) + k-block*64, zero outside co's group. */
int mmsam_conv3x3_kblocks(int Cin, int Cout, int groups);
/* Output channels per n-tile (<= 64) the kernel and the weight packing use for this grouping: the stride that
 * minimises (n-tiles x k-blocks), i.e. the kernel's L2 -> shared-memory traffic. */
int mmsam_conv3x3_nstride(int Cin, int Cout, int groups);
int mmsam_conv3x3_bf16(const void* x, const void* w_packed, void* out, int B, int H, int W, int Cin, int Cout,
                       int groups, int max_ctas, void* stream);

/* Gram matrix over pixels, as per-pixel-chunk partial sums (no floating-point atomics: the consumers add the chunks in a
 * fixed order, so results are bit-reproducible):
 *   S_part[c, b, i, j] = sum_{pix in chunk c} X[b,pix,qoff+i] * X[b,pix,koff+j]      fp32 [nchunks, B, n, n]
 *   nq_part / nk_part [nchunks, B, n]: partial squared norms of the q / k columns (optional, both or neither)
 * with nchunks = mmsam_gram_chunks(n, B, HW, norms). blk > 0: only the block-diagonal (per-head) elements are written,
 * the rest of S_part is left untouched. AttentionBase q@k^T / F.normalize (adapter_modules_...new.py:98-103) and GFFM
 * energies (:250-254). X is [B, HW, ld] bf16. */
int mmsam_gram_chunks(int n, int B, int HW, int norms);
int mmsam_gram_bf16(const void* X, long long ld, int qoff, int koff, int n, int B, int HW, int blk, float* S_part,
                    float* nq_part, float* nk_part, void* stream);

/* AttentionBase after the Gram pass (adapter_modules_...new.py:100-107): att = softmax(cos(q_a, k_j) * temperature[h])
 * per head from the chunk partials, with `proj` and scale2 folded in:
 *   weff[b, i, h*ch + j] = scale2 * sum_a wproj[i, h*ch + a] * att[b, h, a, j]        bf16 [B, ci, ci]
 * so that (attn @ v) -> proj is one GEMM over the pixels with weff[b] as the weight. temperature fp32 [heads],
 * wproj fp32 [ci, ci] (proj.weight), scale2 fp32 [1] (device). */
int mmsam_gfe_weff_bf16(const float* S_part, const float* nq_part, const float* nk_part, int nchunks, int B, int ci,
                        int heads, const float* temperature, const float* wproj, const float* scale2, void* weff,
                        void* stream);

/* GFFM attention maps (:250-259) from the partials of the cross-modal energy E [B, ci, ci]:
 * ax[b,i,:] = softmax_j E[b,i,j], ay[b,j,:] = softmax_i E[b,i,j]; bf16 [B, ci, ci] each. ci <= 1024. */
int mmsam_gffm_softmax_bf16(const float* E_part, int nchunks, int B, int ci, void* ax, void* ay, void* stream);

/* GFFM LayerNorm-over-HW statistics + FFRM gate (:262-264, 148-162) from the mmsam_colstats_bf16 partials:
 * mu, rstd fp32 [B, C]; gate[b, c] = 1 + sigmoid(relu(GroupNorm_groups(wffrm . GAP(LN_HW(o)))[c])); sum_w = sum of the
 * LayerNorm weight over the HW positions, mean_b = mean of its bias. wffrm fp32 [C, C] (1x1 conv, no bias). */
int mmsam_ffrm_gate_f32(const float* colstats_part, int nchunks, int B, int HW, int C, double sum_w, double mean_b,
                        float ln_eps, const float* wffrm, const float* gn_weight, const float* gn_bias, int groups,
                        float gn_eps, float* mu, float* rstd, float* gate, void* stream);

/* CoordinateAttention vectors (:187-215) from the pools of mmsam_combine_pool_bf16: per row y / column x the pooled mean
 * [C] -> conv1 (C -> mip) -> eval BN (scale, shift) -> h_swish -> conv_h | conv_w (mip -> C) -> sigmoid:
 * ah fp32 [B, H, C], aw fp32 [B, W, C]. w1 [mip, C], wh / ww [C, mip]. */
int mmsam_ca_vectors_f32(const float* ph, const float* pw_part, int nstrips, int B, int H, int W, int C, int mip,
                         const float* w1, const float* b1, const float* bn_scale, const float* bn_shift, const float* wh,
                         const float* bh, const float* ww, const float* bw, float* ah, float* aw, void* stream);

/* Per-512-pixel-chunk column statistics {sum o, sum o^2, sum o*w[pix]} of o [B,HW,C] bf16 for GFFM's
 * LayerNorm over the spatial axis (:262-264): part fp32 [mmsam_colstats_chunks(HW), B, C, 3]. */
int mmsam_colstats_chunks(int HW);
int mmsam_colstats_bf16(const void* o, const float* wpix, float* part, int B, int HW, int C, void* stream);

/* u[r, c] = gelu(a[r, c]) * a[r, C + c], a bf16 [rows, 2C] -> u bf16 [rows, C]  (Mlp gate, :129-130). */
int mmsam_gate_bf16(const void* a, void* u, long long rows, int C, void* stream);

/* f = s1 * ((o - mu) * rstd * w[pix] + b[pix]) * gate + s2 * lo  (GFFM LayerNorm over HW, FFRM gate, Scale2;
 * :262-264, 158-162, 279-280) plus the coordinate-attention pools: ph [B,H,C] row sums and per-strip column sums
 * pw_part [ceil(H/RS), B, W, C], RS = mmsam_combine_pool_rows(H). mu/rstd/gate fp32 [B,C]; wpix/bpix fp32 [H*W]. */
int mmsam_combine_pool_rows(int H);
int mmsam_combine_pool_bf16(const void* o, const void* lo, const float* mu, const float* rstd, const float* gate,
                            const float* wpix, const float* bpix, float s1, float s2, void* f, float* ph,
                            float* pw_part, int B, int H, int W, int C, void* stream);

/* out = f * (1 + aw[b,x,c] * ah[b,y,c])  (CoordinateAttention + CA residual, :187-221); ah [B,H,C], aw [B,W,C]. */
int mmsam_ca_apply_bf16(const void* f, const float* ah, const float* aw, void* out, int B, int H, int W, int C,
                        void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MMSAM_B200_H_ */
