"""Oracle: multi-scale deformable attention core (CPU, any float dtype). TEST INFRASTRUCTURE.

Follows the arithmetic of the reference CUDA kernel ms_deformable_im2col_gpu_kernel /
ms_deform_attn_im2col_bilinear (segmentation/ops/src/cuda/ms_deform_im2col_cuda.cuh:33-84,
237-299): h_im = loc_y*H - 0.5, w_im = loc_x*W - 0.5, the sample contributes only when
-1 < h_im < H and -1 < w_im < W, each of the four corners contributes only when it is inside the
map. This is equivalent to the reference's CPU ground truth ms_deform_attn_core_pytorch
(segmentation/ops/functions/ms_deform_attn_func.py:53-75; grid_sample bilinear / zeros /
align_corners=False) and is pinned against it in tests/test_oracle_golden.py.
"""
import torch


def ms_deform_attn_core(value, spatial_shapes, sampling_locations, attention_weights):
    """value [N,S,M,D]; spatial_shapes: sequence of (H,W); sampling_locations [N,Lq,M,L,P,2] (x,y);
    attention_weights [N,Lq,M,L,P]  ->  [N,Lq,M*D]"""
    N, S, M, D = value.shape
    _, Lq, _, L, P, _ = sampling_locations.shape
    shapes = [(int(h), int(w)) for h, w in (spatial_shapes.tolist() if torch.is_tensor(spatial_shapes) else spatial_shapes)]
    out = value.new_zeros((N, Lq, M, D))
    n_idx = torch.arange(N).view(N, 1, 1, 1).expand(N, Lq, M, P)
    m_idx = torch.arange(M).view(1, 1, M, 1).expand(N, Lq, M, P)
    start = 0
    for l, (H, W) in enumerate(shapes):
        v = value[:, start:start + H * W]                      # [N, HW, M, D]
        start += H * W
        loc = sampling_locations[:, :, :, l]                   # [N,Lq,M,P,2]
        aw = attention_weights[:, :, :, l]                     # [N,Lq,M,P]
        w_im = loc[..., 0] * W - 0.5
        h_im = loc[..., 1] * H - 0.5
        inside = (h_im > -1) & (w_im > -1) & (h_im < H) & (w_im < W)
        h0 = torch.floor(h_im)
        w0 = torch.floor(w_im)
        lh, lw = h_im - h0, w_im - w0
        h0, w0 = h0.long(), w0.long()
        acc = value.new_zeros((N, Lq, M, P, D))
        for dh, dw, wt in ((0, 0, (1 - lh) * (1 - lw)), (0, 1, (1 - lh) * lw),
                           (1, 0, lh * (1 - lw)), (1, 1, lh * lw)):
            hh, ww = h0 + dh, w0 + dw
            ok = inside & (hh >= 0) & (hh <= H - 1) & (ww >= 0) & (ww <= W - 1)
            pos = (hh.clamp(0, H - 1) * W + ww.clamp(0, W - 1))
            g = v[n_idx, pos, m_idx]                           # [N,Lq,M,P,D]
            acc = acc + g * (wt * ok.to(value.dtype)).unsqueeze(-1)
        out = out + (acc * aw.unsqueeze(-1)).sum(3)
    return out.reshape(N, Lq, M * D)
