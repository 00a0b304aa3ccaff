"""`ops.modules.MSDeformAttn` with the reference's signature, attributes and parameter names
(segmentation/ops/modules/ms_deform_attn.py:28-130); compute runs on the sm_100a kernels."""
import torch
import torch.nn.functional as F

from ... import kernels as K
from ...nn_modules import MSDeformAttnParams


class MSDeformAttn(MSDeformAttnParams):
    def __init__(self, d_model=256, n_levels=4, n_heads=8, n_points=4, ratio=1.0):
        super().__init__(d_model, n_levels, n_heads, n_points, ratio)
        self._packed = None

    def _pack(self, dev):
        key = (str(dev), self.value_proj.weight._version, self.sampling_offsets.weight._version,
               self.attention_weights.weight._version, self.output_proj.weight._version,
               self.attention_weights.bias._version, self.sampling_offsets.bias._version)
        if self._packed is None or self._packed[0] != key:
            bf = lambda t: t.detach().to(device=dev, dtype=torch.bfloat16).contiguous()
            f32 = lambda t: t.detach().to(device=dev, dtype=torch.float32).contiguous()
            wq = torch.cat((self.sampling_offsets.weight.detach(), self.attention_weights.weight.detach()), 0)
            bq = torch.cat((self.sampling_offsets.bias.detach(), self.attention_weights.bias.detach()), 0)
            self._packed = (key, dict(vw=bf(self.value_proj.weight), vb=f32(self.value_proj.bias), qw=bf(wq), qb=f32(bq),
                                      ow=bf(self.output_proj.weight), ob=f32(self.output_proj.bias)))
        return self._packed[1]

    @torch.no_grad()
    def forward(self, query, reference_points, input_flatten, input_spatial_shapes, input_level_start_index,
                input_padding_mask=None):
        N, Len_q, _ = query.shape
        N, Len_in, _ = input_flatten.shape
        assert (input_spatial_shapes[:, 0] * input_spatial_shapes[:, 1]).sum() == Len_in
        if reference_points.shape[-1] not in (2, 4):
            raise ValueError("Last dim of reference_points must be 2 or 4, but get {} instead.".format(
                reference_points.shape[-1]))
        if not query.is_cuda:
            raise RuntimeError("MSDeformAttn (B200 build) needs CUDA tensors: there is no CPU fallback")
        p = self._pack(query.device)
        M, L, P = self.n_heads, self.n_levels, self.n_points
        odt = query.dtype
        q2 = query.reshape(N * Len_q, -1).to(torch.bfloat16).contiguous()
        f2 = input_flatten.reshape(N * Len_in, -1).to(torch.bfloat16).contiguous()
        value = K.gemm(f2, p["vw"], bias=p["vb"])
        if input_padding_mask is not None:
            value = value.masked_fill(input_padding_mask.reshape(-1)[:, None], 0.0)
        qp = K.gemm(q2, p["qw"], bias=p["qb"], out_dtype=torch.float32)
        shapes = input_spatial_shapes.to(torch.int64).contiguous()
        lsi = input_level_start_index.to(torch.int64).contiguous()
        rp = reference_points
        fused_ok = (rp.shape[-1] == 2 and rp.shape[0] == 1 and rp.shape[2] == 1 and P == 4
                    and (value.shape[1] // M) % 8 == 0)
        if fused_ok:
            ref = rp.reshape(Len_q, 2).float().contiguous()
            o = K.msda_fused(value.view(N, Len_in, -1), shapes, lsi, qp, ref, M, L, P)
        else:
            off = qp[:, : M * L * P * 2].reshape(N, Len_q, M, L, P, 2)
            aw = F.softmax(qp[:, M * L * P * 2:].reshape(N, Len_q, M, L * P), -1).view(N, Len_q, M, L, P)
            if rp.shape[-1] == 2:
                norm = torch.stack([shapes[..., 1], shapes[..., 0]], -1).float()
                loc = rp.float()[:, :, None, :, None, :] + off / norm[None, None, None, :, None, :]
            else:
                loc = rp.float()[:, :, None, :, None, :2] + off / P * rp.float()[:, :, None, :, None, 2:] * 0.5
            loc = loc.expand(N, Len_q, M, L, P, 2).contiguous()
            o = K.msda_forward(value.view(N, Len_in, M, -1), shapes, lsi, loc, aw.contiguous())
        out = K.gemm(o.view(N * Len_q, -1), p["ow"], bias=p["ob"])
        return out.view(N, Len_q, -1).to(odt)
