"""Seeded weight perturbation for parity runs. TEST INFRASTRUCTURE.

The reference's random init is degenerate for parity purposes: pos_embed / rel_pos are zeros
(base/image_encoder.py:248-250, 462-463), MSDeformAttn offsets/attention weights are zero
(ops/modules/ms_deform_attn.py:64-77), Injector.gamma = init_values, GFFM / MobileNetV2 scales are 0
(adapter_modules_...new.py:238-239, 292), BN running stats are (0, 1). This re-randomises exactly
those tensors (deterministically, by key order) so that every arithmetic path carries signal.
"""
import torch


def perturb_state_dict(sd, seed=1, head_std=0.5):
    g = torch.Generator().manual_seed(seed)
    out = {}

    def rn(shape, std):
        return torch.randn(shape, generator=g) * std

    for k in sorted(sd.keys()):
        v = sd[k]
        if not torch.is_floating_point(v):
            out[k] = v.clone()
            continue
        v = v.clone().float()
        last = k.rsplit(".", 1)[-1]
        if k.endswith("pos_embed"):
            v = rn(v.shape, 0.02)
        elif last in ("rel_pos_h", "rel_pos_w"):
            v = rn(v.shape, 0.2)
        elif k.endswith("sampling_offsets.weight"):
            v = rn(v.shape, 0.01)
        elif k.endswith("attention_weights.weight"):
            v = rn(v.shape, 0.02)
        elif k.endswith("attention_weights.bias"):
            v = rn(v.shape, 0.5)
        elif k.endswith("sampling_offsets.bias"):
            pass  # keep the reference's directional grid init
        elif k.endswith("injector.gamma"):
            v = 0.3 + rn(v.shape, 0.1)
        elif "twin_conv.stages_" in k and last == "gamma":
            # ConvNeXt layer scale (twin_convnext.py:58-61, init 1e-6): pretrained values are O(0.1-1); at 1e-6 the whole
            # residual branch (7x7 dwconv, LN, pw1 + GELU, pw2) would be invisible to every parity test
            v = 0.3 + rn(v.shape, 0.1)
        elif k.endswith("gammax.scale") or k.endswith("gammay.scale"):
            v = torch.tensor(0.1) + rn((), 0.02)
        elif "local_feature_encoder" in k and last == "scale":
            v = torch.tensor(0.1) + rn((), 0.02)
        elif last == "running_mean":
            v = rn(v.shape, 0.1)
        elif last == "running_var":
            v = 0.5 + torch.rand(v.shape, generator=g)
        elif "conv_seg.weight" in k:
            v = rn(v.shape, head_std)
        elif last == "bias":
            v = v + rn(v.shape, 0.02)
        elif last == "weight" and v.dim() == 1:
            v = v + rn(v.shape, 0.02)  # LN / BN / GN affine
        out[k] = v
    return out


def synthetic_batch(batch, size, kind="lidar", seed=1234):
    """SURVEY.md §8(d) synthetic inputs: [B, 6, H, W] fp32 NCHW, RGB normalised + sparse aux."""
    H, W = (size, size) if isinstance(size, int) else size
    xs = []
    mean = torch.tensor([0.485, 0.456, 0.406]).view(3, 1, 1)
    std = torch.tensor([0.229, 0.224, 0.225]).view(3, 1, 1)
    for i in range(batch):
        g = torch.Generator().manual_seed(seed + i)
        rgb = torch.randint(0, 256, (3, H, W), generator=g).float() / 255.0
        rgb = (rgb - mean) / std
        if kind == "lidar":
            mask = (torch.rand(3, H, W, generator=g) < 0.1).float()
            aux = mask * torch.randint(1, 256, (3, H, W), generator=g).float() / 255.0
        elif kind == "thermal":
            aux = (torch.randint(0, 256, (1, H, W), generator=g).float() / 255.0).expand(3, H, W)
        else:
            aux = torch.rand(3, H, W, generator=g)
        xs.append(torch.cat((rgb, aux), 0))
    return torch.stack(xs)
