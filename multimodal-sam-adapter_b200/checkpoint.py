"""Checkpoint ingestion with the reference's rules. Host-side, CPU tensors.

* unwrap / revise_keys: mmcv `load_checkpoint` (mmcv_custom/checkpoint.py:319-515): take `state_dict` (or `model`) out
  of the file's dict, strip a leading `module.` (DataParallel), load non-strictly and report what did not match.
* SAM image encoder: tools/SAM_checkpoint_convert.py:15-33 keeps the `image_encoder.*` tensors of a released SAM
  checkpoint, drops its `neck.*`, strips the prefix (-> `pretrained/sam_vit_l_image_encoder_no_neck.pth`).
* Twin ConvNeXt: base/twin_convnext.py:399-443 loads ONE single-tower ConvNeXt checkpoint into both towers by inserting
  `_x` / `_y` before the first dot of every key (`stages.0...` -> `stages_x.0...`, `norm0...` -> `norm0_x...`), after
  stripping `backbone.` / `module.`.
"""
import re
from collections import OrderedDict

import torch


def unwrap_state_dict(ckpt):
    """The tensor dict inside whatever torch.load returned."""
    if isinstance(ckpt, dict):
        for k in ("state_dict", "model"):
            if k in ckpt and isinstance(ckpt[k], dict):
                return ckpt[k]
    return ckpt


def revise_keys(sd, rules=((r"^module\.", ""),)):
    out = OrderedDict()
    for k, v in sd.items():
        for pat, rep in rules:
            k = re.sub(pat, rep, k)
        out[k] = v
    return out


def convert_sam_image_encoder(sd):
    """Released SAM checkpoint -> ImageEncoderViT keys (no neck)."""
    return OrderedDict((k.replace("image_encoder.", ""), v) for k, v in sd.items()
                       if "image_encoder" in k and "neck" not in k)


def twin_convnext_keys(sd):
    """Single-tower ConvNeXt state dict -> (x-tower dict, y-tower dict) with the reference's key surgery."""
    sd = unwrap_state_dict(sd)
    flat = OrderedDict((k[9:] if k.startswith("backbone.") else k, v) for k, v in sd.items())
    if flat and next(iter(flat)).startswith("module."):
        flat = OrderedDict((k[7:], v) for k, v in flat.items())
    sx, sy = OrderedDict(), OrderedDict()
    for k, v in flat.items():
        dot = k.find(".")
        if dot != -1:
            sx[k[:dot] + "_x" + k[dot:]] = v
            sy[k[:dot] + "_y" + k[dot:]] = v
        else:
            sx[k + "_x"] = v
            sy[k + "_y"] = v
    return sx, sy


def load_checkpoint(module, path_or_sd, strict=False, rules=((r"^module\.", ""),), map_location="cpu"):
    """mmcv-style load: returns (missing_keys, unexpected_keys). Shape mismatches raise as in torch."""
    ckpt = torch.load(path_or_sd, map_location=map_location) if isinstance(path_or_sd, str) else path_or_sd
    sd = revise_keys(unwrap_state_dict(ckpt), rules)
    res = module.load_state_dict(sd, strict=strict)
    return list(res.missing_keys), list(res.unexpected_keys)


def load_twin_convnext(twin, path_or_sd, map_location="cpu"):
    """Both towers of a TwinConvNeXt from one single-tower checkpoint; returns the keys neither load consumed."""
    ckpt = torch.load(path_or_sd, map_location=map_location) if isinstance(path_or_sd, str) else path_or_sd
    sx, sy = twin_convnext_keys(ckpt)
    rx = twin.load_state_dict(sx, strict=False)
    ry = twin.load_state_dict(sy, strict=False)
    return sorted(set(rx.unexpected_keys) | set(ry.unexpected_keys))
