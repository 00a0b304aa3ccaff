"""Depthwise conv at the bench shapes (batch 8): ConvNeXt 7x7 stages 0-3 and the ConvFFN 3x3 over three grids."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mmsam_b200  # noqa
from mmsam_b200 import kernels as K
B = 8
for (name, ks, C, grids, act) in (("cnx0 7x7", 7, 96, [(256, 256)], None), ("cnx1 7x7", 7, 192, [(128, 128)], None),
                                  ("cnx2 7x7", 7, 384, [(64, 64)], None), ("cnx3 7x7", 7, 768, [(32, 32)], None),
                                  ("convffn 3x3", 3, 256, [(128, 128), (64, 64), (32, 32)], "gelu")):
    S = sum(h * w for h, w in grids)
    x = torch.randn(B, S, C, device="cuda").to(torch.bfloat16)
    if ks == 7 and os.environ.get("DW_F32", "1") == "1":
        x = x.float()                       # the ConvNeXt towers' fp32 residual stream
    w = torch.randn(ks * ks, C, device="cuda")
    b = torch.randn(C, device="cuda")
    out = torch.empty(x.shape, dtype=torch.bfloat16, device="cuda")
    for _ in range(3):
        K.dwconv(x, w, b, ks, grids, B, C, S * C, S * C, act=act, out=out)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10):
        K.dwconv(x, w, b, ks, grids, B, C, S * C, S * C, act=act, out=out)
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 10
    fma = B * S * C * ks * ks
    print(f"dwconv {name}: {ms * 1e3:.0f} us, {B * S * C * 4 / ms / 1e6:.0f} GB/s (in+out), {fma / ms / 1e9:.2f} TFMA/s")
# MobileNetV2 depthwise 3x3 + ReLU6 of the neck's four levels (2 ci channels)
for (hw, C) in ((256, 192), (128, 384), (64, 768), (32, 1536)):
    S = hw * hw
    x = torch.randn(B, S, C, device="cuda").to(torch.bfloat16)
    w = torch.randn(9, C, device="cuda")
    out = torch.empty_like(x)
    for _ in range(3):
        K.dwconv(x, w, None, 3, [(hw, hw)], B, C, S * C, S * C, act="relu6", out=out)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10):
        K.dwconv(x, w, None, 3, [(hw, hw)], B, C, S * C, S * C, act="relu6", out=out)
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 10
    print(f"dwconv neck 3x3 {hw}^2 x {C}: {ms * 1e3:.0f} us, {B * S * C * 4 / ms / 1e6:.0f} GB/s (in+out)")
