"""Backbone tail (resize of the ViT token map + add + eval-BN, segmentation/.../..._new.py:316-337) at batch 8:
the 256^2 / 128^2 / 64^2 / 32^2 levels, fp32 token map in, bf16 maps out. python tools/bench_resize.py"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mmsam_b200  # noqa
from mmsam_b200 import kernels as K
B, C = 8, 1024
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
src = torch.randn(B, 64, 64, C, device="cuda")
sc, sh = torch.rand(C, device="cuda") + 0.5, torch.randn(C, device="cuda")
tot = 0.0
for ho in (256, 128, 64, 32):
    base = torch.randn(B, ho, ho, C, device="cuda").to(torch.bfloat16)
    out = torch.empty_like(base)
    run = lambda: K.resize_add_affine(src, (64, 64), (ho, ho), B, C, base=base, scale=sc, shift=sh, out=out)
    run(); run()
    ts = []
    for _ in range(5):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); run(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    us = sorted(ts)[2] * 1e3
    by = 2 * base.numel() * 2 + src.numel() * 4
    tot += us
    print(f"{ho}^2: {us:.0f} us, {by / us / 1e3:.0f} GB/s")
print(f"total {tot / 1e3:.2f} ms")
