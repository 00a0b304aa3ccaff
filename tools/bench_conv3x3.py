"""Grouped 3x3 conv (csrc/conv3x3.cu) at the fusion neck's shapes (batch 8, 1024^2 input): qkv2 (groups 32) and
Mlp.dwconv (2 channels per group) of the four pyramid levels. python tools/bench_conv3x3.py"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mmsam_b200  # noqa
from mmsam_b200 import kernels as K

B = 8
tot = 0.0
for lvl, (hw, ci) in enumerate([(256, 96), (128, 192), (64, 384), (32, 768)]):
    for name, C, groups, cnt in (("qkv2", 3 * ci, 32, 2), ("mlp.dwconv", 4 * ci, 2 * ci, 1)):
        w = torch.randn(C, C // groups, 3, 3, device="cuda") * 0.1
        wp = K.pack_conv3x3_weight(w, groups)
        x = torch.randn(B * hw * hw, C, device="cuda").to(torch.bfloat16)
        out = torch.empty_like(x)
        def run():
            K.conv3x3(x, wp, B, hw, hw, C, C, groups, out=out)
        for _ in range(2):
            run()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(10):
            run()
        e.record(); torch.cuda.synchronize()
        us = s.elapsed_time(e) * 100
        tot += us * cnt
        by = 2 * x.numel() * 2
        print(f"level {lvl} {name:10s} C={C:5d} groups={groups:4d}: {us:7.1f} us  {by / us / 1e3:6.0f} GB/s (in+out)  x{cnt}")
print(f"per forward: {tot / 1e3:.2f} ms")
