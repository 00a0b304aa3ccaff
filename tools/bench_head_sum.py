"""resize_sum_affine at the Segformer head's shape (batch 8, 256^2 x 512 + 128^2, 64^2, 32^2). python tools/bench_head_sum.py"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mmsam_b200  # noqa
from mmsam_b200 import kernels as K
B, C = 8, 512
base = torch.randn(B * 256 * 256, C, device="cuda").to(torch.bfloat16)
srcs = [torch.randn(B * h * h, C, device="cuda").to(torch.bfloat16) for h in (128, 64, 32)]
shift = torch.randn(C, device="cuda")
out = torch.empty_like(base)
for _ in range(3):
    K.resize_sum_affine(base, srcs, [(128, 128), (64, 64), (32, 32)], (256, 256), B, C, shift=shift, relu=True, out=out)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(10):
    K.resize_sum_affine(base, srcs, [(128, 128), (64, 64), (32, 32)], (256, 256), B, C, shift=shift, relu=True, out=out)
e.record(); torch.cuda.synchronize()
us = s.elapsed_time(e) * 100
print(f"resize_sum_affine head shape: {us:.0f} us, {(2 * base.numel() * 2 + sum(t.numel() * 2 for t in srcs)) / us / 1e3:.0f} GB/s (compulsory)")
