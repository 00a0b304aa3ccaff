"""Fused ConvNeXt block tail (mmsam_convnext_mlp_bf16) vs the three launches it replaces (LayerNorm + GEMM(GELU) + GEMM(+res)) at
the step's shapes (batch 8, 1024^2: stage 0 C=96 M=524288, stage 1 C=192 M=131072, stage 2 C=384 M=32768)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mmsam_b200  # noqa
from mmsam_b200 import kernels as K


def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n * 1e3


for C, M in ((96, 524288), (192, 131072), (384, 32768)):
    y = torch.randn(M, C, device="cuda").to(torch.bfloat16)
    w1 = (torch.randn(4 * C, C, device="cuda") / C ** 0.5).to(torch.bfloat16)
    w2 = (torch.randn(C, 4 * C, device="cuda") / (4 * C) ** 0.5).to(torch.bfloat16)
    cs, b1, b2, gm = w1.float().sum(1), torch.randn(4 * C, device="cuda"), torch.randn(C, device="cuda"), torch.rand(C, device="cuda") * 1e-3
    lw, lb = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
    t = torch.randn(M, C, device="cuda")
    us = timed(lambda: K.convnext_mlp(y, w1, cs, b1, w2, b2, gm, t, 1e-6))
    yn = torch.empty_like(y)
    hbuf = torch.empty(M, 4 * C, device="cuda", dtype=torch.bfloat16)

    def old():
        K.layernorm(y, lw, lb, 1e-6, out=yn)
        K.gemm(yn, w1, bias=b1, act="gelu", out=hbuf)
        K.gemm(hbuf, w2, bias=b2, scale=gm, residual=t, out=t)
    us0 = timed(old)
    fl = 2.0 * M * C * 4 * C * 2
    by = M * C * (2 + 4 + 4)
    print(f"C={C} M={M}: fused {us:.1f} us ({fl / us / 1e6:.0f} TFLOP/s, {by / us / 1e3:.0f} GB/s algorithmic)  |  LN + 2 GEMMs {us0:.1f} us")
