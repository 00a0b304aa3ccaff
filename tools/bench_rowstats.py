"""rowstats (LayerNorm statistics pass) vs layernorm at the adapter's shapes (batch 8). python tools/bench_rowstats.py"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mmsam_b200  # noqa
from mmsam_b200 import kernels as K


def t(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n * 1e3


for rows, C in ((172032, 1024), (32768, 1024), (524288, 96), (131072, 192), (32768, 384), (8192, 768)):
    x = torch.randn(rows, C, device="cuda").to(torch.bfloat16)
    g, b = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
    st = torch.empty(rows, 2, device="cuda")
    us_s = t(lambda: K.rowstats(x, 1e-6, out=st))
    y = torch.empty_like(x)
    us_l = t(lambda: K.layernorm(x, g, b, 1e-6, out=y))
    by = rows * C * 2
    print(f"rows {rows} C {C}: rowstats {us_s:.1f} us ({by / us_s / 1e3:.0f} GB/s read), layernorm {us_l:.1f} us ({2 * by / us_l / 1e3:.0f} GB/s r+w)")
