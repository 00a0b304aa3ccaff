"""GEMM kernel alone for ncu: qkv-shaped (K=1024, no act), lin1 (GELU), lin2 (K=4096 + residual)."""
import math, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mmsam_b200
from mmsam_b200 import kernels as K
for (M, N, Kd, act, res) in ((32768, 3072, 1024, None, False), (32768, 4096, 1024, "gelu", False), (32768, 1024, 4096, None, True)):
    a = torch.randn(M, Kd, device="cuda").to(torch.bfloat16)
    w = (torch.randn(N, Kd, device="cuda") / math.sqrt(Kd)).to(torch.bfloat16)
    b = torch.randn(N, device="cuda")
    r = torch.randn(M, N, device="cuda").to(torch.bfloat16) if res else None
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    for _ in range(2):
        K.gemm(a, w, bias=b, act=act, residual=r, out=out)
    torch.cuda.synchronize()
print("done")
