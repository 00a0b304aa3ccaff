"""Memory-bound residual GEMMs alone (for ncu --set full): proj shape (fp32 residual in / out) and the ConvFFN fc2 shape (bf16)."""
import math, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mmsam_b200  # noqa
from mmsam_b200 import kernels as K
for (M, N, Kd, f32) in ((32768, 1024, 1024, True), (172032, 1024, 256, False)):
    a = torch.randn(M, Kd, device="cuda").to(torch.bfloat16)
    w = (torch.randn(N, Kd, device="cuda") / math.sqrt(Kd)).to(torch.bfloat16)
    b = torch.randn(N, device="cuda")
    dt = torch.float32 if f32 else torch.bfloat16
    r = torch.randn(M, N, device="cuda").to(dt)
    out = torch.empty(M, N, device="cuda", dtype=dt)
    for _ in range(3):
        K.gemm(a, w, bias=b, residual=r, out=out, out_dtype=dt)
    torch.cuda.synchronize()
print("done")
