import os, sys, torch, math
sys.path.insert(0, "/root/repo")
import mmsam_b200
from mmsam_b200 import kernels as K
def timed(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n * 1e3
for (M, N, Kd, f32) in ((172032, 1024, 256, False), (172032, 1024, 512, False), (32768, 1024, 1024, True), (32768, 1024, 4096, True)):
    a = torch.randn(M, Kd, device="cuda").to(torch.bfloat16)
    w = (torch.randn(N, Kd, device="cuda") / math.sqrt(Kd)).to(torch.bfloat16)
    b = torch.randn(N, device="cuda")
    dt = torch.float32 if f32 else torch.bfloat16
    r = torch.randn(M, N, device="cuda").to(dt)
    out = torch.empty(M, N, device="cuda", dtype=dt)
    us = timed(lambda: K.gemm(a, w, bias=b, residual=r, out=out, out_dtype=dt))
    by = M * Kd * 2 + M * N * (8 if f32 else 4)
    print(f"M={M} N={N} K={Kd} {'f32' if f32 else 'bf16'} residual: {us:.1f} us  {2.0*M*N*Kd/us/1e6:.0f} TFLOP/s  {by/us/1e3:.0f} GB/s")
