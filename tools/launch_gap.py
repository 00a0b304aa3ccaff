"""Per-node cost of a CUDA-graph replay for this library's kernels: chains of tiny launches (work ~ 0) on one stream.
python tools/launch_gap.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mmsam_b200  # noqa
from mmsam_b200 import kernels as K

def timed_graph(fn, n, reps=20):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(n):
                fn()
    torch.cuda.synchronize()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        g.replay()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps / n * 1e3

x = torch.randn(256, 1024, device="cuda").to(torch.bfloat16)
w = torch.ones(1024, device="cuda"); bz = torch.zeros(1024, device="cuda")
o = torch.empty_like(x)
a = torch.randn(256, 256, device="cuda").to(torch.bfloat16)
wt = torch.randn(256, 256, device="cuda").to(torch.bfloat16)
og = torch.empty(256, 256, device="cuda", dtype=torch.bfloat16)
ln = lambda: K.layernorm(x, w, bz, 1e-6, out=o)
gm = lambda: K.gemm(a, wt, out=og)
def both():
    ln(); gm()
print(f"layernorm chain (tiny):            {timed_graph(ln, 200):6.2f} us per node")
print(f"tcgen05 GEMM chain (one tile):     {timed_graph(gm, 200):6.2f} us per node")
print(f"alternating layernorm / GEMM:      {timed_graph(both, 100) / 2:6.2f} us per node")
# same at full width: 148 CTAs with 227 KB of shared memory each
A = torch.randn(32768, 256, device="cuda").to(torch.bfloat16)
W = torch.randn(1024, 256, device="cuda").to(torch.bfloat16)
O = torch.empty(32768, 1024, device="cuda", dtype=torch.bfloat16)
big = lambda: K.gemm(A, W, out=O)
t1 = timed_graph(big, 50)
X = torch.randn(32768, 1024, device="cuda")
Ob = torch.empty(32768, 1024, device="cuda", dtype=torch.bfloat16)
lnb = lambda: K.layernorm(X, w, bz, 1e-6, out=Ob)
t2 = timed_graph(lnb, 50)
def alt():
    lnb(); big()
t3 = timed_graph(alt, 25)
print(f"full-chip GEMM 32768x1024x256: {t1:6.1f} us, layernorm 32768x1024: {t2:6.1f} us, alternating pair: {t3:6.1f} us -> {t3 - t1 - t2:+.1f} us per pair over the sum")
