"""Per-kernel timing on the GPU box (CUDA events, L2 flushed between iterations)."""
import json
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mmsam_b200  # noqa
from mmsam_b200 import kernels as KN

PEAKS = {"hbm_gbs": 6451.2, "bf16_tflops": 1658.5}
try:
    PEAKS.update(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))))
except Exception:
    pass

_flush = None


def timeit(fn, iters=10, warmup=3, flush=True):
    global _flush
    if _flush is None:
        _flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush:
            _flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


def bench_gemm(M, N, K, act=None, res=False, bn=0, name=""):
    a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda") / math.sqrt(K)).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda")
    r = torch.randn(M, N, device="cuda").to(torch.bfloat16) if res else None
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    ms = timeit(lambda: KN.gemm(a, w, bias=bias, act=act, residual=r, out=out, block_n=bn))
    ms_t = timeit(lambda: torch.nn.functional.linear(a, w))
    tf = 2.0 * M * N * K / ms / 1e9
    print(f"gemm {name:14s} M={M:6d} N={N:5d} K={K:5d} act={act} res={res} bn={bn}: {ms:.3f} ms {tf:7.1f} TFLOP/s "
          f"({tf / PEAKS['bf16_tflops'] * 100:.1f}% of measured peak) | cuBLAS {ms_t:.3f} ms")


def bench_msda(N, Lq, shapes, name, M=16, D=32, P=4):
    shapes_t = torch.as_tensor(shapes, dtype=torch.long)
    S = int(shapes_t.prod(1).sum()); L = len(shapes)
    lsi = torch.cat((shapes_t.new_zeros((1,)), shapes_t.prod(1).cumsum(0)[:-1])).cuda()
    value = torch.randn(N, S, M, D, device="cuda").to(torch.bfloat16)
    # realistic locations: reference grid + a few pixels of offset
    base = torch.rand(N, Lq, 1, 1, 1, 2, device="cuda")
    loc = (base + torch.randn(N, Lq, M, L, P, 2, device="cuda") * 0.03).contiguous()
    aw = torch.softmax(torch.randn(N, Lq, M, L * P, device="cuda"), -1).view(N, Lq, M, L, P).contiguous()
    out = torch.empty(N, Lq, M * D, device="cuda", dtype=torch.bfloat16)
    sh = shapes_t.cuda()
    ms = timeit(lambda: KN.msda_forward(value, sh, lsi, loc, aw, out=out))
    by = N * (S * M * D * 2 + Lq * M * L * P * 2 * 4 + Lq * M * L * P * 4 + Lq * M * D * 2)
    gbs = by / ms / 1e6
    print(f"msda {name:10s} N={N} Lq={Lq} S={S} L={L}: {ms:.3f} ms, {by / 1e6:.1f} MB algorithmic -> {gbs:.0f} GB/s "
          f"({gbs / PEAKS['hbm_gbs'] * 100:.1f}% of measured HBM peak)")


def bench_ln(rows, C):
    x = torch.randn(rows, C, device="cuda").to(torch.bfloat16)
    w = torch.ones(C, device="cuda"); b = torch.zeros(C, device="cuda")
    out = torch.empty_like(x)
    ms = timeit(lambda: KN.layernorm(x, w, b, 1e-6, out=out))
    gbs = rows * C * 4 / ms / 1e6
    print(f"layernorm rows={rows} C={C}: {ms:.3f} ms {gbs:.0f} GB/s ({gbs / PEAKS['hbm_gbs'] * 100:.1f}%)")


if __name__ == "__main__":
    B = 8
    which = sys.argv[1:] or ["gemm", "msda", "ln"]
    if "gemm" in which:
        bench_gemm(B * 4900, 3072, 1024, name="qkv(window)")
        bench_gemm(B * 4096, 3072, 1024, name="qkv(global)")
        bench_gemm(B * 4900, 1024, 1024, res=True, name="proj")
        bench_gemm(B * 4096, 4096, 1024, act="gelu", name="mlp.lin1")
        bench_gemm(B * 4096, 4096, 1024, act=None, name="mlp.lin1-noact")
        bench_gemm(B * 4096, 1024, 4096, res=True, name="mlp.lin2")
        bench_gemm(B * 4096, 1024, 4096, res=True, bn=128, name="mlp.lin2/128")
        bench_gemm(B * 21504, 512, 1024, name="value_proj")
        bench_gemm(B * 65536, 384, 96, act="gelu", name="cnx0.pw1")
        bench_gemm(B * 65536, 96, 384, res=True, name="cnx0.pw2")
        bench_gemm(B * 4096, 1536, 384, act="gelu", name="cnx2.pw1")
        bench_gemm(B * 4096, 384, 1536, res=True, name="cnx2.pw2")
        bench_gemm(8192, 8192, 8192, name="square8k")
    if "msda" in which:
        bench_msda(B, 4096, [(128, 128), (64, 64), (32, 32)], "injector")
        bench_msda(B, 21504, [(64, 64)], "extractor")
    if "ln" in which:
        bench_ln(B * 4096, 1024)
        bench_ln(B * 21504, 1024)
        bench_ln(B * 65536, 96)
        bench_ln(B * 16384, 192)
        bench_ln(B * 4096, 384)
        bench_ln(B * 1024, 768)
