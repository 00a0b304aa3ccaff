"""Gram kernel + its consumer at the neck's shapes (batch 8): AttentionBase (qkv [HW, 3ci], per-head blocks, norms ->
gfe_weff) and GFFM ([HW, 2ci] -> gffm_softmax). MMSAM_GRAM_TARGET = CTAs per image the chunk plan aims at."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mmsam_b200  # noqa
from mmsam_b200 import kernels as K
B = 8
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timeit(fn, n=7):
    fn(); fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return sorted(ts)[n // 2] * 1e3


tot = 0.0
for (HW, ci) in ((65536, 96), (16384, 192), (4096, 384), (1024, 768)):
    heads = 8
    temp = torch.ones(heads, device="cuda"); wp = torch.randn(ci, ci, device="cuda") * 0.05; s2 = torch.ones(1, device="cuda")
    x = torch.randn(B * HW, 3 * ci, device="cuda").to(torch.bfloat16)
    S, nq, nk = K.gram(x, 3 * ci, 0, ci, ci, B, HW, blk=ci // heads, norms=True)
    t_g = timeit(lambda: K.gram(x, 3 * ci, 0, ci, ci, B, HW, blk=ci // heads, norms=True))
    t_c = timeit(lambda: K.gfe_weff(S, nq, nk, temp, wp, s2, heads))
    g = torch.randn(B * HW, 2 * ci, device="cuda").to(torch.bfloat16)
    E = K.gram(g, 2 * ci, 0, ci, ci, B, HW, blk=0)
    t_g2 = timeit(lambda: K.gram(g, 2 * ci, 0, ci, ci, B, HW, blk=0))
    t_c2 = timeit(lambda: K.gffm_softmax(E))
    print(f"HW={HW} ci={ci}: attn gram {t_g:.0f} us ({S.shape[0]} chunks) + gfe_weff {t_c:.0f} us | gffm gram {t_g2:.0f} us ({E.shape[0]} chunks) + softmax {t_c2:.0f} us")
    tot += 2 * (t_g + t_c) + t_g2 + t_c2
print(f"per step (2 x attn + 1 x gffm per level): {tot / 1e3:.2f} ms")
