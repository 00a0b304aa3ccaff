"""Gram kernel at the neck's shapes (batch 8): AttentionBase (qkv [HW, 3ci], per-head blocks, norms) and GFFM ([HW, 2ci])."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mmsam_b200  # noqa
from mmsam_b200 import kernels as K
B = 8
for (HW, ci) in ((65536, 96), (16384, 192), (4096, 384), (1024, 768)):
    for (name, ld, koff, blk, norms) in (("attn", 3 * ci, ci, ci // 8, True), ("gffm", 2 * ci, ci, 0, False)):
        x = torch.randn(B * HW, ld, device="cuda").to(torch.bfloat16)
        for _ in range(2):
            K.gram(x, ld, 0, koff, ci, B, HW, blk=blk, norms=norms)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(5):
            K.gram(x, ld, 0, koff, ci, B, HW, blk=blk, norms=norms)
        e.record(); torch.cuda.synchronize()
        ms = s.elapsed_time(e) / 5
        by = B * HW * 2 * ci * 2
        print(f"gram {name} HW={HW} ci={ci}: {ms * 1e3:.0f} us (incl. the zero-fill of S), {by / ms / 1e6:.0f} GB/s of the 2ci channels read")
