"""Attention kernel alone (for ncu --set full): global ViT-L shape with B images, and the window shape."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mmsam_b200  # noqa
from mmsam_b200 import kernels as K

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
nh = 16
for (Bp, Kh, Kw) in ((B, 64, 64), (B * 25, 14, 14)):
    T = Kh * Kw
    qkv = torch.randn(Bp, T, 3 * nh * 64, device="cuda").to(torch.bfloat16)
    th = K.relpos_table(torch.randn(2 * Kh - 1, 64, device="cuda") * 0.2, Kh)
    tw = K.relpos_table(torch.randn(2 * Kw - 1, 64, device="cuda") * 0.2, Kw)
    out = K.attention(qkv, nh, (Kh, Kw), th, tw)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(3):
        K.attention(qkv, nh, (Kh, Kw), th, tw, out=out)
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 3
    fl = 4.0 * Bp * nh * T * T * 64
    print(f"attention Bp={Bp} T={T}: {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s (QK^T+PV only)")
