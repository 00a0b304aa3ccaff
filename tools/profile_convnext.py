"""ConvNeXt block kernels alone (for ncu --set full): the 7x7 depthwise conv on the fp32 stream and the fused MLP tail at the
stage-2 shape of the bench (batch 8, 64 x 64 map, C = 384) and the stage-0 shape (256 x 256, C = 96)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mmsam_b200  # noqa
from mmsam_b200 import kernels as K

B = 8
for C, hw in ((384, 64), (96, 256)):
    M = B * hw * hw
    t = torch.randn(M, C, device="cuda")
    w = torch.randn(49, C, device="cuda") / 7
    b = torch.randn(C, device="cuda")
    y = K.dwconv(t, w, b, 7, [(hw, hw)], B, C, hw * hw * C, hw * hw * C)
    w1 = (torch.randn(4 * C, C, device="cuda") / C ** 0.5).to(torch.bfloat16)
    w2 = (torch.randn(C, 4 * C, device="cuda") / (4 * C) ** 0.5).to(torch.bfloat16)
    cs, b1, b2, gm = w1.float().sum(1), torch.randn(4 * C, device="cuda"), torch.randn(C, device="cuda"), torch.rand(C, device="cuda")
    for _ in range(2):
        y = K.dwconv(t, w, b, 7, [(hw, hw)], B, C, hw * hw * C, hw * hw * C)
        K.convnext_mlp(y.view(M, C), w1, cs, b1, w2, b2, gm, t, 1e-6)
    torch.cuda.synchronize()
print("done")
