import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mmsam_b200  # noqa
from mmsam_b200 import kernels as K
nh, Bp, Kh, Kw = 16, 200, 14, 14
T = Kh * Kw
qkv = torch.randn(Bp, T, 3 * nh * 64, device="cuda").to(torch.bfloat16)
th = K.relpos_table(torch.randn(27, 64, device="cuda") * 0.2, Kh)
tw = K.relpos_table(torch.randn(27, 64, device="cuda") * 0.2, Kw)
out = K.attention(qkv, nh, (Kh, Kw), th, tw)
for _ in range(3):
    K.attention(qkv, nh, (Kh, Kw), th, tw, out=out)
torch.cuda.synchronize()
