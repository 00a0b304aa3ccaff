import math, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mmsam_b200
from mmsam_b200 import kernels as k
B, H, W, Cin, groups = 1, 8, 16, 64, 32
g = torch.Generator().manual_seed(1)
x = torch.randn(B, H, W, Cin, generator=g).to(torch.bfloat16)
w = (torch.randn(Cin, Cin // groups, 3, 3, generator=g) / 4).to(torch.bfloat16)
wp = k.pack_conv3x3_weight(w.float().cuda(), groups)
out = k.conv3x3(x.reshape(-1, Cin).cuda(), wp, B, H, W, Cin, Cin, groups)
torch.cuda.synchronize()
ref = torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), w.float(), padding=1, groups=groups).permute(0, 2, 3, 1)
print("max err", (out.float().cpu().view(B, H, W, Cin) - ref).abs().max().item())
