import sys, torch
sys.path.insert(0, "/root/repo")
import mmsam_b200
from mmsam_b200 import kernels as K
lg = torch.randn(8 * 256 * 256, 32, device="cuda")
for _ in range(3): K.upsample_argmax(lg, 8, (256, 256), 25, (1024, 1024), None)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(10): K.upsample_argmax(lg, 8, (256, 256), 25, (1024, 1024), None)
e.record(); torch.cuda.synchronize()
print(f"upsample_argmax 8 x 256^2 x 25 -> 1024^2: {s.elapsed_time(e) * 100:.0f} us")
