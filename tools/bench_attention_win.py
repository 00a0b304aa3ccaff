"""SAM window attention at the step's shape (batch 8, 64 x 64 tokens -> 200 windows x 16 heads x 196^2): row-mapped store
(mmsam_attention_bf16 + out_map) vs the fused un-partition store (mmsam_attention_window_bf16)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mmsam_b200  # noqa
from mmsam_b200 import kernels as K
from mmsam_b200.engine import window_maps
B, H, W, nh = 8, 64, 64, 16
sc = window_maps(B, H, W, 14, 8, "cuda")
qkv = torch.randn(sc["win_bp"], 196, 3 * nh * 64, device="cuda").to(torch.bfloat16)
th = K.relpos_table(torch.randn(27, 64, device="cuda") * 0.2, 14)
tw = K.relpos_table(torch.randn(27, 64, device="cuda") * 0.2, 14)
out = torch.empty(B * H * W, nh * 64, dtype=torch.bfloat16, device="cuda")


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n * 1e3


print(f"row-mapped store:      {timed(lambda: K.attention(qkv, nh, (14, 14), th, tw, out=out, out_map=sc['win_inv'], out_rows=B * H * W)):.1f} us")
print(f"un-partitioned store:  {timed(lambda: K.attention_window(qkv, nh, B, H, W, th, tw, out=out)):.1f} us")
