"""clock64 timeline of CTA 0, first item, of the two-tile global attention kernel (MMSAM_ATT_TRACE=1)."""
import ctypes, os, sys
os.environ["MMSAM_ATT_TRACE"] = "1"
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mmsam_b200  # noqa
from mmsam_b200 import kernels as K, _lib
nh, Kh, Kw, Bp = 16, 64, 64, 8
T = Kh * Kw
qkv = torch.randn(Bp, T, 3 * nh * 64, device="cuda").to(torch.bfloat16)
th = K.relpos_table(torch.randn(2 * Kh - 1, 64, device="cuda") * 0.2, Kh)
tw = K.relpos_table(torch.randn(2 * Kw - 1, 64, device="cuda") * 0.2, Kw)
for _ in range(2):
    K.attention(qkv, nh, (Kh, Kw), th, tw)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * (3 * 64 * 8))()
lib = ctypes.CDLL(_lib.LIB_PATH)
assert lib.mmsam_dbg_attn_glb_trace(buf) == 1
E = [[[buf[(r * 64 + j) * 8 + e] for e in range(8)] for j in range(64)] for r in range(3)]
t0 = E[0][0][0]
print("block j | group A: s_full wait start, +wait, +softmax (of which pv_done wait), +arrive | group B: same | MMA: kv wait start, +kv, p_full A seen, s_free A seen / QK_A(j+1) issued, p_full B seen, QK_B(j+1) issued")
for j in range(0, 32):
    a, b, m = E[0][j], E[1][j], E[2][j]
    print(f" j={j:2d} | A {a[0]-t0:7d} +{a[1]-a[0]:5d} +{a[2]-a[1]:5d} ({a[5]-a[4]:4d}) +{a[3]-a[2]:4d} | B {b[0]-t0:7d} +{b[1]-b[0]:5d} +{b[2]-b[1]:5d} ({b[5]-b[4]:4d}) +{b[3]-b[2]:4d} | MMA {m[4]-t0:7d} +{m[5]-m[4]:4d}  pA {m[0]-t0:7d} qkA {m[1]-t0:7d} pB {m[2]-t0:7d} qkB {m[3]-t0:7d}")
