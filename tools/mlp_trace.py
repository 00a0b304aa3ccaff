"""clock64 timeline of CTA pair 0 of the fused ConvNeXt MLP kernel (MMSAM_MLP_TRACE=1): python tools/mlp_trace.py [C]"""
import ctypes, os, sys
os.environ["MMSAM_MLP_TRACE"] = "1"
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mmsam_b200  # noqa
from mmsam_b200 import kernels as K, _lib
C = int(sys.argv[1]) if len(sys.argv) > 1 else 384
M = 32768 * 384 // C
y = torch.randn(M, C, device="cuda").to(torch.bfloat16)
w1 = (torch.randn(4 * C, C, device="cuda") / C ** 0.5).to(torch.bfloat16)
w2 = (torch.randn(C, 4 * C, device="cuda") / (4 * C) ** 0.5).to(torch.bfloat16)
cs, b1, b2, gm = w1.float().sum(1), torch.randn(4 * C, device="cuda"), torch.randn(C, device="cuda"), torch.rand(C, device="cuda")
t = torch.randn(M, C, device="cuda")
for _ in range(3):
    K.convnext_mlp(y, w1, cs, b1, w2, b2, gm, t, 1e-6)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * (3 * 64 * 12))()
lib = ctypes.CDLL(_lib.LIB_PATH)
assert lib.mmsam_dbg_mlp_trace(buf) == 1
T = [[[buf[(r * 64 + g) * 12 + e] for e in range(12)] for g in range(64)] for r in range(3)]
t0 = T[0][0][0]
print("MMA thread (leader): per chunk g: G1 wait-start, w1 ready, G1 issued+committed | G2(g): wait start, h_ready, w2/o ready, issued")
for g in range(4, 30):
    m = T[0][g]
    print(f" g={g:2d} G1: {m[0]-t0:7d} +{m[1]-m[0]:5d} (w1 wait) +{m[2]-m[1]:4d} (issue) | G2: {m[3]-t0:7d} +{m[4]-m[3]:5d} (h_ready wait) +{m[5]-m[4]:5d} (w2 wait) +{m[6]-m[5]:4d} (issue) | period {T[0][g][0]-T[0][g-1][0]:5d}")
for r in (1, 2):
    e0 = T[r][0][0]
    print(f"epilogue warp 0 of CTA rank {r-1}: wait start, +hacc_full wait, +tmem ld, +math, +tmem st, +arrive")
    for g in range(4, 30):
        e = T[r][g]
        print(f" g={g:2d} {e[0]-e0:7d} +{e[1]-e[0]:5d} +{e[2]-e[1]:4d} +{e[3]-e[2]:4d} +{e[4]-e[3]:4d} +{e[5]-e[4]:4d} | period {T[r][g][0]-T[r][g-1][0]:5d}")
NCH = 4 * C // 128
print("tile boundaries (leader CTA, warp 0): last chunk arrive -> next A wait start -> A normalised / a_ready arrive -> o_full wait start -> o_full -> drain done | MMA: a_ready wait start -> seen")
for tb in range(1, 4):
    g = tb * NCH
    if g >= 64 or T[1][g][6] == 0:
        break
    e, l, m = T[1][g], T[1][g - 1], T[0][g]
    print(f" tile {tb}: {l[5]-t0} -> {e[6]-t0} -> {e[8]-t0} -> {l[9]-t0} -> {l[10]-t0} -> {l[11]-t0} | MMA {m[7]-t0} -> {m[8]-t0}; first G1 issued {m[2]-t0}")
# same-SM correlation (leader CTA): G1 commit of chunk g -> epilogue wake
print("leader: G1(g) committed -> epilogue warp 0 sees hacc_full; epilogue arrive -> MMA thread sees h_ready")
for g in range(4, 16):
    print(f" g={g:2d} commit->wake {T[1][g][1]-T[0][g][2]:6d}   arrive->h_ready seen {T[0][g][4]-T[1][g][5]:6d}")
