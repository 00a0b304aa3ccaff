"""Print the clock64 timeline of CTA 0 of the window-attention kernel (MMSAM_ATT_TRACE=1)."""
import ctypes, os, sys
os.environ["MMSAM_ATT_TRACE"] = "1"
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mmsam_b200  # noqa
from mmsam_b200 import kernels as K, _lib
nh, Bp, T = 16, 200, 196
bias = len(sys.argv) < 2 or sys.argv[1] != "nobias"
th = K.relpos_table(torch.randn(27, 64, device="cuda") * 0.2, 14) if bias else None
tw = K.relpos_table(torch.randn(27, 64, device="cuda") * 0.2, 14) if bias else None
if os.environ.get("WIN_LEGACY"):        # partitioned input / output (the row-mapped path)
    qkv = torch.randn(Bp, T, 3 * nh * 64, device="cuda").to(torch.bfloat16)
    for _ in range(3):
        out = K.attention(qkv, nh, (14, 14), th, tw)
else:                                   # what the model runs: 64 x 64 token images, window_unpartition fused into the store
    qkv = torch.randn(Bp, T, 3 * nh * 64, device="cuda").to(torch.bfloat16)      # 8 images x 25 windows (64 -> 70 tokens padded)
    for _ in range(3):
        out = K.attention_window(qkv, nh, 8, 64, 64, th, tw)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * (5 * 16 * 8))()
lib = ctypes.CDLL(_lib.LIB_PATH)
assert lib.mmsam_dbg_attention_win_trace(buf) == 1
t = [[[buf[(r * 16 + i) * 8 + e] for e in range(8)] for i in range(16)] for r in range(5)]
t0 = min(v for r in t for i in r for v in i if v > 0)
names = ["wait S", "S ready", "bias done", "pass1 done", "P written", "O ready", "stored"]
for g in (0, 1):
    print(f"softmax group {g} (warp {4 * g}, lane 0): cycles since the first stamp")
    for i in range(2, 10):
        print(f"  item {i}: " + "  ".join(f"{names[e]} {t[g][i][e] - t0:7d}" for e in range(7)))
    for i in range(3, 10):
        d = [t[g][i][e + 1] - t[g][i][e] for e in range(6)]
        print(f"  item {i} deltas: S-wait {d[0]:5d} | bias {d[1]:5d} | pass1 {d[2]:5d} | pass2 {d[3]:5d} | O-wait {d[4]:5d} | O-load {t[g][i][7] - t[g][i][5]:5d} | store {t[g][i][6] - t[g][i][7]:5d} | period {t[g][i][1] - t[g][i - 1][1]:6d}")
print("MMA thread: per item [S_A wait, S_A issue, PV_A wait, PV_A issue, S_B wait, S_B issue, PV_B wait, PV_B issue]")
for i in range(2, 10):
    print(f"  item {i}: " + " ".join(f"{t[2][i][e] - t0:7d}" for e in range(8)))

for g in (0, 1):
    for i in range(3, 6):
        e = t[3 + g][i]
        print(f"store phase group {g} item {i}: O loaded->arrive {e[0] - t[g][i][7]:5d} | inv/orow {e[1] - e[0]:5d} | staging {e[2] - e[1]:5d} | fence {e[3] - e[2]:5d} | bulk {e[4] - e[3]:5d} | commit {e[5] - e[4]:5d}")
