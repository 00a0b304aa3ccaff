"""Every distinct GEMM configuration of one bench forward, timed in isolation (20 back-to-back launches):
count x time, TFLOP/s and compulsory GB/s, sorted by its share of the step. python tools/gemm_shapes.py [batch]"""
import os
import sys
from collections import OrderedDict

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from oracle.perturb import synthetic_batch  # noqa: E402
from mmsam_b200 import kernels as K  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
seg, sd = bench.build_model()
seg = seg.cuda()
eng = seg.backbone.engine(seg.decode_head)
x = synthetic_batch(B, 1024).cuda()
eng.segment(x)
torch.cuda.synchronize()

calls = OrderedDict()
og = K.gemm


def rec(a, w, bias=None, act=None, scale=None, residual=None, out=None, out_dtype=torch.bfloat16, row_map=None,
        out_rows=None, pixel_shuffle=None, block_n=0, max_ctas=0):
    r = og(a, w, bias=bias, act=act, scale=scale, residual=residual, out=out, out_dtype=out_dtype, row_map=row_map,
           out_rows=out_rows, pixel_shuffle=pixel_shuffle, block_n=block_n, max_ctas=max_ctas)
    key = (a.shape[0], w.shape[0], a.shape[1], a.stride(0), r.stride(0), act, bias is not None, scale is not None,
           residual is not None, "map" if row_map is not None else ("ps" if pixel_shuffle is not None else "id"),
           str(r.dtype).replace("torch.", ""), block_n)
    if key not in calls:
        calls[key] = [0, (a, w, bias, act, scale, residual, r, row_map, pixel_shuffle, block_n, max_ctas)]
    calls[key][0] += 1
    return r


K.gemm = rec
import mmsam_b200.engine as E  # noqa: E402
import mmsam_b200.neck as NK  # noqa: E402
E.K.gemm = rec
NK.K.gemm = rec
eng.segment(x)
torch.cuda.synchronize()
K.gemm = og

rows = []
for key, (cnt, (a, w, bias, act, scale, residual, out, row_map, ps, bn, mc)) in calls.items():
    def run():
        og(a, w, bias=bias, act=act, scale=scale, residual=residual, out=out, row_map=row_map, pixel_shuffle=ps,
           block_n=bn, max_ctas=mc)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(20):
        run()
    e.record()
    torch.cuda.synchronize()
    us = s.elapsed_time(e) / 20 * 1e3
    M, N, Kd = key[0], key[1], key[2]
    by = M * Kd * 2 + N * Kd * 2 + M * N * (4 if key[10] == "float32" else 2) * (2 if key[8] else 1)
    rows.append((cnt * us, cnt, us, 2.0 * M * N * Kd / us / 1e6, by / us / 1e3, key))
tot = sum(r[0] for r in rows)
print(f"total GEMM time {tot / 1e3:.2f} ms over {sum(r[1] for r in rows)} launches, {len(rows)} distinct configs")
print("  share   tot_us  cnt     us  TFLOP/s   GB/s  (M, N, K, lda, ldo, act, bias, scale, res, rows, out, bn)")
for t, cnt, us, tf, gbs, key in sorted(rows, key=lambda r: -r[0]):
    print(f"  {t / tot * 100:5.1f}% {t:8.0f} {cnt:4d} {us:6.1f} {tf:8.0f} {gbs:6.0f}  {key}")
