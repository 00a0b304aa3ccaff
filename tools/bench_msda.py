"""MSDeformAttn fused kernel at the bench shapes (batch 8, ViT-L 1024^2): injector (Lq=4096, 3 levels) and
extractor (Lq=21504, 1 level). Algorithmic bytes per SURVEY.md 8(d). python tools/bench_msda.py"""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mmsam_b200  # noqa
from mmsam_b200 import kernels as K
PEAK = 6451.2
TILE = tuple(int(v) for v in os.environ["MSDA_TILE"].split("x")) if os.environ.get("MSDA_TILE") else None
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def run(name, N, qgrids, vshapes, M=16, D=32, P=4, noise=0.32, staged=True, margin=2):
    L = len(vshapes)
    sh = torch.as_tensor(vshapes, dtype=torch.long)
    S = int(sh.prod(1).sum())
    lsi = torch.cat((sh.new_zeros((1,)), sh.prod(1).cumsum(0)[:-1]))
    refs = []
    for (h, w) in qgrids:
        ys, xs = torch.meshgrid((torch.arange(h) + 0.5) / h, (torch.arange(w) + 0.5) / w, indexing="ij")
        refs.append(torch.stack([xs.reshape(-1), ys.reshape(-1)], -1))
    ref = torch.cat(refs).float().cuda()
    Lq = ref.shape[0]
    value = torch.randn(N, S, M * D, device="cuda").to(torch.bfloat16)
    # MSDeformAttn._reset_parameters-like offsets: direction of the head x (p + 1) pixels, plus noise
    th = torch.arange(M).float() * (2 * torch.pi / M)
    gi = torch.stack([th.cos(), th.sin()], -1)
    gi = gi / gi.abs().max(-1, keepdim=True)[0]
    bias = (gi.view(M, 1, 1, 2) * torch.arange(1, P + 1).view(1, 1, P, 1)).expand(M, L, P, 2).contiguous()
    off = bias.reshape(-1)
    anchor = qgrids[0] if len(qgrids) == 1 else vshapes[0]
    geom = K.MsdaGeometry(vshapes, qgrids, anchor, TILE or ((8, 8) if len(qgrids) == 1 else (8, 16)), bias, M, L, P, margin=margin) if staged else None
    qproj = torch.randn(N * Lq, M * L * P * 3, device="cuda")
    qproj[:, :M * L * P * 2] = qproj[:, :M * L * P * 2] * noise + off.cuda()
    out = torch.empty(N, Lq, M * D, device="cuda", dtype=torch.bfloat16)
    shc, lsic = sh.cuda(), lsi.cuda()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        K.msda_fused(value, shc, lsic, qproj, ref, M, L, P, out, geom=geom)
    ts = []
    for _ in range(10):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); K.msda_fused(value, shc, lsic, qproj, ref, M, L, P, out, geom=geom); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    ms = ts[len(ts) // 2]
    by = N * (S * M * D * 2 + Lq * M * L * P * 3 * 4 + Lq * M * D * 2)
    units = N * Lq * M * L * P * 4
    name = f"{name} {'staged' if staged and not geom.unsupported else 'L1-gather'} noise {noise}"
    print(f"msda {name}: {ms * 1e3:.0f} us, {by / 1e6:.0f} MB algorithmic -> {by / ms / 1e6:.0f} GB/s ({by / ms / 1e6 / PEAK * 100:.1f}% of measured HBM peak); "
          f"{units / ms / 1e6:.0f} G gathered 64-B units/s")


import itertools
for staged, noise in itertools.product((False, True), (0.32, 2.0)):
    if not staged and noise != 0.32:
        continue
    run("injector (Lq 4096, L3)", 8, [(64, 64)], [(128, 128), (64, 64), (32, 32)], noise=noise, staged=staged)
    run("extractor (Lq 21504, L1)", 8, [(128, 128), (64, 64), (32, 32)], [(64, 64)], noise=noise, staged=staged)
