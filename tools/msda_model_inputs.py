"""MSDeformAttn on the tensors the ViT-L model really produces (bench.py's model and frames): captures every fused
MSDeformAttn call of one forward, prints the distribution of the sampling offsets around the sampling_offsets-bias prior
(what the staged kernel's boxes are placed by), the share of samples outside their box for a given margin, and the time of
the L1-gather and the staged kernel on those tensors. python tools/msda_model_inputs.py [margin ...]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as BN  # noqa
import mmsam_b200  # noqa
from mmsam_b200 import kernels as K
import mmsam_b200.engine as E

margins = [int(a) for a in sys.argv[1:]] or [2]
seg, sd = BN.build_model()
seg = seg.cuda()
P = BN.PIPELINE
seg.set_input_pipeline(P["mean"], P["std"], to_rgb=P["to_rgb"], norm_by_max=P["norm_by_max"])
eng = seg.backbone.engine(seg.decode_head)
rgb, aux = BN.synthetic_frames(8, 1024, 1234)
frames = seg.u8_input(rgb.cuda(), aux.cuda())
calls = []
orig = K.msda_fused


def hook(value, shapes, lsi, qproj, ref, n_heads, n_levels, n_points=4, out=None, geom=None):
    r = orig(value, shapes, lsi, qproj, ref, n_heads, n_levels, n_points, out, geom=geom)
    calls.append((value.clone(), shapes, lsi, qproj.clone(), ref, n_heads, n_levels, n_points))
    return r


E.K.msda_fused = hook
os.environ["MMSAM_TOWER_STREAMS"] = "0"; os.environ["MMSAM_NECK_STREAMS"] = "0"
eng.segment(frames)
torch.cuda.synchronize()
E.K.msda_fused = orig
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


WARM = os.environ.get("MSDA_WARM", "0") == "1"      # 1: no L2 flush between the timed launches


def timeit(fn):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(7):
        if not WARM:
            flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return sorted(ts)[len(ts) // 2] * 1e3


msdas = [e["attn"] for it in eng.inter for e in it["ext"]] if hasattr(eng, "inter") else None
for i, (value, shapes, lsi, qproj, ref, M, L, Pn) in enumerate(calls):
    N, S, MD = value.shape
    Lq = ref.shape[0]
    off = qproj[:, :M * L * Pn * 2].view(N, Lq, M, L, Pn, 2).float()
    mean = off.mean((0, 1))                     # [M, L, P, 2]: the learnt prior (bias) shows up as the mean
    dev = off - mean
    kind = "ext" if L == 1 else "inj"
    print(f"call {i} {kind}: Lq {Lq} S {S} L {L}: |offset - mean| std {dev.std().item():.3f} px, "
          f"99% {dev.abs().flatten()[::97].quantile(0.99).item():.2f} px, max {dev.abs().max().item():.1f} px; "
          f"mean magnitude {mean.abs().mean().item():.2f} px", flush=True)
    if i in (0, 1, len(calls) - 1):
        lv = [tuple(int(v) for v in r) for r in shapes.tolist()]
        s1 = [(64, 64)]
        s3 = [(128, 128), (64, 64), (32, 32)]
        levels, qgrids = (s3, s1) if kind == "inj" else (s1, s3)
        assert lv == levels, (lv, levels)
        t0 = timeit(lambda: orig(value, shapes, lsi, qproj, ref, M, L, Pn))
        line = f"   L1-gather {t0:.0f} us"
        for mg in margins:
            g = K.MsdaGeometry(levels, qgrids, (64, 64), (8, 8) if kind == "inj" else (8, 16), mean.cpu(), M, L, Pn, margin=mg)
            t1 = timeit(lambda: orig(value, shapes, lsi, qproj, ref, M, L, Pn, geom=g))
            line += f" | staged margin {mg}: {t1:.0f} us" + (" (unsupported)" if g.unsupported else "")
        print(line, flush=True)
