"""Throughput of BASELINE configs 4 and 5 (parity for both is in tests/test_model_gpu.py at reduced size):
  4  FMB-shaped RGB+thermal, 800x800 network input (600x800 frames zero-padded), ViT-L ...NEWwithcp, 14 classes,
     whole_dim_cut (logits cropped to 600x800): padded 56x56-token windows, interpolated global rel-pos.
  5a MUSES-shaped RGB+LiDAR 1080x1920 frames, 19 classes, slide inference: 6 crops of 1024^2 (stride 640) per frame,
     batched into one forward, logits averaged by overlap count.
python tools/bench_configs.py"""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from common import build_segmentor
from oracle.perturb import synthetic_batch


def timed(fn, n=3):
    fn(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n


# ---- config 4 ----
cfg4 = dict(bench.VITL, img_size=800, modalities_name=["rgb", "thermal"], conv_drop_path_rate=0.3)
head4 = dict(bench.VITL_HEAD, num_classes=14)
seg, _ = build_segmentor(cfg4, head4, test_cfg=dict(mode="whole_dim_cut", rescale=False, dim=(600, 800), cut_dim=(800, 600)),
                         btype="SAMAdapterbimodalMixModNewInTwinConvNEWwithcp")
seg = seg.cuda()
B = 8
x = synthetic_batch(B, 800, kind="thermal").cuda()
x[:, :, 600:] = 0
ms = timed(lambda: seg.encode_decode_labels(x, (800, 800), (600, 800)))
print(f"config 4 (FMB 800x800 -> 600x800 labels, batch {B}): {ms:.1f} ms per batch, {B / ms * 1e3:.1f} img/s")
del seg
torch.cuda.empty_cache()

# ---- config 5a ----
head5 = dict(bench.VITL_HEAD, num_classes=19)
seg, _ = build_segmentor(bench.VITL, head5, test_cfg=dict(mode="slide", crop_size=(1024, 1024), stride=(640, 640)))
seg = seg.cuda()
frames = synthetic_batch(1, (1080, 1920)).cuda()
ms = timed(lambda: seg.slide_labels(frames))
print(f"config 5a (MUSES 1080x1920 slide, 6 crops batched, 1 frame): {ms:.1f} ms per frame, {1e3 / ms:.2f} frames/s, {6e3 / ms:.1f} crops/s")
