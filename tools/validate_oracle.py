"""Build-container check: oracle/model.py vs the UNMODIFIED reference modules (via tools/ref_shim.py)."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import ref_shim  # noqa: E402
from oracle import model as om  # noqa: E402
from oracle.perturb import perturb_state_dict, synthetic_batch  # noqa: E402

TINY = dict(img_size=128, modalities_name=["rgb", "lidar"], modalities_ch=[3, 3], init_values=1e-6,
            gamma_init_values=1e-6, patch_size=16, embed_dim=64, depth=4, num_heads=2, mlp_ratio=4,
            drop_path_rate=0.3, drop_multimodal_path=0, conv_inplane=16, n_points=4, deform_num_heads=2,
            cffn_ratio=0.25, deform_ratio=0.5, with_cp=True, interaction_indexes=[[0, 0], [1, 1], [2, 2], [3, 3]],
            global_attn_indexes=[1, 3], window_size=14, arch=dict(depths=[1, 1, 2, 1], channels=[32, 64, 128, 256]),
            checkpoint="none", pretrained_size=256)

VITB = dict(img_size=512, modalities_name=["rgb", "lidar"], modalities_ch=[3, 3], init_values=1e-6,
            gamma_init_values=1e-6, patch_size=16, embed_dim=768, depth=12, num_heads=12, mlp_ratio=4,
            drop_path_rate=0.3, drop_multimodal_path=0, conv_inplane=48, n_points=4, deform_num_heads=12,
            cffn_ratio=0.25, deform_ratio=0.5, with_cp=True, interaction_indexes=[[0, 2], [3, 5], [6, 8], [9, 11]],
            global_attn_indexes=[2, 5, 8, 11], window_size=14, arch="small", checkpoint="none")


def run(cfg, name, dtype=torch.float32):
    torch.manual_seed(0)
    net = ref_shim.build_backbone(cfg)
    sd = perturb_state_dict(net.state_dict(), seed=1)
    net.load_state_dict(sd)
    net = net.to(dtype)
    sd = {k: (v.to(dtype) if torch.is_floating_point(v) else v) for k, v in sd.items()}
    x = synthetic_batch(1, cfg["img_size"]).to(dtype)
    with torch.no_grad():
        t0 = time.time(); ref, _ = net(x); t1 = time.time()
        out = om.backbone_forward(sd, cfg, x); t2 = time.time()
    print(f"{name}: reference {t1 - t0:.2f}s oracle {t2 - t1:.2f}s")
    for i, (a, b) in enumerate(zip(ref, out)):
        rel = ((a - b).norm() / a.norm()).item()
        print(f"  f{i + 1} {tuple(a.shape)} rel-L2 {rel:.3e} max-abs {((a - b).abs().max()).item():.3e} |ref| {a.abs().mean().item():.3f}")


if __name__ == "__main__":
    run(TINY, "tiny fp64", torch.float64)
    run(TINY, "tiny fp32")
    if "--vitb" in sys.argv:
        run(VITB, "ViT-B 512 fp32")
