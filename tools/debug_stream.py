"""Debug: stream_labels vs per-batch encode_decode_labels agreement (tiny segmentor). python tools/debug_stream.py"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import build_segmentor, TINY, TINY_HEAD
from oracle.perturb import synthetic_batch
seg, sd = build_segmentor(TINY, TINY_HEAD, test_cfg=dict(mode="whole_dim", rescale=True, dim=(128, 128)))
seg = seg.cuda()
batches = [synthetic_batch(2, 128, seed=40 + i).pin_memory() for i in range(5)]
for graph in (True, False):
    seg.use_cuda_graph = graph
    want = [seg.encode_decode_labels(b.cuda(), (128, 128)).cpu().clone() for b in batches]
    again = [seg.encode_decode_labels(b.cuda(), (128, 128)).cpu().clone() for b in batches]
    got = [lab.clone() for lab in seg.stream_labels(iter(batches), (128, 128))]
    print("graph", graph, "self", [round((a == w).float().mean().item(), 5) for a, w in zip(again, want)])
    print("graph", graph, "stream", [round((g == w).float().mean().item(), 5) for g, w in zip(got, want)])
    print("graph", graph, "stream vs others", [[round((g == w).float().mean().item(), 3) for w in want] for g in got])
