"""One forward of the bench workload bracketed by cudaProfilerStart/Stop, for
   ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file ... python tools/profile_step.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from oracle.perturb import synthetic_batch  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
seg, sd = bench.build_model()
seg = seg.cuda()
eng = seg.backbone.engine(seg.decode_head)
x = synthetic_batch(B, 1024).cuda()
eng.segment(x)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
eng.segment(x)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("done")
