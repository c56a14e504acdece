#!/usr/bin/env python
"""Run a few forwards of one workload on cuda:0 (the target of ncu captures; prints nothing that is a bench value).

    python scripts/run_forward.py [--workload cfg2] [--iters 2]
"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import cds_mvsnet_b200 as C  # noqa: E402
from cds_mvsnet_b200 import synthetic  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="cfg2")
ap.add_argument("--iters", type=int, default=2)
args = ap.parse_args()
torch.set_grad_enabled(False)
cfg = dict(synthetic.CONFIGS[args.workload])
z = np.load(os.path.join(ROOT, "tests", "golden", "weights_both_dtu_blended.npz"))
model = C.CDSMVSNet(refine=False, ndepths=cfg["ndepths"], depth_interals_ratio=cfg["ratios"])
model.load_state_dict({k: torch.from_numpy(z[k]) for k in z.files})
model = model.cuda().eval()
s = synthetic.make_sample(cfg, "plane", seed=0)
imgs, dv = (s.imgs.clamp(0, 1) * 255.0).round().to(torch.uint8).cuda(), s.depth_values.cuda()   # 8-bit images, as bench.py holds them
proj = {k: v.cuda() for k, v in s.proj_matrices.items()}
eng = model.engine(torch.device("cuda", 0))
for _ in range(args.iters):
    eng.forward(imgs, proj, dv, 0.01)
torch.cuda.synchronize()
print("done")
