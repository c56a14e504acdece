#!/usr/bin/env python
"""Measured errors behind the tolerances of tests/test_gpu_train.py (run on the GPU box; prints one line per check)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cds_mvsnet_b200 as C  # noqa: E402
from cds_mvsnet_b200 import losses  # noqa: E402
from oracle import oracle as O  # noqa: E402

z = np.load(os.path.join(ROOT, "tests", "golden", "train_ops.npz"))
g = {k: torch.from_numpy(z[k]) for k in z.files}
cu = lambda t: t.cuda()
for rep in range(3):
    for tag in ("planes", "pix"):
        fea = cu(g["warp_src_fea"]).requires_grad_(True)
        C.homo_warping_3D(fea, cu(g["warp_src_proj"]), cu(g["warp_ref_proj"]), cu(g[f"warp_depth_{tag}"])).backward(cu(g["warp_grad_out"]))
        want = g[f"warp_grad_src_{tag}"]
        print(f"warp backward [{tag}] rep {rep}: rel-L1 {O.rel_l1(fea.grad.cpu(), want):.3e} (test bound 2e-4)  max-abs {(fea.grad.cpu() - want).abs().max():.3e} (3e-3)")
inputs = {f"stage{i}": {"depth": cu(g[f"loss_in_stage{i}.depth"]).requires_grad_(True), "norm_curv": cu(g[f"loss_in_stage{i}.norm_curv"]).requires_grad_(True),
                        "feat_distance": cu(g[f"lossf_in_stage{i}.feat_distance"]).requires_grad_(True), "feat_target": cu(g[f"lossf_in_stage{i}.feat_target"])}
          for i in (1, 2, 3)}
gts = {f"stage{i}": cu(g[f"loss_gt_stage{i}"]) for i in (1, 2, 3, 4)}
masks = {f"stage{i}": cu(g[f"loss_mask_stage{i}"]) for i in (1, 2, 3, 4)}
total, dl = losses.final_loss(inputs, gts, masks, dlossw=g["loss_dlossw"].tolist(), depth_interval=cu(g["loss_interval"]))
total.backward()
print(f"final_loss total rel err {abs(total.item() - g['lossf_total'].item()) / abs(g['lossf_total'].item()):.3e} (2e-6 rel + 1e-5 abs)")
for i in (1, 2, 3):
    a, b = inputs[f"stage{i}"]["feat_distance"].grad.cpu(), g[f"lossf_grad_stage{i}.feat_distance"]
    print(f"stage{i} feat grad: max rel err {((a - b).abs() / b.abs().clamp_min(1e-12)).max():.3e} (2e-5 rel + 1e-8 abs), |grad| max {b.abs().max():.2e}")
    a, b = inputs[f"stage{i}"]["depth"].grad.cpu(), g[f"lossf_grad_stage{i}.depth"]
    print(f"stage{i} depth grad: max rel err {((a - b).abs() / b.abs().clamp_min(1e-12)).max():.3e} (1e-5 rel + 1e-9 abs), |grad| max {b.abs().max():.2e}")
