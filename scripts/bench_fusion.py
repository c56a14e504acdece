#!/usr/bin/env python
"""Timing of the geometric-consistency filter on cuda:0 at the reference's protocol size
(1600x1184, 10 source views, test.py:327): fused single pass vs the op-level chain, with the algorithmic bytes of each."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cds_mvsnet_b200 import fusion as F, synthetic  # noqa: E402

torch.set_grad_enabled(False)
H, W, V = 1184, 1600, 10
s = {k: v.cuda() for k, v in synthetic.make_fusion_sample(H, W, V, seed=0).items()}
args = (s["ref_depth"], s["src_depths"], s["ref_cam"], s["src_cams"], 1.0, 0.01, 3)


def timed(fn, iters=20):
    for _ in range(3):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def chain():
    xyd, inr = F.get_reproj(*args[:4])
    masks, mask = F.vis_filter(args[0], xyd, inr, *args[4:])
    return F.ave_fusion(args[0], xyd, masks)


P = H * W
fused_ms = timed(lambda: F.geometric_filter(*args))
chain_ms = timed(chain)
fused_bytes = P * (4 + 4 * V + 4 * V + 1 + 4 + 12)            # ref depth, V source depths, masks, vis_mask, ave, points
line = {"op": "geometric_filter", "shape": f"{W}x{H}, {V} source views", "fused_ms": fused_ms, "op_chain_ms": chain_ms,
        "fused_algorithmic_mb": fused_bytes / 1e6, "fused_gbs": fused_bytes / fused_ms / 1e6}
if "--cpu" in sys.argv:
    from oracle import fusion_oracle as FO
    c = {k: v.cpu() for k, v in s.items()}
    t0 = time.perf_counter()
    xyd, inr = FO.get_reproj(c["ref_depth"], c["src_depths"], c["ref_cam"], c["src_cams"])
    masks, mask = FO.vis_filter(c["ref_depth"], xyd, inr, 1.0, 0.01, 3)
    FO.ave_fusion(c["ref_depth"], xyd, masks)
    line["cpu_oracle_ms"] = (time.perf_counter() - t0) * 1e3
    line["cpu_threads"] = torch.get_num_threads()
print(json.dumps(line))
