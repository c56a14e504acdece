#!/usr/bin/env python
"""Timing of the training-slice kernels on cuda:0 at the three stage shapes of BASELINE cfg4
(DTU training, 640x512, batch 4, D = 48/32/8, C = 32/16/8): homo_warping_3D forward and backward (op-level A1
contract, fp32 NCHW / NCDHW), depth_regression backward, one stage of final_loss forward + backward.  Algorithmic bytes:
warp = C*V*4 (volume written / gradient volume read) + 4*V (hypotheses) + C*P*4 (features / feature gradient);
regress backward = 4*(P + 2*V); loss = 4*4*P forward, 4*5*P backward.  `--cpu` adds the oracle's backward on the host."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cds_mvsnet_b200 as C  # noqa: E402
from cds_mvsnet_b200 import losses, synthetic  # noqa: E402
from oracle import oracle as O  # noqa: E402  (only for the camera composition helper and the --cpu leg)

HBM = 6551.4
try:
    HBM = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


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


rows = []
B = 4
s = synthetic.make_sample(dict(W=640, H=512, N=3, ndepths=(48, 32, 8), ratios=(4.0, 1.5, 0.75), B=B, Dtot=192, interval=2.65))
for stage, (Cc, D, scale) in enumerate(((32, 48, 4), (16, 32, 2), (8, 8, 1))):
    h, w = 512 // scale, 640 // scale
    pm = s.proj_matrices[f"stage{stage + 1}"]
    refP, srcP = O.compose_projection(pm[:, 0]).cuda(), O.compose_projection(pm[:, 1]).cuda()
    torch.manual_seed(stage)
    dv = (500 + 300 * torch.rand(B, D, h, w)).cuda()
    x = torch.randn(B, Cc, h, w, device="cuda", requires_grad=True)
    g = torch.randn(B, Cc, D, h, w, device="cuda")
    P, V = B * h * w, B * D * h * w
    wbytes = Cc * V * 4 + 4 * V + Cc * P * 4
    with torch.no_grad():
        fwd_ms = timed(lambda: C.homo_warping_3D(x, srcP, refP, dv))
    out = C.homo_warping_3D(x, srcP, refP, dv)

    def bwd():
        x.grad = None
        out.backward(g, retain_graph=True)
    bwd_ms = timed(bwd)
    row = {"stage": stage + 1, "shape": f"B{B} C{Cc} D{D} {h}x{w}", "warp_fwd_ms": fwd_ms, "warp_bwd_ms": bwd_ms, "warp_algorithmic_mb": wbytes / 1e6,
           "warp_fwd_gbs": wbytes / fwd_ms / 1e6, "warp_bwd_gbs": wbytes / bwd_ms / 1e6, "warp_bwd_frac_of_hbm": wbytes / bwd_ms / 1e6 / HBM}
    p = torch.softmax(torch.randn(B, D, h, w, device="cuda"), 1).requires_grad_(True)
    gd = torch.randn(B, h, w, device="cuda")
    dep = C.depth_regression(p, dv)

    def rb():
        p.grad = None
        dep.backward(gd, retain_graph=True)
    rb_ms = timed(rb)
    row.update(regress_bwd_ms=rb_ms, regress_bwd_gbs=4 * (P + 2 * V) / rb_ms / 1e6)
    est = (dv[:, 0] + torch.randn(B, h, w, device="cuda") * 4).requires_grad_(True)
    curv = torch.rand(B, 1, h, w, device="cuda", requires_grad=True)
    gt, mask, iv = dv[:, 0].contiguous(), (torch.rand(B, h, w, device="cuda") > 0.3).float(), torch.full((B,), 2.65, device="cuda")

    def loss_step():
        est.grad = curv.grad = None
        dl, cm = losses._StageLossFn.apply(est, curv, gt, mask, iv)
        (dl + 0.1 * cm).backward()
    row.update(stage_loss_fwd_bwd_ms=timed(loss_step), stage_loss_algorithmic_mb=4 * 9 * P / 1e6)
    if "--cpu" in sys.argv:
        gc, sc, rc, dc = g.cpu(), srcP.cpu(), refP.cpu(), dv.cpu()
        t0 = time.perf_counter()
        O.homo_warp_backward(gc, sc, rc, dc)
        row.update(cpu_oracle_warp_bwd_ms=(time.perf_counter() - t0) * 1e3, cpu_threads=torch.get_num_threads())
    rows.append(row)
print(json.dumps({"op": "training-slice kernels at cfg4 stage shapes", "hbm_peak_gbs": HBM, "stages": rows}))
