#!/usr/bin/env python
"""One training step (forward with gt_depths + final_loss + backward) of the reference's CDSMVSNet on the GPU: unpatched
(cuDNN / ATen, TF32 off and on) against patch(level="leaf") (this repository's training kernels for DynamicConv and CostRegNet,
differentiable warp / regression, fused loss).  Correctness of the patched step is tests/test_gpu_train.py; this is its cost.

    python scripts/bench_train_step.py [--hw 512x640] [--n 3] [--steps 5]
"""
import argparse, json, os, sys
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cds_mvsnet_b200 as C  # noqa: E402
from cds_mvsnet_b200 import losses, synthetic  # noqa: E402
from oracle import ref_live  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--hw", default="512x640")
ap.add_argument("--n", type=int, default=3)
ap.add_argument("--steps", type=int, default=5)
args = ap.parse_args()
H, Wd = (int(v) for v in args.hw.split("x"))
cfg = dict(W=Wd, H=H, N=args.n, ndepths=(48, 32, 8), ratios=(4.0, 1.5, 0.75), B=1, Dtot=192, interval=2.65)
z = np.load(os.path.join(ROOT, "tests", "golden", "weights_both_dtu_blended.npz"))
sd = {k: torch.from_numpy(z[k]) for k in z.files}
s = synthetic.make_sample(cfg, "plane", seed=0)
dev = "cuda"
imgs, dv = s.imgs.to(dev), s.depth_values.to(dev)
proj = {k: v.to(dev) for k, v in s.proj_matrices.items()}
gt = s.gt_depth.to(dev)
gts = {"stage1": gt[:, ::4, ::4].contiguous(), "stage2": gt[:, ::2, ::2].contiguous(), "stage3": gt, "stage4": gt}
masks = {k: torch.ones_like(v) for k, v in gts.items()}
rmodel, rmodule, _, _ = ref_live.load()
rlosses = __import__("models.losses", fromlist=["final_loss"])


def run(model, loss_fn, steps):
    model.train()
    interval = torch.tensor([cfg["interval"]], device=dev)
    def step():
        model.zero_grad(set_to_none=True)
        out = model(imgs, proj, dv, gt_depths=gts, temperature=0.01)
        total, _ = loss_fn(out, gts, masks, dlossw=[0.5, 1.0, 2.0], depth_interval=interval)
        total.backward()
        return float(total)
    step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.reset_peak_memory_stats()
    e0.record()
    for _ in range(steps):
        last = step()
    e1.record()
    torch.cuda.synchronize()
    return {"ms_per_step": e0.elapsed_time(e1) / steps, "loss": last, "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9}


res = {"config": f"{H}x{Wd} N={args.n} D=48/32/8 B=1, one training step = forward(gt_depths) + final_loss + backward"}
for tf32 in (False, True):
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = tf32
    torch.backends.cudnn.benchmark = True
    m = ref_live.build_model(sd, cfg["ndepths"], cfg["ratios"], device=dev, rmodel=rmodel)
    res["reference_tf32" if tf32 else "reference_fp32"] = run(m, rlosses.final_loss, args.steps)
    del m
torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
saved = C.patch(rmodel, rmodule, level="leaf")
try:
    m = ref_live.build_model(sd, cfg["ndepths"], cfg["ratios"], device=dev, rmodel=rmodel)
    res["patched_leaf"] = run(m, losses.final_loss, args.steps)
finally:
    C.unpatch(saved)
print(json.dumps(res, indent=1))
