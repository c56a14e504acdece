#!/usr/bin/env python
"""Per-kernel CUDA-event times of this repository's kernels inside one patched training step (640x512, N=3)."""
import os, sys, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cds_mvsnet_b200 as C
from cds_mvsnet_b200 import losses, synthetic, _lib
from oracle import ref_live
cfg = dict(W=640, H=512, N=3, ndepths=(48, 32, 8), ratios=(4.0, 1.5, 0.75), B=1, Dtot=192, interval=2.65)
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
C.patch(rmodel, rmodule, level="leaf")
model = ref_live.build_model(sd, cfg["ndepths"], cfg["ratios"], device=dev, rmodel=rmodel).train()
interval = torch.tensor([cfg["interval"]], device=dev)
def step():
    model.zero_grad(set_to_none=True)
    out = model(imgs, proj, dv, gt_depths=gts, temperature=0.01)
    total, _ = losses.final_loss(out, gts, masks, dlossw=[0.5, 1.0, 2.0], depth_interval=interval)
    total.backward()
step(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with _lib.LaunchProfile() as prof:
    e0.record(); step(); e1.record()
torch.cuda.synchronize()
tab = prof.summary()
rows = sorted(((d["ms_total"], d["launches"], k[0]) for k, d in tab.items()), reverse=True)
agg = {}
for ms, n, name in rows:
    a = agg.setdefault(name, [0.0, 0]); a[0] += ms; a[1] += n
print("step", e0.elapsed_time(e1), "ms; our kernels", sum(r[0] for r in rows), "ms")
for name, (ms, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"  {ms:8.3f} ms  {n:4d} launches  {name}")
