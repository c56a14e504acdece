"""Diagnostic (GPU box): is the end-to-end 'noise' error tail-dominated?  Several seeds, oracle computed on the box."""
import os, sys
import numpy as np, torch
sys.path.insert(0, ".")
import cds_mvsnet_b200 as C
from cds_mvsnet_b200 import synthetic
from oracle import oracle as O
torch.set_grad_enabled(False)
O.FAST_GATHER = True
z = np.load("tests/golden/weights_both_dtu_blended.npz"); sd = {k: torch.from_numpy(z[k]) for k in z.files}
cfg = dict(W=160, H=128, N=4, ndepths=(48, 32, 8), ratios=(4.0, 1.5, 0.75), B=1, Dtot=192, interval=2.65)
variants = {"no tc": {"CDS_USE_TC": "0"}, "tc, no split": {"CDS_SPLIT": "0"}, "tc, split": {}}
acc = {k: [] for k in variants}
for seed in range(4):
    s = synthetic.make_sample(cfg, "noise", seed=seed)
    ref = O.cdsmvsnet_forward(sd, s.imgs, s.proj_matrices, s.depth_values, cfg["ndepths"], cfg["ratios"], 0.01)["depth"]
    for name, env in variants.items():
        for k in ("CDS_USE_TC", "CDS_SPLIT"): os.environ.pop(k, None)
        os.environ.update(env)
        m = C.CDSMVSNet(ndepths=cfg["ndepths"], depth_interals_ratio=cfg["ratios"]); m.load_state_dict(sd); m = m.cuda().eval()
        d = m(s.imgs.cuda(), {k: v.cuda() for k, v in s.proj_matrices.items()}, s.depth_values.cuda(), temperature=0.01)["depth"].cpu()
        e = (d - ref).abs().flatten()
        acc[name].append((O.rel_l1(d, ref), e.median().item(), e.quantile(0.99).item(), e.max().item(), (e > 2.0).float().mean().item()))
        print(seed, name, ["%.3g" % v for v in acc[name][-1]], flush=True)
for name, rows in acc.items():
    print(f"{name:14s} mean rel-L1 {np.mean([r[0] for r in rows]):.3e}  median|e| {np.mean([r[1] for r in rows]):.4f} mm  p99 {np.mean([r[2] for r in rows]):.3f} mm")
