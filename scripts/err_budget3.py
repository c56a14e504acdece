"""Diagnostic (GPU box): end-to-end 'noise' depth error (rel-L1 per stage, mean over seeds) under kernel toggles."""
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
KEYS = ("CDS_USE_TC", "CDS_SPLIT", "CDS_TC_CONV2D", "CDS_TC_GATHER", "CDS_TC_ROLL", "CDS_TC_DYN", "CDS_TC_CONV3D", "CDS_TC_VIS")
variants = {"no tc": {"CDS_USE_TC": "0"}, "all tc": {}, "tc - conv2d": {"CDS_TC_CONV2D": "0"}, "tc - gather": {"CDS_TC_GATHER": "0"},
            "tc - roll": {"CDS_TC_ROLL": "0"}, "tc + split": {"CDS_SPLIT": "1", "CDS_TC_CONV2D": "0"}, "tc - dyn": {"CDS_TC_DYN": "0"}}
seeds = [int(a) for a in sys.argv[1:]] or list(range(6))
acc = {k: [] for k in variants}
for seed in seeds:
    s = synthetic.make_sample(cfg, "noise", seed=seed)
    ref = O.cdsmvsnet_forward(sd, s.imgs, s.proj_matrices, s.depth_values, cfg["ndepths"], cfg["ratios"], 0.01)
    for name, env in variants.items():
        for k in KEYS: os.environ.pop(k, None)
        os.environ.update(env)
        m = C.CDSMVSNet(ndepths=cfg["ndepths"], depth_interals_ratio=cfg["ratios"]); m.load_state_dict(sd); m = m.cuda().eval()
        out = m(s.imgs.cuda(), {k: v.cuda() for k, v in s.proj_matrices.items()}, s.depth_values.cuda(), temperature=0.01)
        acc[name].append([O.rel_l1(out[f"stage{i}"]["depth"].cpu(), ref[f"stage{i}"]["depth"]) for i in (1, 2, 3)])
    print(seed, {k: ["%.2e" % v for v in acc[k][-1]] for k in acc}, flush=True)
for name, rows in acc.items():
    r = np.array(rows)
    print(f"{name:14s} mean rel-L1 per stage {r.mean(0)}  max {r.max(0)}")
