"""Diagnostic (GPU box): end-to-end 'noise' depth error (rel-L1 per stage) of the default path and of named env-var variants,
per seed, against the CPU oracle.  usage: python scripts/err_budget4.py [--hw 128x160] [--seeds 0-7] name=ENV1:v,ENV2:v ..."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import cds_mvsnet_b200 as C
from cds_mvsnet_b200 import synthetic
from oracle import oracle as O
torch.set_grad_enabled(False)
O.FAST_GATHER = True
z = np.load("tests/golden/weights_both_dtu_blended.npz"); sd = {k: torch.from_numpy(z[k]) for k in z.files}
H, W, NV, seeds, variants = 128, 160, 4, list(range(8)), {"default": {}}
for a in sys.argv[1:]:
    if a.startswith("--hw="): H, W = (int(v) for v in a[5:].split("x"))
    elif a.startswith("--seeds="): lo, hi = a[8:].split("-"); seeds = list(range(int(lo), int(hi) + 1))
    elif a.startswith("--n="): NV = int(a[4:])
    else:
        name, _, envs = a.partition("=")
        variants[name] = dict(e.split(":") for e in envs.split(",") if e)
cfg = dict(W=W, H=H, N=NV, ndepths=(48, 32, 8), ratios=(4.0, 1.5, 0.75), B=1, Dtot=192, interval=2.65)
KEYS = sorted({k for v in variants.values() for k in v})
acc = {k: [] for k in variants}
for seed in seeds:
    s = synthetic.make_sample(cfg, "noise", seed=seed)
    ref = O.cdsmvsnet_forward(sd, s.imgs, s.proj_matrices, s.depth_values, cfg["ndepths"], cfg["ratios"], 0.01)
    for name, env in variants.items():
        for k in KEYS: os.environ.pop(k, None)
        os.environ.update({k: v for k, v in env.items() if k != "STORAGE"})
        st = getattr(torch, env.get("STORAGE", "float16"))
        m = C.CDSMVSNet(ndepths=cfg["ndepths"], depth_interals_ratio=cfg["ratios"], storage=st); m.load_state_dict(sd); m = m.cuda().eval()
        out = m.engine(torch.device("cuda", 0)).forward(s.imgs.cuda(), {k: v.cuda() for k, v in s.proj_matrices.items()}, s.depth_values.cuda(), 0.01)
        acc[name].append([O.rel_l1(out[f"stage{i}"]["depth"].cpu(), ref[f"stage{i}"]["depth"]) for i in (1, 2, 3)])
        del m, out
        torch.cuda.empty_cache()
    print(seed, {k: ["%.2e" % v for v in acc[k][-1]] for k in acc}, flush=True)
for name, rows in acc.items():
    r = np.array(rows)
    print(f"{name:14s} mean rel-L1 per stage {r.mean(0)}  max {r.max(0)}  seeds over 1e-3: {[seeds[i] for i in np.nonzero(r.max(1) >= 1e-3)[0]]}")
