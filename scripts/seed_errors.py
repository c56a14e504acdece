"""GPU box: depth rel-L1 per stage of the DEFAULT path against the oracle on noise seeds at 128x160 (N = 4) -- run it under an
environment switch (e.g. CDS_ENTROPY_PACKED=1) to see what a kernel alternative does to the parity margin."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
import cds_mvsnet_b200 as C
from cds_mvsnet_b200 import synthetic
from oracle import oracle as O
torch.set_grad_enabled(False)
O.FAST_GATHER = True
z = np.load("tests/golden/weights_both_dtu_blended.npz"); sd = {k: torch.from_numpy(z[k]) for k in z.files}
cfg = dict(W=160, H=128, N=4, ndepths=(48, 32, 8), ratios=(4.0, 1.5, 0.75), B=1, Dtot=192, interval=2.65)
seeds = [int(a) for a in sys.argv[1:]] or list(range(8))
m = C.CDSMVSNet(ndepths=cfg["ndepths"], depth_interals_ratio=cfg["ratios"]); m.load_state_dict(sd); m = m.cuda().eval()
rows = []
for seed in seeds:
    s = synthetic.make_sample(cfg, "noise", seed=seed)
    ref = O.cdsmvsnet_forward(sd, s.imgs, s.proj_matrices, s.depth_values, cfg["ndepths"], cfg["ratios"], 0.01)
    out = m(s.imgs.cuda(), {k: v.cuda() for k, v in s.proj_matrices.items()}, s.depth_values.cuda(), temperature=0.01)
    rows.append([O.rel_l1(out[f"stage{i}"]["depth"].cpu(), ref[f"stage{i}"]["depth"]) for i in (1, 2, 3)])
    print(seed, ["%.2e" % v for v in rows[-1]], flush=True)
r = np.array(rows)
print("mean", r.mean(0), "max", r.max(0))
