"""Diagnostic (GPU box): end-to-end depth error of the fp16 path under different kernel toggles (env vars read at engine build)."""
import os, sys, itertools
import numpy as np, torch
sys.path.insert(0, ".")
import cds_mvsnet_b200 as C
from cds_mvsnet_b200 import synthetic
from oracle import oracle as O
torch.set_grad_enabled(False)
z = np.load("tests/golden/weights_both_dtu_blended.npz"); sd = {k: torch.from_numpy(z[k]) for k in z.files}
g = np.load("tests/golden/e2e_small3_noise.npz")
cfg = dict(W=160, H=128, N=4, ndepths=(48, 32, 8), ratios=(4.0, 1.5, 0.75), B=2, Dtot=192, interval=2.65)
s = synthetic.make_sample(cfg, "noise", seed=0)
def run(env, storage=torch.float16):
    for k in ("CDS_USE_TC", "CDS_TC_DYN", "CDS_TC_CONV3D", "CDS_TC_VIS", "CDS_SPLIT"): os.environ.pop(k, None)
    os.environ.update(env)
    m = C.CDSMVSNet(ndepths=cfg["ndepths"], depth_interals_ratio=cfg["ratios"], storage=storage); m.load_state_dict(sd); m = m.cuda().eval()
    out = m(s.imgs.cuda(), {k: v.cuda() for k, v in s.proj_matrices.items()}, s.depth_values.cuda(), temperature=0.01)
    return " ".join(f"{O.rel_l1(out[f'stage{i}']['depth'].cpu(), torch.from_numpy(g[f'stage{i}_depth'])):.2e}" for i in (1, 2, 3))
print("fp32 storage                 ", run({}, torch.float32))
print("fp16, no tc                  ", run({"CDS_USE_TC": "0"}))
print("fp16, tc dyn only, no split  ", run({"CDS_TC_CONV3D": "0", "CDS_TC_VIS": "0", "CDS_SPLIT": "0"}))
print("fp16, tc dyn only, split     ", run({"CDS_TC_CONV3D": "0", "CDS_TC_VIS": "0"}))
print("fp16, tc conv3d only         ", run({"CDS_TC_DYN": "0", "CDS_TC_VIS": "0"}))
print("fp16, tc vis only            ", run({"CDS_TC_DYN": "0", "CDS_TC_CONV3D": "0"}))
print("fp16, all tc, no split       ", run({"CDS_SPLIT": "0"}))
print("fp16, all tc, split          ", run({}))
print("fp16, all tc, split (repeat) ", run({}))
