"""Debug aid (GPU box): compare every intermediate of the fused cascade with the CPU oracle."""
import ctypes
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import cds_mvsnet_b200 as C
from cds_mvsnet_b200 import synthetic
from cds_mvsnet_b200._lib import call, ptr
from oracle import oracle as O

torch.set_grad_enabled(False)
z = np.load("tests/golden/weights_both_dtu_blended.npz")
sd = {k: torch.from_numpy(z[k]) for k in z.files}
g = np.load("tests/golden/epipole.npz")
pm = torch.from_numpy(np.stack((g["cam_ref"], g["cam_src"]), 1)).cuda().contiguous()
B = pm.shape[0]
coef = torch.zeros(1, B, 1, 12, device="cuda")
epi = torch.full((2, 1, B, 2), -7.0, device="cuda")
call("cds_camera_setup", (ctypes.c_void_p * 1)(pm.data_ptr()), 1, 0, B, 2, ptr(coef), ptr(epi))
torch.cuda.synchronize()
print("epi gpu", epi.flatten().tolist())
print("epi ref", g["e_ref"].tolist(), g["e_src"].tolist())

storage = torch.float32 if len(sys.argv) < 2 else getattr(torch, sys.argv[1])
cfg = dict(W=160, H=128, N=3, ndepths=(48, 32, 8), ratios=(4.0, 1.5, 0.75), B=1, Dtot=192, interval=2.65)
s = synthetic.make_sample(cfg, "plane", seed=0)
ref = O.cdsmvsnet_forward(sd, s.imgs, s.proj_matrices, s.depth_values, cfg["ndepths"], cfg["ratios"], 0.01, return_intermediates=True)
m = C.CDSMVSNet(ndepths=cfg["ndepths"], depth_interals_ratio=cfg["ratios"], storage=storage)
m.load_state_dict(sd)
m = m.cuda().eval()
out = m(s.imgs.cuda(), {k: v.cuda() for k, v in s.proj_matrices.items()}, s.depth_values.cuda(), temperature=0.01)
eng = m.engine(torch.device("cuda", 0))
buf = eng.buf._t
V, Bn = cfg["N"] - 1, cfg["B"]
print("epipoles engine", buf["cam.epi"].flatten().tolist())
cams3 = torch.unbind(s.proj_matrices["stage3"], 1)
for i in range(1, cfg["N"]):
    Fm = O.fundamental_matrix(cams3[0], cams3[i])
    print("  oracle pair", i, O.epipole_from_F(Fm).tolist(), O.epipole_from_F(Fm.transpose(1, 2)).tolist())


def rel(a, b):
    return O.rel_l1(a.float().cpu(), b)


for st, fname in enumerate(("f.fea1", "f.fea2", "f.fea3")):
    fea = buf[fname]  # [2*V*B, h, w, C]
    for v in range(V):
        for side, key in enumerate(("ref", "src")):
            r = ref["_features"][v][key][f"stage{st + 1}"]
            idx = (side * V + v) * Bn
            print(f"stage{st+1} pair{v} {key}: fea {rel(fea[idx:idx+Bn].permute(0,3,1,2), r[0]):.2e} "
                  f"ncsq {rel(buf[f'f.ncsq{st}'][idx:idx+Bn].unsqueeze(1), r[1]):.2e} ncabs {rel(buf[f'f.ncabs{st}'][idx:idx+Bn].unsqueeze(1), r[2]):.2e}")
for st in range(3):
    it = ref[f"stage{st + 1}"]["_inter"]
    print(f"stage{st+1}: samples {rel(buf[f's{st}.samples'], it['depth_samples']):.2e}")
    for v in range(V):
        print(f"   view{v}: entropy {rel(buf[f's{st}.entropy'][v], it['entropy'][v][:, 0]):.2e} vis {rel(buf[f's{st}.vis'][v], it['vis'][v][:, 0]):.2e}")
    print(f"   volume {rel(buf[f's{st}.volume'].permute(0, 1, 5, 2, 3, 4).reshape(it['volume'].shape), it['volume']):.2e} logits {rel(buf[f's{st}.cr.logits'], it['logits']):.2e} "
          f"depth {rel(out[f'stage{st+1}']['depth'], ref[f'stage{st+1}']['depth']):.2e} conf {(out[f'stage{st+1}']['photometric_confidence'].cpu() - ref[f'stage{st+1}']['photometric_confidence']).abs().mean():.2e}")
