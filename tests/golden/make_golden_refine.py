"""Golden outputs of the LIVE reference with refine=True (the configuration of all three pretrained checkpoints): the
Refinement network alone and the whole cascade.  Run in the build container:

    python tests/golden/make_golden_refine.py     # -> tests/golden/refine.npz, weights_refine_both_dtu_blended.npz
"""
import contextlib
import io
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("CDS_REF_PATH", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

from models import model as ref_model  # noqa: E402
from models import module as ref_module  # noqa: E402

from cds_mvsnet_b200 import synthetic  # noqa: E402
from oracle import oracle as O  # noqa: E402

torch.set_grad_enabled(False)
T = 0.01


def main():
    ck = torch.load(os.path.join(REF, "pretrained/both_dtu_blended/cds_mvsnet.ckpt"), map_location="cpu", weights_only=False)
    sd = {k: (v.float() if v.is_floating_point() else v) for k, v in O.strip_module_prefix(ck["state_dict"]).items()}
    rsd = {k: v for k, v in sd.items() if k.startswith("refine_network.")}
    np.savez_compressed(os.path.join(HERE, "weights_refine_both_dtu_blended.npz"), **{k: v.numpy() for k, v in rsd.items()})
    print("refine_network entries:", len(rsd))

    # ---- the Refinement network alone (module.py:318-370)
    net = ref_module.Refinement()
    net.load_state_dict({k[len("refine_network."):]: v for k, v in rsd.items()})
    net.eval()
    torch.manual_seed(3)
    img = torch.rand(2, 3, 64, 96)
    d0 = 160 + 120 * torch.rand(2, 1, 32, 48)
    dmin, dmax = torch.tensor([150.0, 155.0]), torch.tensor([300.0, 290.0])
    out = net(img, d0, dmin, dmax)
    mine = O.refinement(sd, img, d0, dmin, dmax)
    print("  oracle vs reference Refinement: max-abs", (mine - out).abs().max().item())

    # ---- the whole cascade with refine=True: images at 128x192, cameras / cascade at 64x96
    cfg = dict(W=96, H=64, N=3, ndepths=(16, 8, 8), ratios=(4.0, 1.5, 0.75), B=1, Dtot=192, interval=2.65)
    s = synthetic.make_sample(cfg, "noise", seed=0)                  # cameras, depth_values for the half-resolution cascade
    torch.manual_seed(5)
    imgs = torch.rand(1, 3, 3, 128, 192)
    with contextlib.redirect_stdout(io.StringIO()):
        m = ref_model.CDSMVSNet(refine=True, ndepths=cfg["ndepths"], depth_interals_ratio=cfg["ratios"])
    m.load_state_dict({k: v for k, v in sd.items() if k in m.state_dict()}, strict=True)
    m.eval()
    o = m(imgs, s.proj_matrices, s.depth_values, temperature=T)
    mine_o = O.cdsmvsnet_forward(sd, imgs, s.proj_matrices, s.depth_values, cfg["ndepths"], cfg["ratios"], T, refine=True)
    print("  oracle vs reference cascade(refine=True): depth rel-L1", O.rel_l1(mine_o["depth"], o["depth"]),
          " refined rel-L1", O.rel_l1(mine_o["refined_depth"], o["refined_depth"]))
    np.savez_compressed(os.path.join(HERE, "refine.npz"), img=img.numpy(), depth_0=d0.numpy(), depth_min=dmin.numpy(), depth_max=dmax.numpy(),
                        refined=out.numpy(), e2e_imgs=imgs.numpy(), e2e_depth=o["depth"].numpy(), e2e_refined=o["refined_depth"].numpy(),
                        e2e_conf=o["photometric_confidence"].numpy(), cfg=np.array([96, 64, 3, 1, 192, 16, 8, 8]),
                        ratios=np.array(cfg["ratios"]), interval=np.array(cfg["interval"]))
    print("  wrote refine.npz", os.path.getsize(os.path.join(HERE, "refine.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
