"""Refinement network and the refine=True cascade against golden outputs of the live reference
(tests/golden/make_golden_refine.py)."""
import numpy as np
import pytest
import torch

import cds_mvsnet_b200 as C
from cds_mvsnet_b200 import synthetic
from oracle import oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"
T = 0.01
torch.set_grad_enabled(False)


def _weights(golden, pretrained_sd):
    sd = dict(pretrained_sd)
    sd.update(golden("weights_refine_both_dtu_blended"))
    return sd


def test_refinement_module_vs_live_reference(golden, pretrained_sd):
    g = golden("refine")
    sd = _weights(golden, pretrained_sd)
    net = C.Refinement()
    net.load_state_dict({k[len("refine_network."):]: v for k, v in sd.items() if k.startswith("refine_network.")}, strict=True)
    net = net.to(DEV).eval()
    out = net(g["img"].to(DEV), g["depth_0"].to(DEV), g["depth_min"].to(DEV), g["depth_max"].to(DEV))
    assert out.shape == g["refined"].shape
    torch.testing.assert_close(out.cpu(), g["refined"], rtol=1e-5, atol=2e-4)     # fp32 both sides; depths ~ 150..300
    with pytest.raises(RuntimeError, match="depth_0"):
        net(g["img"].to(DEV), g["depth_0"].to(DEV)[..., :-1], g["depth_min"].to(DEV), g["depth_max"].to(DEV))


@pytest.mark.parametrize("storage,tol", [(torch.float32, 2e-5), (torch.float16, 1e-3)])
def test_cascade_with_refinement_vs_live_reference(golden, pretrained_sd, storage, tol):
    g = golden("refine")
    W_, H_, N, B, Dtot = (int(v) for v in g["cfg"][:5])
    cfg = dict(W=W_, H=H_, N=N, B=B, Dtot=Dtot, ndepths=tuple(int(v) for v in g["cfg"][5:]), ratios=tuple(float(r) for r in g["ratios"]),
               interval=float(g["interval"]))
    s = synthetic.make_sample(cfg, "noise", seed=0)          # cameras / depth range of the half-resolution cascade
    sd = _weights(golden, pretrained_sd)
    m = C.CDSMVSNet(refine=True, ndepths=cfg["ndepths"], depth_interals_ratio=cfg["ratios"], storage=storage)
    m.load_state_dict({k: v for k, v in sd.items() if k in m.state_dict()}, strict=True)
    m = m.to(DEV).eval()
    out = m(g["e2e_imgs"].to(DEV), {k: v.to(DEV) for k, v in s.proj_matrices.items()}, s.depth_values.to(DEV), temperature=T)
    assert out["depth"].shape == g["e2e_depth"].shape and out["refined_depth"].shape == g["e2e_refined"].shape
    assert out["refined_depth"].shape[-1] == 2 * out["depth"].shape[-1]
    e_d, e_r = O.rel_l1(out["depth"].cpu(), g["e2e_depth"]), O.rel_l1(out["refined_depth"].cpu(), g["e2e_refined"])
    print(f"refine=True storage={storage}: depth rel-L1 {e_d:.3e}  refined rel-L1 {e_r:.3e}")
    assert e_d < tol and e_r < tol
    with pytest.raises(RuntimeError, match="divisible by 64"):
        m(g["e2e_imgs"].to(DEV)[..., :96, :], {k: v.to(DEV) for k, v in s.proj_matrices.items()}, s.depth_values.to(DEV), temperature=T)
