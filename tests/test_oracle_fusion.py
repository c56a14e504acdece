"""The fusion oracle (CPU restatement of the reference's fusion.py) against golden outputs of the LIVE reference."""
import numpy as np
import torch

from cds_mvsnet_b200 import synthetic
from oracle import fusion_oracle as FO

torch.set_grad_enabled(False)


def test_fusion_oracle_matches_live_reference(golden):
    g = golden("fusion_small")
    H, W, V, seed, B = (int(x) for x in g["cfg"])
    s = synthetic.make_fusion_sample(H, W, V, seed=seed, batch=B)
    pt = tuple(float(x) for x in g["thresholds"][:3])
    disp, dth, vth = (float(x) for x in g["thresholds"][3:])
    sd = s["src_depths"].clone()
    for i in range(V):
        sd[:, i] *= FO.prob_filter(s["src_confs"][:, i], pt).float()
    xyd, inr = FO.get_reproj(s["ref_depth"], sd, s["ref_cam"], s["src_cams"])
    masks, mask = FO.vis_filter(s["ref_depth"], xyd, inr, disp, dth, vth)
    ave = FO.ave_fusion(s["ref_depth"], xyd, masks)
    pts = FO.back_project(ave, s["ref_cam"])
    final = FO.prob_filter(s["ref_conf"], pt) & mask
    as_t = lambda k: g[k] if torch.is_tensor(g[k]) else torch.from_numpy(np.asarray(g[k]))
    for name, got in (("reproj_xyd", xyd), ("in_range", inr), ("masks", masks), ("ave", ave), ("points", pts)):
        torch.testing.assert_close(got, as_t(name), rtol=1e-6, atol=1e-4)
    assert torch.equal(mask, as_t("vis_mask").bool()) and torch.equal(final, as_t("final_mask").bool())
