"""Geometric-consistency filter on the CUDA kernels against golden outputs of the live reference
(tests/golden/make_golden_fusion.py) and against the CPU oracle on other shapes / seeds."""
import numpy as np
import pytest
import torch

from cds_mvsnet_b200 import fusion as F
from cds_mvsnet_b200 import synthetic
from oracle import fusion_oracle as FO

pytestmark = pytest.mark.gpu
DEV = "cuda"
torch.set_grad_enabled(False)


def _cuda(s):
    return {k: v.to(DEV) for k, v in s.items()}


def _oracle(s, pt, disp, dth, vth):
    sd = s["src_depths"].clone()
    for i in range(sd.size(1)):
        sd[:, i] *= FO.prob_filter(s["src_confs"][:, i], pt).float()
    xyd, inr = FO.get_reproj(s["ref_depth"], sd, s["ref_cam"], s["src_cams"])
    masks, mask = FO.vis_filter(s["ref_depth"], xyd, inr, disp, dth, vth)
    ave = FO.ave_fusion(s["ref_depth"], xyd, masks)
    return dict(reproj_xyd=xyd, in_range=inr, masks=masks, vis_mask=mask, ave=ave, points=FO.back_project(ave, s["ref_cam"]),
                prob_mask=FO.prob_filter(s["ref_conf"], pt), src_depths_masked=sd)


def _clean_footprints(s, sd_masked):
    """[n,v,1,h,w] bool: the four bilinear taps of the reference pixel's projection into source view v all carry a depth.
    Where a tap lands on a masked (zero-depth) pixel the sampled (x, y, depth) is a blend with the projection of that
    camera's centre: huge, ill-conditioned values (a rounding of the bilinear weight moves them by centimetres) on BOTH
    sides, rejected by every mask -- positions are compared on clean footprints only, masks everywhere."""
    n, v, _, h, w = sd_masked.shape
    g = FO.pixel_grids(h, w).unsqueeze(0)
    world = FO.cam2world(FO.img2cam(g, s["ref_depth"], s["ref_cam"]), s["ref_cam"])
    out = torch.zeros(n, v, 1, h, w, dtype=torch.bool)
    for vi in range(v):
        img = FO.cam2img(FO.world2cam(world, s["src_cams"][:, vi]), s["src_cams"][:, vi])[..., :2, 0]      # [n,h,w,2]
        gx = ((img[..., 0] / w * 2 - 1).clamp(-1.1, 1.1) + 1) / 2 * (w - 1)
        gy = ((img[..., 1] / h * 2 - 1).clamp(-1.1, 1.1) + 1) / 2 * (h - 1)
        x0, y0 = gx.floor().long(), gy.floor().long()
        ok = (x0 >= 0) & (x0 + 1 < w) & (y0 >= 0) & (y0 + 1 < h)
        x0c, y0c = x0.clamp(0, w - 2), y0.clamp(0, h - 2)
        d = sd_masked[:, vi, 0]
        bi = torch.arange(n).view(n, 1, 1).expand_as(x0c)
        for dy in (0, 1):
            for dx in (0, 1):
                ok &= d[bi, y0c + dy, x0c + dx] > 0
        out[:, vi, 0] = ok
    return out


def _check(got, ref, clean):
    """Positions / depths to 5e-3 on clean footprints; masks may differ only where a test sits on its threshold."""
    assert torch.equal(got["in_range"].cpu(), ref["in_range"])
    xyd = ref["reproj_xyd"]
    sel = (clean & ref["in_range"].bool()).expand_as(xyd)
    assert sel.float().mean() > 0.1
    d = (got["reproj_xyd"].cpu() - xyd).abs()
    assert d[sel].max() < 5e-3, d[sel].max()
    flips = (got["masks"].cpu() != ref["masks"]).float().mean().item()
    assert flips < 2e-3, flips
    assert (got["vis_mask"].cpu() != ref["vis_mask"]).float().mean().item() < 2e-3
    same = (got["masks"].cpu() == ref["masks"]).all(dim=1)
    assert ((got["ave"].cpu() - ref["ave"]).abs()[same] < 5e-3).all()
    assert ((got["points"].cpu() - ref["points"]).abs()[same.expand(-1, 3, -1, -1)] < 1e-2).all()


def test_fused_filter_vs_live_reference_golden(golden):
    g = golden("fusion_small")
    H, W, V, seed, B = (int(x) for x in g["cfg"])
    s = synthetic.make_fusion_sample(H, W, V, seed=seed, batch=B)
    pt = tuple(float(x) for x in g["thresholds"][:3])
    disp, dth, vth = (float(x) for x in g["thresholds"][3:])
    c = _cuda(s)
    out = F.geometric_filter(c["ref_depth"], c["src_depths"], c["ref_cam"], c["src_cams"], disp, dth, vth, ref_conf=c["ref_conf"],
                             srcs_conf=c["src_confs"], prob_thresh=pt, want_reproj=True)
    ref = {k: (v if torch.is_tensor(v) else torch.from_numpy(np.asarray(v))) for k, v in g.items() if k not in ("cfg", "thresholds")}
    sd = s["src_depths"].clone()
    for i in range(V):
        sd[:, i] *= FO.prob_filter(s["src_confs"][:, i], pt).float()
    _check(out, ref, _clean_footprints(s, sd))
    assert torch.equal(out["final_mask"].cpu(), out["prob_mask"].cpu() & out["vis_mask"].cpu())
    assert (out["final_mask"].cpu() != ref["final_mask"]).float().mean().item() < 2e-3
    assert 0.05 < ref["vis_mask"].float().mean() < 0.95      # the fixture exercises both outcomes
    assert torch.equal(c["src_depths"].cpu(), s["src_depths"])  # caller's depths are not modified


@pytest.mark.parametrize("hw,V,B,seed", [((64, 80), 2, 2, 1), ((37, 53), 5, 1, 2), ((120, 160), 10, 1, 3)])
def test_op_level_dropins_vs_oracle(hw, V, B, seed):
    """get_reproj / vis_filter / ave_fusion / prob_filter with the reference's signatures, on odd sizes, batches and 10 views."""
    s = synthetic.make_fusion_sample(hw[0], hw[1], V, seed=seed, batch=B)
    pt, disp, dth, vth = (0.2, 0.1), 1.0, 0.01, 2
    ref = _oracle(s, pt, disp, dth, vth)
    c = _cuda(s)
    sd = c["src_depths"].clone()
    for i in range(V):
        sd[:, i] *= F.prob_filter(c["src_confs"][:, i], pt).float()
    assert torch.equal(sd.cpu(), ref["src_depths_masked"])
    assert torch.equal(F.prob_filter(c["ref_conf"], pt).cpu(), ref["prob_mask"])
    xyd, inr = F.get_reproj(c["ref_depth"], sd, c["ref_cam"], c["src_cams"])
    assert xyd.shape == ref["reproj_xyd"].shape and inr.shape == ref["in_range"].shape
    masks, mask = F.vis_filter(c["ref_depth"], xyd, inr, disp, dth, vth)
    ave = F.ave_fusion(c["ref_depth"], xyd, masks)
    assert mask.dtype == torch.bool and masks.shape == ref["masks"].shape and ave.shape == ref["ave"].shape
    got = dict(reproj_xyd=xyd, in_range=inr, masks=masks, vis_mask=mask, ave=ave, points=ref["points"].to(DEV))
    clean = _clean_footprints(s, ref["src_depths_masked"])
    _check(got, ref, clean)
    # the op-level chain and the fused pass are the same arithmetic
    fused = F.geometric_filter(c["ref_depth"], sd, c["ref_cam"], c["src_cams"], disp, dth, vth, want_reproj=True)
    assert torch.equal(fused["reproj_xyd"], xyd) and torch.equal(fused["masks"], masks) and torch.equal(fused["ave"], ave)
    _check(fused, ref, clean)


def test_fusion_rejects_bad_inputs():
    s = synthetic.make_fusion_sample(32, 40, 2, seed=0)
    with pytest.raises(RuntimeError, match="CUDA"):
        F.get_reproj(s["ref_depth"], s["src_depths"], s["ref_cam"], s["src_cams"])
    c = _cuda(s)
    with pytest.raises(AssertionError):
        F.get_reproj(c["ref_depth"], c["src_depths"], c["ref_cam"][:, :1], c["src_cams"])
    with pytest.raises(AssertionError):
        F.prob_filter(c["ref_conf"], (0.1, 0.1, 0.1, 0.1, 0.1))
