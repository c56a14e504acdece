"""The oracle restatement against the golden vectors generated from the live reference
(tests/golden/make_golden.py).  CPU only.  Tolerances are the fp32 re-association noise floor
measured for the reference itself (SURVEY.md 8c fixture 5: ~1e-6 relative)."""
import hashlib

import pytest
import torch

from cds_mvsnet_b200 import synthetic
from oracle import oracle as O

T = 0.01
torch.set_grad_enabled(False)


def close(a, b, atol, rtol=1e-5):
    assert a.shape == b.shape, (a.shape, b.shape)
    torch.testing.assert_close(a, b, atol=atol, rtol=rtol)


def test_warp(golden):
    g = golden("warp")
    close(O.homo_warp(g["src_fea"], g["src_proj"], g["ref_proj"], g["depth_planes"]), g["out_planes"], 2e-6)
    close(O.homo_warp(g["src_fea"], g["src_proj"], g["ref_proj"], g["depth_pix"]), g["out_pix"], 2e-6)
    # the far-out-of-frustum plane is all zero padding
    assert g["out_pix"][:, :, 0].abs().max() == 0


def test_fast_gather(golden):
    """The grid_sample shortcut bench.py times on the CPU is the same function."""
    g = golden("warp")
    O.FAST_GATHER = True
    try:
        close(O.homo_warp(g["src_fea"], g["src_proj"], g["ref_proj"], g["depth_pix"]), g["out_pix"], 2e-6)
    finally:
        O.FAST_GATHER = False


def test_compose_projection(golden):
    g = golden("warp")
    close(O.compose_projection(g["cams"][:, 1]), g["src_proj"], 0, 0)


def test_epipole(golden):
    g = golden("epipole")
    Fm = O.fundamental_matrix(g["cam_ref"], g["cam_src"])
    close(Fm, g["F"], 1e-6)
    close(O.epipole_from_F(Fm), g["e_ref"], 1e-2, 1e-5)
    close(O.epipole_from_F(Fm.transpose(1, 2)), g["e_src"], 1e-2, 1e-5)


@pytest.mark.parametrize("tag,D,ratio,scale", [("s1", 48, 4.0, 4), ("s2", 32, 1.5, 2), ("s3", 8, 0.75, 1)])
def test_hypotheses(golden, tag, D, ratio, scale):
    g = golden("hypotheses")
    dv = g["depth_values"]
    B = dv.shape[0]
    out = g[f"{tag}_out"]
    H, W = out.shape[2] * scale, out.shape[3] * scale
    mine = O.depth_hypotheses(g[f"{tag}_prev"], D, ratio * (dv[:, 1] - dv[:, 0]), H, W, dv[:, 0], dv[:, -1], scale)
    close(mine, out, 3e-4, 0)
    assert mine.min() >= dv[0, 0] and mine.max() <= dv[0, -1]


def test_tail(golden):
    g = golden("tail")
    p = torch.softmax(g["logits"], 1)
    close(O.depth_regression(p, g["depth_samples"]), g["depth"], 1e-4)
    close(O.conf_regression(p), g["conf"], 1e-6)
    # window at the borders: index 0 sums p[0..2], index D-1 sums p[D-2..D-1]
    assert g["conf"][0, 0, 0] == pytest.approx(float(p[0, 0:3, 0, 0].sum()), abs=1e-6)
    assert g["conf"][0, 0, 1] == pytest.approx(float(p[0, 5:8, 0, 1].sum()), abs=1e-6)


@pytest.mark.parametrize("name", ["conv00", "conv01", "conv10", "conv20", "out1", "out3"])
def test_dynamic_conv(golden, pretrained_sd, name):
    g = golden("dynconv")
    pre = f"feature.{name}" + ("" if name.startswith("out") else ".conv")
    y, nc = O.dynamic_conv(g[f"{name}_x"], pretrained_sd, pre, O.FEATURE_DYN_KSIZES[name], g[f"{name}_epi"], T)
    close(y, g[f"{name}_y"], 2e-6)
    close(nc, g[f"{name}_nc"], 2e-6)


def test_feature_net(golden, pretrained_sd):
    g = golden("featurenet")
    out = O.feature_net(g["img"], pretrained_sd, g["epi"], T)
    for st in ("stage1", "stage2", "stage3"):
        for j, nm in enumerate(("fea", "nc_sum", "nc_abs")):
            close(out[st][j], g[f"{st}_{nm}"], 5e-5)
        assert out[st][0].abs().max() <= 1.0  # tanh range


@pytest.mark.parametrize("st", [0, 1, 2])
def test_vis_and_costreg(golden, pretrained_sd, st):
    g = golden("nets3d")
    close(O.vis_net(g[f"vis{st}_x"], pretrained_sd, f"stage_net.vis.{st}"), g[f"vis{st}_y"], 2e-6)
    close(O.cost_reg_net(g[f"cr{st}_x"], pretrained_sd, f"cost_regularization.{st}"), g[f"cr{st}_y"], 5e-5)


def _e2e(golden, pretrained_sd, tag, family):
    g = golden(tag)
    W, H, N, B, Dtot = (int(v) for v in g["cfg"][:5])
    nd = tuple(int(v) for v in g["cfg"][5:])
    cfg = dict(W=W, H=H, N=N, B=B, Dtot=Dtot, ndepths=nd, ratios=tuple(float(r) for r in g["ratios"]),
               interval=float(g["interval"]))
    s = synthetic.make_sample(cfg, family, seed=0)
    digest = hashlib.sha256(s.imgs.numpy().tobytes()).hexdigest()[:16]
    assert digest == bytes(g["imgs_sha"].numpy()).decode(), "synthetic generator drifted from the golden inputs"
    out = O.cdsmvsnet_forward(pretrained_sd, s.imgs, s.proj_matrices, s.depth_values, nd, cfg["ratios"], T)
    return g, s, out, nd


@pytest.mark.parametrize("tag,family", [("e2e_cfg1_noise", "noise"), ("e2e_small3_plane", "plane"),
                                        ("e2e_small3_noise", "noise")])
def test_end_to_end(golden, pretrained_sd, tag, family):
    g, s, out, nd = _e2e(golden, pretrained_sd, tag, family)
    for st in range(len(nd)):
        nm = f"stage{st + 1}"
        assert O.rel_l1(out[nm]["depth"], g[f"{nm}_depth"]) < 1e-5
        assert (out[nm]["photometric_confidence"] - g[f"{nm}_photometric_confidence"]).abs().mean() < 1e-4
        close(out[nm]["norm_curv"], g[f"{nm}_norm_curv"], 1e-5, 1e-4)
    if family == "plane":
        # known-answer: the photo-consistent slanted plane is recovered
        assert (out["depth"] - s.gt_depth).abs().mean() < 8.0


def test_refinement_oracle_matches_live_reference(golden, pretrained_sd):
    """Refinement network and the refine=True cascade of the oracle against the live reference's outputs."""
    from cds_mvsnet_b200 import synthetic
    g = golden("refine")
    sd = dict(pretrained_sd)
    sd.update(golden("weights_refine_both_dtu_blended"))
    out = O.refinement(sd, g["img"], g["depth_0"], g["depth_min"], g["depth_max"])
    torch.testing.assert_close(out, g["refined"], rtol=1e-6, atol=1e-4)
    W_, H_, N, B, Dtot = (int(v) for v in g["cfg"][:5])
    cfg = dict(W=W_, H=H_, N=N, B=B, Dtot=Dtot, ndepths=tuple(int(v) for v in g["cfg"][5:]), ratios=tuple(float(r) for r in g["ratios"]),
               interval=float(g["interval"]))
    s = synthetic.make_sample(cfg, "noise", seed=0)
    o = O.cdsmvsnet_forward(sd, g["e2e_imgs"], s.proj_matrices, s.depth_values, cfg["ndepths"], cfg["ratios"], 0.01, refine=True)
    assert O.rel_l1(o["depth"], g["e2e_depth"]) < 1e-5 and O.rel_l1(o["refined_depth"], g["e2e_refined"]) < 1e-5
