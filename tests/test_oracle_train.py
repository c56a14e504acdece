"""Training slice: the oracle's backward / loss restatements against the gradients of the LIVE reference
(tests/golden/make_golden_train.py -> train_ops.npz).  CPU only."""
import pytest
import torch

from oracle import oracle as O

torch.set_grad_enabled(False)


def close(a, b, atol, rtol=1e-5):
    assert a.shape == b.shape, (a.shape, b.shape)
    torch.testing.assert_close(a, b, atol=atol, rtol=rtol)


def test_warp_backward(golden):
    g = golden("train_ops")
    for tag in ("planes", "pix"):
        got = O.homo_warp_backward(g["warp_grad_out"], g["warp_src_proj"], g["warp_ref_proj"], g[f"warp_depth_{tag}"])
        close(got, g[f"warp_grad_src_{tag}"], 4e-6)
    # the adjoint identity <warp(x), g> = <x, warp^T(g)> ties the backward restatement to the forward one
    x = g["warp_src_fea"]
    fwd = O.homo_warp(x, g["warp_src_proj"], g["warp_ref_proj"], g["warp_depth_pix"])
    lhs = (fwd.double() * g["warp_grad_out"].double()).sum()
    rhs = (x.double() * g["warp_grad_src_pix"].double()).sum()
    assert abs(lhs - rhs) <= 1e-6 * abs(lhs)


def test_depth_regression_backward(golden):
    g = golden("train_ops")
    for tag in ("planes", "pix"):
        gp, gdv = O.depth_regression_backward(g["regress_grad_depth"], g["regress_p"], g[f"warp_depth_{tag}"])
        close(gp, g[f"regress_grad_p_{tag}"], 0.0, 0.0)
        close(gdv, g[f"regress_grad_dv_{tag}"], 1e-5)


def loss_case(g):
    inputs = {f"stage{i}": {"depth": g[f"loss_in_stage{i}.depth"], "norm_curv": g[f"loss_in_stage{i}.norm_curv"]} for i in (1, 2, 3)}
    inputs["refined_depth"] = g["loss_in_refined_depth"]
    gts = {f"stage{i}": g[f"loss_gt_stage{i}"] for i in (1, 2, 3, 4)}
    masks = {f"stage{i}": g[f"loss_mask_stage{i}"] for i in (1, 2, 3, 4)}
    return inputs, gts, masks


def test_final_loss(golden):
    g = golden("train_ops")
    inputs, gts, masks = loss_case(g)
    w = g["loss_dlossw"].tolist()
    total, dl = O.final_loss(inputs, gts, masks, dlossw=w, depth_interval=g["loss_interval"])
    close(total, g["loss_total"], 1e-6)
    close(dl, g["loss_depth"], 1e-6)
    for i in (1, 2, 3):
        k = f"stage{i}"
        ge, gc = O.stage_loss_backward(inputs[k]["depth"], gts[k], masks[k], g["loss_interval"], w[i - 1], 0.1 * w[i - 1])
        close(ge, g[f"loss_grad_{k}.depth"], 1e-9, 1e-5)
        close(gc.unsqueeze(1), g[f"loss_grad_{k}.norm_curv"], 1e-9, 1e-5)
    ge, _ = O.stage_loss_backward(inputs["refined_depth"], gts["stage4"], masks["stage4"], g["loss_interval"], 2.0, 0.0)
    close(ge, g["loss_grad_refined_depth"], 1e-9, 1e-5)


def test_final_loss_with_feat_term(golden):
    g = golden("train_ops")
    inputs, gts, masks = loss_case(g)
    del inputs["refined_depth"]
    w = g["loss_dlossw"].tolist()
    for i in (1, 2, 3):
        k = f"stage{i}"
        inputs[k]["feat_distance"], inputs[k]["feat_target"] = g[f"lossf_in_{k}.feat_distance"], g[f"lossf_in_{k}.feat_target"]
        close(O.feat_loss(inputs[k]["feat_distance"], inputs[k]["feat_target"], masks[k]), g[f"lossf_value_{k}"], 1e-6)
        gf = O.feat_loss_backward(inputs[k]["feat_distance"], inputs[k]["feat_target"], masks[k], 5 * w[i - 1])
        close(gf, g[f"lossf_grad_{k}.feat_distance"], 1e-8, 1e-5)
    total, dl = O.final_loss(inputs, gts, masks, dlossw=w, depth_interval=g["loss_interval"])
    close(total, g["lossf_total"], 1e-5)
    close(dl, g["lossf_depth"], 1e-6)


@pytest.mark.parametrize("seed,h,w,C_,M", [(0, 2, 2, 1, 40), (1, 5, 9, 3, 200), (2, 16, 7, 4, 500), (3, 3, 31, 2, 1)])
def test_scatter_is_the_adjoint_of_the_gather(seed, h, w, C_, M):
    """<gather(x), g> = <x, scatter(g)> on ragged shapes with samples on and beyond every border (zero padding drops them on
    both sides alike), including the degenerate single-sample case."""
    torch.manual_seed(seed)
    u = torch.rand(2, M) * (w + 4) - 2.5
    v = torch.rand(2, M) * (h + 4) - 2.5
    u[:, 0], v[:, 0] = w - 1.0, h - 1.0            # exactly on the last pixel: only the (0,0) tap is inside
    x, g = torch.randn(2, C_, h, w), torch.randn(2, C_, M)
    lhs = (O.bilinear_gather_zeros(x, u, v).double() * g.double()).sum()
    rhs = (x.double() * O.bilinear_scatter_zeros(g, u, v, h, w).double()).sum()
    assert abs(lhs - rhs) <= 1e-5 * max(1.0, abs(lhs))


def test_scatter_of_nothing_is_zero():
    z = O.bilinear_scatter_zeros(torch.randn(1, 2, 3), torch.full((1, 3), -50.0), torch.full((1, 3), 1e6), 4, 5)
    assert z.shape == (1, 2, 4, 5) and z.abs().max() == 0


def test_temperature_schedule():
    """trainer/trainer.py:45-49."""
    from cds_mvsnet_b200.losses import temperature_for_epoch
    want = {1: 1.0, 2: 10 ** -0.5, 3: 0.1, 4: 10 ** -1.5, 5: 0.01, 17: 0.01}
    for e, t in want.items():
        assert abs(temperature_for_epoch(e) - t) < 1e-12
