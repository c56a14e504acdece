"""Golden vectors of the LIVE reference for the training slice: autograd of homo_warping_3D and
depth_regression, and models/losses.py:final_loss with its gradients.  Run in the build container:

    python tests/golden/make_golden_train.py      # -> tests/golden/train_ops.npz
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

with contextlib.redirect_stdout(io.StringIO()):
    from models.utils.warping import homo_warping_3D as ref_warp  # noqa: E402
    from models.module import depth_regression as ref_regress  # noqa: E402
    from models.losses import final_loss as ref_loss  # noqa: E402

from cds_mvsnet_b200 import synthetic  # noqa: E402
from oracle import oracle as O  # noqa: E402


def dev(tag, a, b):
    print(f"    oracle vs reference  {tag:<34s} max-abs {(a - b).abs().max().item():.3e}   rel-L1 {O.rel_l1(a, b):.3e}")


def main():
    torch.manual_seed(7)
    out = {}
    # ---- warp backward: d/d src_fea of sum(out * grad_out), plane depths and per-pixel depths
    s = synthetic.make_sample(dict(W=96, H=64, N=3, ndepths=(8,), ratios=(1.0,), B=2, Dtot=192, interval=2.65))
    pm = s.proj_matrices["stage2"]
    h, w, C, D = 32, 48, 5, 6
    ref_P, src_P = O.compose_projection(pm[:, 0]), O.compose_projection(pm[:, 2])
    fea = torch.randn(2, C, h, w)
    dv_planes = torch.linspace(430, 930, D).unsqueeze(0).repeat(2, 1)
    dv_pix = (dv_planes.reshape(2, D, 1, 1) + 40 * torch.rand(2, D, h, w)).contiguous()
    dv_pix[:, 0] = 5.0                      # a plane far outside the frustum: every tap dropped
    g_out = torch.randn(2, C, D, h, w)
    for tag, dv in (("planes", dv_planes), ("pix", dv_pix)):
        f = fea.clone().requires_grad_(True)
        ref_warp(f, src_P, ref_P, dv).backward(g_out)
        out[f"warp_grad_src_{tag}"] = f.grad.clone()
        dev(f"warp backward [{tag}]", O.homo_warp_backward(g_out, src_P, ref_P, dv), f.grad)
    out.update(warp_src_fea=fea, warp_src_proj=src_P, warp_ref_proj=ref_P, warp_depth_planes=dv_planes, warp_depth_pix=dv_pix,
               warp_grad_out=g_out)

    # ---- depth_regression backward
    p = torch.softmax(torch.randn(2, D, h, w), 1)
    g_d = torch.randn(2, h, w)
    for tag, dv in (("planes", dv_planes), ("pix", dv_pix)):
        pp, dd = p.clone().requires_grad_(True), dv.clone().requires_grad_(True)
        ref_regress(pp, dd).backward(g_d)
        out[f"regress_grad_p_{tag}"], out[f"regress_grad_dv_{tag}"] = pp.grad.clone(), dd.grad.clone()
        gp, gdv = O.depth_regression_backward(g_d, p, dv)
        dev(f"regress d/dp [{tag}]", gp, pp.grad)
        dev(f"regress d/d depth [{tag}]", gdv, dd.grad)
    out.update(regress_p=p, regress_grad_depth=g_d)

    # ---- final_loss (eval-mode output dict: no feat_distance), with refined_depth, weights 0.5/1/2
    B = 2
    interval = torch.tensor([2.65, 2.5])
    inputs, gts, masks = {}, {}, {}
    for i, (hh, ww) in enumerate(((8, 12), (16, 24), (32, 48), (64, 96))):
        k = f"stage{i + 1}"
        gt = 425 + 500 * torch.rand(B, hh, ww)
        gts[k] = gt
        masks[k] = (torch.rand(B, hh, ww) > 0.3).float()
        est = gt + torch.randn(B, hh, ww) * interval.reshape(B, 1, 1) * 1.5      # errors on both sides of the |x| = 1 knee
        if i < 3:
            inputs[k] = {"depth": est, "norm_curv": torch.rand(B, 1, hh, ww)}
        else:
            inputs["refined_depth"] = est
    leaves = {}
    live = {}
    for k, v in inputs.items():
        if isinstance(v, dict):
            live[k] = {kk: vv.clone().requires_grad_(True) for kk, vv in v.items()}
            leaves.update({f"{k}.{kk}": vv for kk, vv in live[k].items()})
        else:
            live[k] = v.clone().requires_grad_(True)
            leaves[k] = live[k]
    dlossw = [0.5, 1.0, 2.0]
    total, dl = ref_loss(live, gts, masks, dlossw=dlossw, depth_interval=interval)
    total.backward()
    o_total, o_dl = O.final_loss(inputs, gts, masks, dlossw=dlossw, depth_interval=interval)
    dev("final_loss total", o_total, total.detach())
    dev("final_loss depth_loss", o_dl, dl.detach())
    ge, gc = O.stage_loss_backward(inputs["stage2"]["depth"], gts["stage2"], masks["stage2"], interval, 1.0, 0.1)
    dev("final_loss d/d stage2.depth", ge, leaves["stage2.depth"].grad)
    dev("final_loss d/d stage2.norm_curv", gc.unsqueeze(1), leaves["stage2.norm_curv"].grad)
    out.update(loss_total=total.detach(), loss_depth=dl.detach(), loss_interval=interval, loss_dlossw=torch.tensor(dlossw))
    for k in gts:
        out[f"loss_gt_{k}"], out[f"loss_mask_{k}"] = gts[k], masks[k]
    for k, v in leaves.items():
        out[f"loss_in_{k}"] = v.detach()
        out[f"loss_grad_{k}"] = v.grad
    # ---- the same loss on a training-mode output dict: feat_distance / feat_target present (D + 1 planes, losses.py:25-35)
    torch.manual_seed(8)
    live2, leaves2 = {}, {}
    for i, (hh, ww) in enumerate(((8, 12), (16, 24), (32, 48))):
        k = f"stage{i + 1}"
        nd = 5
        fd = torch.randn(B, nd, hh, ww) * 3
        ft = (torch.rand(B, nd, hh, ww) > 0.8).float()
        live2[k] = {"depth": inputs[k]["depth"].clone().requires_grad_(True), "norm_curv": inputs[k]["norm_curv"].clone().requires_grad_(True),
                    "feat_distance": fd.clone().requires_grad_(True), "feat_target": ft}
        out[f"lossf_in_{k}.feat_distance"], out[f"lossf_in_{k}.feat_target"] = fd, ft
        sel = masks[k].unsqueeze(1).repeat(1, nd, 1, 1) > 0.5       # losses.py:29-34 spelled out with the torch operator
        pos = ft[sel].sum()
        want = torch.nn.functional.binary_cross_entropy_with_logits(fd[sel], ft[sel], reduction="mean",
                                                                    pos_weight=(torch.numel(ft[sel]) - pos) / pos)
        dev(f"feat_loss {k}", O.feat_loss(fd, ft, masks[k]), want)
        out[f"lossf_value_{k}"] = want
    total2, dl2 = ref_loss(live2, gts, masks, dlossw=dlossw, depth_interval=interval)
    total2.backward()
    plain = {k: {kk: vv.detach() for kk, vv in v.items()} for k, v in live2.items()}
    o_total2, _ = O.final_loss(plain, gts, masks, dlossw=dlossw, depth_interval=interval)
    dev("final_loss (feat term) total", o_total2, total2.detach())
    for i in (1, 2, 3):
        k = f"stage{i}"
        gf = O.feat_loss_backward(plain[k]["feat_distance"], plain[k]["feat_target"], masks[k], 5 * dlossw[i - 1])
        dev(f"final_loss d/d {k}.feat_distance", gf, live2[k]["feat_distance"].grad)
        out[f"lossf_grad_{k}.feat_distance"] = live2[k]["feat_distance"].grad
        out[f"lossf_grad_{k}.depth"] = live2[k]["depth"].grad
    out.update(lossf_total=total2.detach(), lossf_depth=dl2.detach())
    path = os.path.join(HERE, "train_ops.npz")
    np.savez_compressed(path, **{k: v.numpy() for k, v in out.items()})
    print(f"wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB)")


if __name__ == "__main__":
    main()
