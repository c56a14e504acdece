"""Generate the golden fixtures in this directory from the LIVE reference.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

It imports the reference's ``models`` package unmodified, loads the pretrained
``both_dtu_blended`` checkpoint, runs every hot-path function of SURVEY.md section 8(a) on
small seeded inputs and stores inputs + outputs as ``*.npz``.  While generating it also checks
the oracle restatement (``oracle/oracle.py``) against the live reference and prints the
deviations, so a regression in the oracle shows up here first.

The fixtures pin semantics the reference itself never pinned (it ships no tests).
"""
from __future__ import annotations

import hashlib
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("CDS_REF_PATH", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)
warnings.filterwarnings("ignore")

import contextlib  # noqa: E402
import io  # noqa: E402

with contextlib.redirect_stdout(io.StringIO()):
    import models.model as ref_model  # noqa: E402
    import models.module as ref_module  # noqa: E402
    import models.dynamic_conv as ref_dyn  # noqa: E402
    from models.utils.warping import homo_warping_3D as ref_warp  # noqa: E402

from cds_mvsnet_b200 import synthetic  # noqa: E402
from oracle import oracle as O  # noqa: E402

torch.set_grad_enabled(False)
torch.manual_seed(0)
T = 0.01


def load_pretrained():
    ck = torch.load(os.path.join(REF, "pretrained/both_dtu_blended/cds_mvsnet.ckpt"), map_location="cpu",
                    weights_only=False)
    sd = O.strip_module_prefix(ck["state_dict"])
    return {k: v.float() if v.is_floating_point() else v for k, v in sd.items() if not k.startswith("refine_network")}


def build_ref(sd, ndepths, ratios):
    with contextlib.redirect_stdout(io.StringIO()):
        m = ref_model.CDSMVSNet(refine=False, ndepths=ndepths, depth_interals_ratio=ratios)
    missing = m.load_state_dict({k: v for k, v in sd.items() if k in m.state_dict()}, strict=False)
    assert not missing.missing_keys, missing.missing_keys
    return m.eval()


def save(name, **arrays):
    out = {k: (v.numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in arrays.items()}
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"  wrote {name}.npz  ({os.path.getsize(path) / 1024:.0f} KiB)")


def dev(tag, a, b):
    d = (a - b).abs().max().item()
    print(f"    oracle vs reference  {tag:<28s} max-abs {d:.3e}   rel-L1 {O.rel_l1(a, b):.3e}")
    return d


def sha(t):
    return hashlib.sha256(t.contiguous().numpy().tobytes()).hexdigest()[:16]


def main():
    sd = load_pretrained()
    save("weights_both_dtu_blended", **{k: v for k, v in sd.items()})
    model3 = build_ref(sd, (48, 32, 8), (4.0, 1.5, 0.75))

    # ---------------------------------------------------------------- A1 warp
    print("A1 homo_warping_3D")
    s = synthetic.make_sample(dict(W=96, H=64, N=3, ndepths=(8,), ratios=(1.0,), B=2, Dtot=192, interval=2.65))
    pm = s.proj_matrices["stage2"]  # half-res intrinsics for a 48x32 feature map
    h, w, C, D = 32, 48, 8, 6
    fea = torch.randn(2, C, h, w)
    ref_P = O.compose_projection(pm[:, 0])
    src_P = O.compose_projection(pm[:, 1])
    dv_planes = torch.linspace(430, 930, D).unsqueeze(0).repeat(2, 1)
    dv_pix = (dv_planes.reshape(2, D, 1, 1) + 40 * torch.rand(2, D, h, w)).contiguous()
    # one far-out-of-frustum plane so zero padding and p_z <= 0 behaviour is pinned as well
    dv_pix[:, 0] = 5.0
    out_planes = ref_warp(fea, src_P, ref_P, dv_planes)
    out_pix = ref_warp(fea, src_P, ref_P, dv_pix)
    dev("warp [B,D]", O.homo_warp(fea, src_P, ref_P, dv_planes), out_planes)
    dev("warp [B,D,h,w]", O.homo_warp(fea, src_P, ref_P, dv_pix), out_pix)
    save("warp", src_fea=fea, src_proj=src_P, ref_proj=ref_P, depth_planes=dv_planes, depth_pix=dv_pix,
         out_planes=out_planes, out_pix=out_pix, cams=pm)

    # ---------------------------------------------------------------- A8 epipoles
    print("A8 compute_Fmatrix / compute_epipole")
    pm3 = s.proj_matrices["stage3"]
    Fm = ref_dyn.compute_Fmatrix(pm3[:, 0], pm3[:, 2])
    e_ref, e_src = ref_dyn.compute_epipole(Fm), ref_dyn.compute_epipole(Fm.transpose(1, 2))
    dev("F", O.fundamental_matrix(pm3[:, 0], pm3[:, 2]), Fm)
    dev("epipole ref", O.epipole_from_F(Fm), e_ref)
    save("epipole", cam_ref=pm3[:, 0], cam_src=pm3[:, 2], F=Fm, e_ref=e_ref, e_src=e_src)

    # ---------------------------------------------------------------- A9 hypotheses
    print("A9 depth hypotheses")
    B, H, W = 2, 32, 48
    dvals = s.depth_values
    dmin, dmax = dvals[:, [0]].unsqueeze(-1).unsqueeze(-1), dvals[:, [-1]].unsqueeze(-1).unsqueeze(-1)
    interval = (dvals[:, 1] - dvals[:, 0]).unsqueeze(-1).unsqueeze(-1)
    cases = {}
    for tag, (D, ratio, scale, prev_scale) in {"s1": (48, 4.0, 4, None), "s2": (32, 1.5, 2, 4), "s3": (8, 0.75, 1, 2)}.items():
        if prev_scale is None:
            cur = dvals
            prev = dvals
        else:
            prev = 425 + 510 * torch.rand(B, H // prev_scale, W // prev_scale)  # spans the clamp at both ends
            cur = torch.nn.functional.interpolate(prev.unsqueeze(1), [H, W], mode="bilinear", align_corners=False).squeeze(1)
        full = ref_module.get_depth_range_samples(cur_depth=cur, ndepth=D, depth_inteval_pixel=ratio * interval,
                                                  dtype=torch.float32, device="cpu", shape=[B, H, W],
                                                  max_depth=dmax, min_depth=dmin)
        smp = torch.nn.functional.interpolate(full.unsqueeze(1), [D, H // scale, W // scale], mode="trilinear",
                                              align_corners=False).squeeze(1)
        mine = O.depth_hypotheses(prev, D, ratio * interval.reshape(B), H, W, dmin.reshape(B), dmax.reshape(B), scale)
        dev(f"hypotheses {tag}", mine, smp)
        cases[f"{tag}_prev"] = prev
        cases[f"{tag}_out"] = smp
    save("hypotheses", depth_values=dvals, **cases)

    # ---------------------------------------------------------------- A5 tail
    print("A5 softmax / depth_regression / conf_regression")
    logits = 3 * torch.randn(2, 8, 16, 24)
    logits[0, :, 0, 0] = torch.tensor([9., 0, 0, 0, 0, 0, 0, 0])      # mass at the first plane
    logits[0, :, 0, 1] = torch.tensor([0., 0, 0, 0, 0, 0, 0, 9])      # mass at the last plane
    dsm = 500 + 300 * torch.rand(2, 8, 16, 24)
    p = torch.softmax(logits, 1)
    depth = ref_module.depth_regression(p, dsm)
    conf = ref_module.conf_regression(p)
    dev("depth", O.depth_regression(p, dsm), depth)
    dev("conf", O.conf_regression(p), conf)
    save("tail", logits=logits, depth_samples=dsm, depth=depth, conf=conf)

    # ---------------------------------------------------------------- A6 DynamicConv
    print("A6 DynamicConv")
    dyn = {}
    for name, cin, hw in (("conv00", 3, (40, 56)), ("conv01", 8, (40, 56)), ("conv10", 16, (24, 32)),
                          ("conv20", 32, (16, 24)), ("out1", 32, (16, 24)), ("out3", 8, (24, 32))):
        x = torch.rand(2, cin, *hw) if cin == 3 else torch.randn(2, cin, *hw)
        epi = torch.tensor([[hw[1] * 1.7, -hw[0] * 0.6], [-30.0, hw[0] / 2.0]])
        mod = getattr(model3.feature, name)
        mod = mod.conv if hasattr(mod, "conv") else mod
        y, nc = mod(x, epipole=epi, temperature=T)
        pre = f"feature.{name}" + ("" if name.startswith("out") else ".conv")
        y2, nc2 = O.dynamic_conv(x, sd, pre, O.FEATURE_DYN_KSIZES[name], epi, T)
        dev(f"{name} out", y2, y)
        dev(f"{name} curv", nc2, nc)
        dyn.update({f"{name}_x": x, f"{name}_epi": epi, f"{name}_y": y, f"{name}_nc": nc})
    save("dynconv", **dyn)

    # ---------------------------------------------------------------- A7 FeatureNet
    print("A7 FeatureNet")
    img = torch.rand(1, 3, 64, 96)
    epi = torch.tensor([[250.0, -40.0]])
    fo = model3.feature(img, epipole=epi, temperature=T)
    fm = O.feature_net(img, sd, epi, T)
    arrays = {"img": img, "epi": epi}
    for st in ("stage1", "stage2", "stage3"):
        for j, nm in enumerate(("fea", "nc_sum", "nc_abs")):
            dev(f"{st} {nm}", fm[st][j], fo[st][j])
            arrays[f"{st}_{nm}"] = fo[st][j]
    save("featurenet", **arrays)

    # ---------------------------------------------------------------- A3 vis-net, A4 CostRegNet
    print("A3 visibility net / A4 CostRegNet")
    arrays = {}
    for st, (C, D, h, w) in enumerate(((32, 16, 16, 24), (16, 8, 16, 16), (8, 8, 24, 32))):
        x2 = torch.cat((2.5 * torch.rand(1, 1, h, w), 0.3 * torch.rand(1, 1, h, w)), 1)
        v = model3.stage_net.vis[st](x2)
        dev(f"vis[{st}]", O.vis_net(x2, sd, f"stage_net.vis.{st}"), v)
        vol = 0.5 * torch.randn(1, C, D, h, w).clamp(-1, 1)
        cr = model3.cost_regularization[st](vol)
        dev(f"costreg[{st}]", O.cost_reg_net(vol, sd, f"cost_regularization.{st}"), cr)
        arrays.update({f"vis{st}_x": x2, f"vis{st}_y": v, f"cr{st}_x": vol, f"cr{st}_y": cr})
    save("nets3d", **arrays)

    # ---------------------------------------------------------------- end to end
    print("A10 end-to-end")
    for tag, cfg, family in (("e2e_cfg1_noise", "cfg1", "noise"),
                             ("e2e_small3_plane", dict(W=160, H=128, N=3, ndepths=(48, 32, 8), ratios=(4.0, 1.5, 0.75),
                                                       B=1, Dtot=192, interval=2.65), "plane"),
                             ("e2e_small3_noise", dict(W=160, H=128, N=4, ndepths=(48, 32, 8), ratios=(4.0, 1.5, 0.75),
                                                       B=2, Dtot=192, interval=2.65), "noise")):
        c = synthetic.CONFIGS[cfg] if isinstance(cfg, str) else cfg
        smp = synthetic.make_sample(cfg, family, seed=0)
        model = build_ref(sd, tuple(c["ndepths"]), tuple(c["ratios"]))
        out = model(smp.imgs, smp.proj_matrices, smp.depth_values, temperature=T)
        mine = O.cdsmvsnet_forward(sd, smp.imgs, smp.proj_matrices, smp.depth_values, c["ndepths"], c["ratios"], T)
        arrays = {"imgs_sha": np.frombuffer(sha(smp.imgs).encode(), dtype=np.uint8),
                  "cfg": np.array([c["W"], c["H"], c["N"], c["B"], c["Dtot"]] + list(c["ndepths"])),
                  "ratios": np.array(c["ratios"]), "interval": np.array(c["interval"])}
        for st in range(len(c["ndepths"])):
            nm = f"stage{st + 1}"
            for key in ("depth", "photometric_confidence", "norm_curv"):
                dev(f"{tag} {nm} {key[:5]}", mine[nm][key], out[nm][key])
                arrays[f"{nm}_{key}"] = out[nm][key]
        if smp.gt_depth is not None:
            err = (out["depth"] - smp.gt_depth).abs().mean().item()
            print(f"    KAT: reference mean |depth - gt| = {err:.3f} mm")
            arrays["gt_depth"] = smp.gt_depth
        save(tag, **arrays)


if __name__ == "__main__":
    main()
