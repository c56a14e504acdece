"""Seeded synthetic inputs for the depth-inference path.

Produces exactly the tensor contract the reference's data layer hands to
``CDSMVSNet.forward`` (reference: datasets/general_eval.py:74,167-200):

* ``imgs``            [B, N, 3, H, W] fp32 in [0, 1]
* ``proj_matrices``   {"stage1".."stage3": [B, N, 2, 4, 4]}; ``[:, i, 0]`` is the 4x4
  world->camera extrinsic, ``[:, i, 1, :3, :3]`` the intrinsic of that stage
  (full-image K / 4 for stage1, x2, x4 for stage2/3)
* ``depth_values``    [B, Dtot] = depth_min + interval * k

Two image families: "noise" (uniform noise, worst case) and "plane" (a textured slanted
plane rendered photo-consistently into every view; ground-truth depth is returned).
Everything here is host-side numpy/torch-CPU and deterministic for a given seed.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch
import torch.nn.functional as F

# Named workloads = BASELINE.json "configs" restated as concrete inputs.
CONFIGS = {
    "cfg1": dict(W=160, H=128, N=3, ndepths=(8,), ratios=(1.0,), B=1, Dtot=192, interval=2.65),
    "cfg2": dict(W=1600, H=1184, N=5, ndepths=(48, 32, 8), ratios=(4.0, 1.5, 0.75), B=1, Dtot=192, interval=2.65),
    "cfg3": dict(W=1920, H=1056, N=7, ndepths=(64, 32, 8), ratios=(4.0, 1.5, 0.75), B=1, Dtot=256,
                 interval=2.65 * 192 / 256),
    "cfg4": dict(W=640, H=512, N=3, ndepths=(48, 32, 8), ratios=(4.0, 1.5, 0.75), B=4, Dtot=192, interval=2.65),
    "cfg5": dict(W=1600, H=1184, N=5, ndepths=(128, 32, 8), ratios=(4.0, 1.5, 0.75), B=1, Dtot=512,
                 interval=2.65 * 192 / 512),
    # the reference's DTU protocol (scripts/dtu_eval.sh: 1536x1152, refine=True): cascade at 768x576 + Refinement network
    "dtu_refine": dict(W=1536, H=1152, N=5, ndepths=(48, 32, 8), ratios=(4.0, 1.5, 0.75), B=1, Dtot=192, interval=2.65,
                       refine=True),
}
DEPTH_MIN = 425.0


@dataclass
class Sample:
    imgs: torch.Tensor
    proj_matrices: dict
    depth_values: torch.Tensor
    gt_depth: torch.Tensor | None = None  # [B, H, W] for the "plane" family


def _rot_y(a: float) -> np.ndarray:
    c, s = math.cos(a), math.sin(a)
    return np.array([[c, 0.0, s], [0.0, 1.0, 0.0], [-s, 0.0, c]], dtype=np.float64)


def make_cameras(n_views: int, H: int, W: int):
    """Full-resolution K (3x3) and N extrinsics (4x4), fp64. View 0 is the reference."""
    K = np.array([[1.2 * W, 0.0, W / 2.0], [0.0, 1.2 * W, H / 2.0], [0.0, 0.0, 1.0]], dtype=np.float64)
    extr = []
    for i in range(n_views):
        s = 0 if i == 0 else (1 if i % 2 == 1 else -1) * math.ceil(i / 2)
        R = _rot_y(-0.06 * s)
        C = np.array([45.0 * s, 12.0 * ((i % 3) - 1), 6.0 * i], dtype=np.float64)
        if i == 0:
            C = np.zeros(3)
        E = np.eye(4, dtype=np.float64)
        E[:3, :3] = R
        E[:3, 3] = -R @ C
        extr.append(E)
    return K, extr


def pack_proj_matrices(K: np.ndarray, extr: list, batch: int, n_stages: int = 3) -> dict:
    """The data layer's [B, N, 2, 4, 4] per-stage dictionary."""
    out = {}
    n = len(extr)
    for s in range(n_stages):
        scale = 0.25 * (2 ** s)
        Ks = K.copy()
        Ks[:2, :] *= scale
        pm = np.zeros((n, 2, 4, 4), dtype=np.float32)
        for i, E in enumerate(extr):
            pm[i, 0] = E.astype(np.float32)
            pm[i, 1, :3, :3] = Ks.astype(np.float32)
        out[f"stage{s + 1}"] = torch.from_numpy(pm).unsqueeze(0).repeat(batch, 1, 1, 1, 1).contiguous()
    return out


def _texture(H: int, W: int, gen: torch.Generator) -> torch.Tensor:
    """Band-limited RGB texture on a 2H x 2W canvas, min-max normalised to [0, 1]."""
    Hc, Wc = 2 * H, 2 * W
    tex = torch.zeros(1, 3, Hc, Wc)
    for cell, amp in ((64, 0.5), (16, 0.3), (4, 0.2), (1, 0.1)):
        h, w = max(2, Hc // cell), max(2, Wc // cell)
        n = torch.rand(1, 3, h, w, generator=gen)
        if (h, w) != (Hc, Wc):
            n = F.interpolate(n, size=(Hc, Wc), mode="bicubic", align_corners=False)
        tex += amp * n
    tex -= tex.amin()
    tex /= tex.amax().clamp_min(1e-12)
    return tex


def _render_plane(tex, K, E_ref, E_i, n, d, H, W):
    """View i of the plane n.X = d (ref camera frame) through the plane homography."""
    # relative pose ref -> view i: X_i = R_rel X_ref + t_rel
    R_ref, t_ref = E_ref[:3, :3], E_ref[:3, 3]
    R_i, t_i = E_i[:3, :3], E_i[:3, 3]
    R_rel = R_i @ R_ref.T
    t_rel = t_i - R_rel @ t_ref
    Hmg = K @ (R_rel + np.outer(t_rel, n) / d) @ np.linalg.inv(K)  # ref pixel -> view-i pixel
    Hinv = np.linalg.inv(Hmg)                                       # view-i pixel -> ref pixel
    ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float64), torch.arange(W, dtype=torch.float64), indexing="ij")
    p = torch.stack((xs, ys, torch.ones_like(xs)), 0).reshape(3, -1)
    q = torch.from_numpy(Hinv) @ p
    u, v = q[0] / q[2], q[1] / q[2]
    # texture canvas is 2H x 2W, the ref image occupies its centre
    Hc, Wc = tex.shape[-2:]
    uc, vc = u + W / 2.0, v + H / 2.0
    gx = (uc / (Wc - 1) * 2 - 1).float()
    gy = (vc / (Hc - 1) * 2 - 1).float()
    grid = torch.stack((gx, gy), -1).reshape(1, H, W, 2)
    return F.grid_sample(tex, grid, mode="bilinear", padding_mode="border", align_corners=True)[0]


def make_sample(cfg: str | dict = "cfg1", family: str = "noise", seed: int = 0, **over) -> Sample:
    c = dict(CONFIGS[cfg]) if isinstance(cfg, str) else dict(cfg)
    c.update(over)
    H, W, N, B = c["H"], c["W"], c["N"], c["B"]
    K, extr = make_cameras(N, H, W)
    n_stages = 3
    # refine=True (the reference's DTU protocol, dtu_eval.sh): the cascade works at half the image resolution on pixels (2i, 2j),
    # so the cameras handed to the model are the full-resolution ones with the first two rows of K halved
    Kc = K.copy()
    if c.get("refine"):
        Kc[:2, :] *= 0.5
    proj = pack_proj_matrices(Kc, extr, B, n_stages)
    dv = (DEPTH_MIN + c["interval"] * torch.arange(c["Dtot"], dtype=torch.float32)).unsqueeze(0).repeat(B, 1)
    gen = torch.Generator().manual_seed(seed)
    gt = None
    if family == "noise":
        imgs = torch.rand(B, N, 3, H, W, generator=gen)
    elif family == "plane":
        nrm = np.array([0.15, -0.1, 1.0])
        nrm /= np.linalg.norm(nrm)
        Kinv = np.linalg.inv(K)
        # depth 650 at the principal ray: n . (K^-1 c) * 650 = d
        d = 650.0 * float(nrm @ (Kinv @ np.array([W / 2.0, H / 2.0, 1.0])))
        views = []
        for b in range(B):
            tex = _texture(H, W, gen)
            views.append(torch.stack([_render_plane(tex, K, extr[0], extr[i], nrm, d, H, W) for i in range(N)], 0))
        imgs = torch.stack(views, 0).clamp(0, 1).float()
        ys, xs = np.meshgrid(np.arange(H, dtype=np.float64), np.arange(W, dtype=np.float64), indexing="ij")
        rays = Kinv @ np.stack((xs.ravel(), ys.ravel(), np.ones(H * W)))
        gt = torch.from_numpy((d / (nrm @ rays)).reshape(H, W)).float().unsqueeze(0).repeat(B, 1, 1)
    else:
        raise ValueError(f"unknown image family {family!r}")
    return Sample(imgs.contiguous(), proj, dv.contiguous(), gt)


def model_args(cfg: str | dict) -> dict:
    c = CONFIGS[cfg] if isinstance(cfg, str) else cfg
    return dict(refine=False, ndepths=tuple(c["ndepths"]), depth_interals_ratio=tuple(c["ratios"]),
                share_cr=False, cr_base_chs=(8,) * len(c["ndepths"]), grad_method="detach")


# --------------------------------------------------------------------------------------------
# Random weights with the reference's state-dict keys/shapes (reference: models/model.py:98-138,
# models/module.py:201-315, models/dynamic_conv.py:81-95).  Used when pretrained weights are not
# at hand (GPU box); BN running stats are randomised so BN folding is exercised.
# --------------------------------------------------------------------------------------------
_FEATURE_DYN = {  # name: (cin, cout, ksizes, branch bias)
    "conv00": (3, 8, (3, 7, 11), False), "conv01": (8, 8, (3, 5, 7), False),
    "conv10": (16, 16, (3, 5), False), "conv11": (16, 16, (3, 5), False),
    "conv20": (32, 32, (1, 3), False), "conv21": (32, 32, (1, 3), False),
    "out1": (32, 32, (1, 3), True), "out2": (16, 16, (1, 3), True), "out3": (8, 8, (1, 3), True),
}
_FEATURE_PLAIN = {"downsample1": (8, 16, 3), "downsample2": (16, 32, 3), "inner1": (48, 16, 1), "inner2": (24, 8, 1)}


def feature_layer_specs():
    return dict(_FEATURE_DYN), dict(_FEATURE_PLAIN)


def random_state_dict(ndepths=(48, 32, 8), seed: int = 123, cr_base: int = 8) -> dict:
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def conv_w(co, ci, *k):
        fan = ci * int(np.prod(k))
        return (torch.rand(co, ci, *k, generator=g) * 2 - 1) * math.sqrt(3.0 / fan)

    def bn(prefix, c):
        sd[prefix + ".weight"] = 0.5 + torch.rand(c, generator=g)
        sd[prefix + ".bias"] = 0.2 * torch.randn(c, generator=g)
        sd[prefix + ".running_mean"] = 0.1 * torch.randn(c, generator=g)
        sd[prefix + ".running_var"] = 0.5 + torch.rand(c, generator=g)
        sd[prefix + ".num_batches_tracked"] = torch.tensor(1, dtype=torch.long)

    for name, (ci, co, ks, bias) in _FEATURE_DYN.items():
        p = f"feature.{name}" + ("" if name.startswith("out") else ".conv")
        for i, k in enumerate(ks):
            sd[f"{p}.att_convs.{i}.weight"] = 0.1 * torch.randn(3, ci, k, k, generator=g)
            sd[f"{p}.convs.{i}.weight"] = conv_w(co, ci, k, k)
            if bias:
                sd[f"{p}.convs.{i}.bias"] = 0.1 * torch.randn(co, generator=g)
        sd[f"{p}.att_weights.0.weight"] = conv_w(4, len(ks), 1, 1)
        bn(f"{p}.att_weights.1", 4)
        sd[f"{p}.att_weights.3.weight"] = conv_w(len(ks), 4, 1, 1)
    for name, (ci, co, k) in _FEATURE_PLAIN.items():
        sd[f"feature.{name}.conv.weight"] = conv_w(co, ci, k, k)
    for s in range(len(ndepths)):
        v = f"stage_net.vis.{s}"
        for j, (ci, co) in enumerate(((2, 16), (16, 16), (16, 16))):
            sd[f"{v}.{j}.conv.weight"] = conv_w(co, ci, 3, 3)
            bn(f"{v}.{j}.bn", co)
        sd[f"{v}.3.weight"] = conv_w(1, 16, 1, 1)
        sd[f"{v}.3.bias"] = 0.1 * torch.randn(1, generator=g)
        cr = f"cost_regularization.{s}"
        cin = (32, 16, 8)[s]
        b = cr_base
        for name, (ci, co) in {"conv0": (cin, b), "conv1": (b, 2 * b), "conv2": (2 * b, 2 * b), "conv3": (2 * b, 4 * b),
                               "conv4": (4 * b, 4 * b), "conv5": (4 * b, 8 * b), "conv6": (8 * b, 8 * b)}.items():
            sd[f"{cr}.{name}.conv.weight"] = conv_w(co, ci, 3, 3, 3)
            bn(f"{cr}.{name}.bn", co)
        for name, (ci, co) in {"conv7": (8 * b, 4 * b), "conv9": (4 * b, 2 * b), "conv11": (2 * b, b)}.items():
            sd[f"{cr}.{name}.conv.weight"] = conv_w(ci, co, 3, 3, 3) * math.sqrt(8.0)  # ConvTranspose3d layout [Cin,Cout,...]
            bn(f"{cr}.{name}.bn", co)
        sd[f"{cr}.prob.weight"] = conv_w(1, b, 3, 3, 3)
    return sd


def make_fusion_sample(H: int, W: int, n_src: int, seed: int = 0, batch: int = 1):
    """Inputs of the geometric-consistency filter (reference fusion.py / test.py:326-352) for the synthetic rig: per-view depth
    maps of the slanted plane of the "plane" family (analytic, in every camera), perturbed so that the masks are mixed
    (Gaussian noise on all maps, gross outliers and zeroed pixels on patches), cams [.,2,4,4] = (extrinsic, K with [1,3,3]=1)
    and 3-channel confidences.  Returns dict of fp32 tensors shaped as test.py's TTDataset yields them."""
    rng = np.random.default_rng(seed)
    K, extr = make_cameras(n_src + 1, H, W)
    nrm = np.array([0.15, -0.1, 1.0]); nrm /= np.linalg.norm(nrm)
    d0 = 650.0 * nrm[2]                       # plane n.X = d0 in the reference (= world) frame, 650 mm at the principal ray
    Kinv = np.linalg.inv(K)
    ys, xs = np.meshgrid(np.arange(H) + 0.5, np.arange(W) + 0.5, indexing="ij")
    rays = np.stack([xs, ys, np.ones_like(xs)], -1) @ Kinv.T          # [H, W, 3], z = 1
    depths, cams = [], []
    for i, E in enumerate(extr):
        R, t = E[:3, :3], E[:3, 3]
        C = -R.T @ t
        # X_w = R^T (depth * ray) + C  ->  n.X_w = d0
        dep = (d0 - nrm @ C) / (rays @ (R @ nrm))
        dep = dep + rng.normal(0.0, 0.4, dep.shape)
        y0, x0 = rng.integers(0, H // 2), rng.integers(0, W // 2)
        dep[y0:y0 + H // 6, x0:x0 + W // 6] *= 1.0 + 0.05 * rng.random()          # a patch that fails the depth test
        y1, x1 = rng.integers(0, H // 2), rng.integers(0, W // 2)
        dep[y1:y1 + H // 8, x1:x1 + W // 8] = 0.0                                   # a patch masked out by the confidence filter
        depths.append(dep.astype(np.float32))
        cam = np.zeros((2, 4, 4), dtype=np.float32)
        cam[0] = E.astype(np.float32)
        cam[1, :3, :3] = K.astype(np.float32)
        cam[1, 3, 3] = 1.0
        cams.append(cam)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    rep = lambda x: x.unsqueeze(0).repeat(batch, *([1] * x.dim())).contiguous()
    conf = rng.random((n_src + 1, 3, H, W)).astype(np.float32)
    return {"ref_depth": rep(t(depths[0])[None]), "src_depths": rep(t(np.stack(depths[1:]))[:, None]),
            "ref_cam": rep(t(cams[0])), "src_cams": rep(t(np.stack(cams[1:]))),
            "ref_conf": rep(t(conf[0])), "src_confs": rep(t(conf[1:]))}
