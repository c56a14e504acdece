"""CPU oracle for the CDS-MVSNet depth-inference path.  TEST INFRASTRUCTURE ONLY.

This file is the checker, never the product: only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.  The shipped path
(``cds_mvsnet_b200``) never routes through it and fails loudly without its CUDA library.

It restates, function by function, what the reference computes (all file:line citations are
into the reference tree, TruongKhang/cds-mvsnet @ 2a84f7a):

* plane-sweep warp ................ models/utils/warping.py:69-104
* per-view cost-volume loop ....... models/model.py:34-60,74,79
* visibility net .................. models/model.py:14, models/module.py:169-198
* 3-D regulariser ................. models/module.py:80-166,270-315
* soft-argmin + confidence ........ models/model.py:85-92, models/module.py:373-391
* depth hypotheses ................ models/model.py:141-143,174-193, models/module.py:394-439
* DynamicConv / FeatureNet ........ models/dynamic_conv.py:81-122, models/module.py:28-71,201-267
* fundamental matrix / epipoles ... models/dynamic_conv.py:7-47
* cascade driver .................. models/model.py:140-223
* backward of warp / soft-argmin .. models/utils/warping.py:79,100-101, models/module.py:373-379 (what autograd derives)
* training loss ................... models/losses.py:6-48

The dense arithmetic the reference delegates to PyTorch 1.6 ATen/cuDNN (conv2d/conv3d/
conv_transpose3d, softmax, instance/batch norm, linalg inverse; third-party, not vendored in
the reference tree) is delegated to the same operators of the torch build in this image
(CPU, fp32); their published definitions are additionally restated in plain C in
``oracle/c/cds_oracle.c`` and cross-checked in ``tests/test_oracle_c.py``.  Everything the
reference builds on top of those operators (the warp's coordinate maths and bilinear gather,
hypothesis generation, confidence window, curvature gate, aggregation) is written out here
explicitly rather than through ``grid_sample`` / ``avg_pool3d`` / ``interpolate``.

Parity pin: the reference ships no tests or golden vectors, so the oracle
is pinned against outputs of the reference itself, imported unmodified in the build container:
``tests/golden/make_golden.py`` generated ``tests/golden/*.npz`` and
``tests/test_oracle_golden.py`` checks every function here against them.

The model is addressed functionally through a state dict with the reference's own key names
(``feature.conv00.conv.att_convs.0.weight`` ...), so pretrained checkpoints drop in.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

BN_EPS = 1e-5
IN_EPS = 1e-5

FEATURE_DYN_KSIZES = {
    "conv00": (3, 7, 11), "conv01": (3, 5, 7), "conv10": (3, 5), "conv11": (3, 5),
    "conv20": (1, 3), "conv21": (1, 3), "out1": (1, 3), "out2": (1, 3), "out3": (1, 3),
}


# ----------------------------------------------------------------------------------------------
# geometry
# ----------------------------------------------------------------------------------------------
def compose_projection(cam: torch.Tensor) -> torch.Tensor:
    """[B,2,4,4] (extrinsic, intrinsic) -> [B,4,4] with rows 0-2 = K @ E[:3,:4] (model.py:40-43)."""
    P = cam[:, 0].clone()
    P[:, :3, :4] = cam[:, 1, :3, :3] @ cam[:, 0, :3, :4]
    return P


def _skew(v: torch.Tensor) -> torch.Tensor:
    z = torch.zeros_like(v[:, 0])
    return torch.stack((torch.stack((z, -v[:, 2], v[:, 1]), 1),
                        torch.stack((v[:, 2], z, -v[:, 0]), 1),
                        torch.stack((-v[:, 1], v[:, 0], z), 1)), 1)


def fundamental_matrix(cam1: torch.Tensor, cam2: torch.Tensor) -> torch.Tensor:
    """F = [K2 R2 (c1 - c2)]x  K2 R2 (K1 R1)^-1   (dynamic_conv.py:19-38)."""
    K1, R1, t1 = cam1[:, 1, :3, :3], cam1[:, 0, :3, :3], cam1[:, 0, :3, 3:4]
    K2, R2, t2 = cam2[:, 1, :3, :3], cam2[:, 0, :3, :3], cam2[:, 0, :3, 3:4]
    c1 = -torch.linalg.inv(R1) @ t1
    c2 = -torch.linalg.inv(R2) @ t2
    P1, P2 = K1 @ R1, K2 @ R2
    e = (P2 @ (c1 - c2)).squeeze(2)
    return _skew(e) @ P2 @ torch.linalg.inv(P1)


def epipole_from_F(Fm: torch.Tensor) -> torch.Tensor:
    """2x2 solve of rows c*F0 +/- (F1+F2), c = 1e3 (dynamic_conv.py:41-47)."""
    c = 1e3
    r1 = c * Fm[:, 0] + Fm[:, 1] + Fm[:, 2]
    r2 = c * Fm[:, 0] - Fm[:, 1] - Fm[:, 2]
    A = torch.stack((r1, r2), 1)
    return (-torch.linalg.inv(A[:, :, :2]) @ A[:, :, 2:3]).squeeze(2)


def warp_coefficients(src_proj: torch.Tensor, ref_proj: torch.Tensor):
    """rot [B,3,3], trans [B,3] of src_proj @ inv(ref_proj) (warping.py:80-82)."""
    M = src_proj @ torch.linalg.inv(ref_proj)
    return M[:, :3, :3], M[:, :3, 3]


def bilinear_gather_zeros(fea: torch.Tensor, u: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
    """Sample fea [B,C,h,w] at pixel coords (u,v) [B,M]; taps outside the image contribute 0.

    Restates grid_sample(bilinear, zeros, align_corners=True) after the reference's
    normalise step (warping.py:95-101), i.e. sampling at the pixel coordinates themselves.
    """
    B, C, h, w = fea.shape
    # reproduce the normalise -> unnormalise round trip the reference goes through in fp32
    un = u / ((w - 1) / 2) - 1
    vn = v / ((h - 1) / 2) - 1
    u = (un + 1) / 2 * (w - 1)
    v = (vn + 1) / 2 * (h - 1)
    x0 = torch.floor(u)
    y0 = torch.floor(v)
    fx, fy = u - x0, v - y0
    flat = fea.reshape(B, C, h * w)
    out = torch.zeros(B, C, u.shape[1], dtype=fea.dtype)
    for dy, dx, wgt in ((0, 0, (1 - fx) * (1 - fy)), (0, 1, fx * (1 - fy)), (1, 0, (1 - fx) * fy), (1, 1, fx * fy)):
        xi, yi = x0 + dx, y0 + dy
        ok = (xi >= 0) & (xi <= w - 1) & (yi >= 0) & (yi <= h - 1)
        idx = (yi.clamp(0, h - 1) * w + xi.clamp(0, w - 1)).long()
        tap = torch.gather(flat, 2, idx.unsqueeze(1).expand(B, C, -1))
        out += tap * (wgt * ok).unsqueeze(1)
    return out


# bench.py's CPU legs set this: the gather then goes through ATen's grid_sample exactly as the reference's
# own code does (warping.py:100-101), so the timed port costs what the reference costs.  Tests keep the
# explicit restatement above (the two agree to 5e-7, tests/test_oracle_golden.py::test_fast_gather).
FAST_GATHER = False


def homo_warp(src_fea, src_proj, ref_proj, depth_values):
    """homo_warping_3D (warping.py:69-104): [B,C,h,w] -> [B,C,D,h,w]."""
    B, C, h, w = src_fea.shape
    D = depth_values.shape[1]
    rot, trans = warp_coefficients(src_proj, ref_proj)
    ys, xs = torch.meshgrid(torch.arange(h, dtype=torch.float32), torch.arange(w, dtype=torch.float32), indexing="ij")
    pix = torch.stack((xs.reshape(-1), ys.reshape(-1), torch.ones(h * w)), 0)       # [3, P]
    ray = rot @ pix.unsqueeze(0)                                                    # [B,3,P]
    dep = depth_values.reshape(B, 1, D, -1)                                         # [B,1,D,1|P]
    p = ray.unsqueeze(2) * dep + trans.reshape(B, 3, 1, 1)                          # [B,3,D,P]
    z = p[:, 2] + 1e-6
    if FAST_GATHER:
        grid = torch.stack(((p[:, 0] / z) / ((w - 1) / 2) - 1, (p[:, 1] / z) / ((h - 1) / 2) - 1), dim=3)
        out = F.grid_sample(src_fea, grid.reshape(B, D * h, w, 2), mode="bilinear", padding_mode="zeros", align_corners=True)
        return out.reshape(B, C, D, h, w)
    u = (p[:, 0] / z).reshape(B, -1)
    v = (p[:, 1] / z).reshape(B, -1)
    return bilinear_gather_zeros(src_fea, u, v).reshape(B, C, D, h, w)


def bilinear_scatter_zeros(grad: torch.Tensor, u: torch.Tensor, v: torch.Tensor, h: int, w: int) -> torch.Tensor:
    """Adjoint of ``bilinear_gather_zeros`` in ``fea``: grad [B,C,M] at coords (u,v) [B,M] -> [B,C,h,w].

    What autograd does for the reference's grid_sample when only ``src_fea`` carries a gradient (the grid is built
    under no_grad, warping.py:79): every sample adds weight * grad to its in-image taps, out-of-image taps are dropped.
    """
    B, C, M = grad.shape
    un = u / ((w - 1) / 2) - 1
    vn = v / ((h - 1) / 2) - 1
    u = (un + 1) / 2 * (w - 1)
    v = (vn + 1) / 2 * (h - 1)
    x0 = torch.floor(u)
    y0 = torch.floor(v)
    fx, fy = u - x0, v - y0
    out = torch.zeros(B, C, h * w, dtype=torch.float64)
    for dy, dx, wgt in ((0, 0, (1 - fx) * (1 - fy)), (0, 1, fx * (1 - fy)), (1, 0, (1 - fx) * fy), (1, 1, fx * fy)):
        xi, yi = x0 + dx, y0 + dy
        ok = (xi >= 0) & (xi <= w - 1) & (yi >= 0) & (yi <= h - 1)
        idx = (yi.clamp(0, h - 1) * w + xi.clamp(0, w - 1)).long()
        contrib = (grad * (wgt * ok).unsqueeze(1)).double()   # fp64 accumulation: the order of the adds is not the reference's
        out.scatter_add_(2, idx.unsqueeze(1).expand(B, C, -1), contrib)
    return out.reshape(B, C, h, w).float()


def homo_warp_backward(grad_out, src_proj, ref_proj, depth_values):
    """d(loss)/d(src_fea) of homo_warping_3D given d(loss)/d(out) [B,C,D,h,w] (warping.py:79: coordinates carry no gradient)."""
    B, C, D, h, w = grad_out.shape
    rot, trans = warp_coefficients(src_proj, ref_proj)
    ys, xs = torch.meshgrid(torch.arange(h, dtype=torch.float32), torch.arange(w, dtype=torch.float32), indexing="ij")
    pix = torch.stack((xs.reshape(-1), ys.reshape(-1), torch.ones(h * w)), 0)
    ray = rot @ pix.unsqueeze(0)
    dep = depth_values.reshape(B, 1, D, -1)
    p = ray.unsqueeze(2) * dep + trans.reshape(B, 3, 1, 1)
    z = p[:, 2] + 1e-6
    u = (p[:, 0] / z).reshape(B, -1)
    v = (p[:, 1] / z).reshape(B, -1)
    return bilinear_scatter_zeros(grad_out.reshape(B, C, -1), u, v, h, w)


# ----------------------------------------------------------------------------------------------
# hypotheses (A9)
# ----------------------------------------------------------------------------------------------
def upsample_bilinear_half_pixel(x: torch.Tensor, H: int, W: int) -> torch.Tensor:
    """F.interpolate(bilinear, align_corners=False) written out.  x [B,h,w] -> [B,H,W]."""
    B, h, w = x.shape

    def axis(n_out, n_in):
        c = (torch.arange(n_out, dtype=torch.float32) + 0.5) * (n_in / n_out) - 0.5
        c = c.clamp_min(0)
        i0 = c.floor().long().clamp_max(n_in - 1)
        i1 = (i0 + 1).clamp_max(n_in - 1)
        f = c - i0.float()
        return i0, i1, f

    y0, y1, fy = axis(H, h)
    x0, x1, fx = axis(W, w)
    top = x[:, y0][:, :, x0] * (1 - fx) + x[:, y0][:, :, x1] * fx
    bot = x[:, y1][:, :, x0] * (1 - fx) + x[:, y1][:, :, x1] * fx
    return top * (1 - fy).unsqueeze(1) + bot * fy.unsqueeze(1)


def depth_hypotheses(cur_depth, ndepth, interval_pixel, H, W, dmin, dmax, scale):
    """Stage hypotheses at the stage resolution [B, D, H/scale, W/scale].

    cur_depth [B,Dtot] (first stage): D planes uniformly spanning [dv[0], dv[-1]]
    (module.py:425-433).  cur_depth [B,h,w] (later stages): bilinear up-sample to (H,W)
    (model.py:177-182), ``cur - ((D-1)//2)*step + d*step`` clamped to [dmin,dmax]
    (module.py:398-417), then the trilinear resize to the stage grid (model.py:191-193), which
    is the identity along D and, per image axis, the mean of the two centre taps of each
    scale-wide block (a 2x2 box mean for scale 2).
    interval_pixel, dmin, dmax: [B] tensors.
    """
    B = cur_depth.shape[0]
    k = torch.arange(ndepth, dtype=torch.float32)
    if cur_depth.dim() == 2:
        lo, hi = cur_depth[:, 0], cur_depth[:, -1]
        step = (hi - lo) / (ndepth - 1)
        planes = lo.unsqueeze(1) + k.unsqueeze(0) * step.unsqueeze(1)
        full = planes.reshape(B, ndepth, 1, 1).expand(B, ndepth, H, W)
    else:
        cur = upsample_bilinear_half_pixel(cur_depth, H, W)
        nl = (ndepth - 1) // 2
        step = interval_pixel.reshape(B, 1, 1)
        start = cur - nl * step
        full = start.unsqueeze(1) + k.reshape(1, -1, 1, 1) * (torch.ones_like(cur) * step).unsqueeze(1)
        lo = dmin.reshape(B, 1, 1, 1)
        hi = dmax.reshape(B, 1, 1, 1)
        full = lo + (full - lo).clamp(min=0)
        full = hi + (full - hi).clamp(max=0)
    if scale == 1:
        return full.contiguous()
    # half-pixel-centre resize by an even integer factor s: output i samples input s*i + (s-1)/2,
    # i.e. the mean of the two centre taps s*i + s/2 - 1 and s*i + s/2 (per image axis)
    a, b = scale // 2 - 1, scale // 2
    rows = 0.5 * full[:, :, a::scale] + 0.5 * full[:, :, b::scale]
    return (0.5 * rows[:, :, :, a::scale] + 0.5 * rows[:, :, :, b::scale]).contiguous()


# ----------------------------------------------------------------------------------------------
# soft-argmin tail (A5)
# ----------------------------------------------------------------------------------------------
def depth_regression(p, depth_values):
    if depth_values.dim() <= 2:
        depth_values = depth_values.reshape(*depth_values.shape, 1, 1)
    return (p * depth_values).sum(1)


def depth_regression_backward(grad_depth, p, depth_values):
    """(d/dp, d/d depth_values) of sum_d p_d * depth_d (module.py:373-379) given d(loss)/d(depth) [B,h,w]."""
    dv = depth_values.reshape(*depth_values.shape, 1, 1) if depth_values.dim() <= 2 else depth_values
    g = grad_depth.unsqueeze(1)
    grad_p = g * dv.expand_as(p)
    grad_dv = g * p
    if depth_values.dim() <= 2:
        grad_dv = grad_dv.sum((2, 3)).reshape(depth_values.shape)
    return grad_p, grad_dv


def conf_regression(p, n: int = 4):
    """p[d-1]+p[d]+p[d+1]+p[d+2] at d = clamp(trunc(sum_k k*p_k)) (module.py:382-391)."""
    B, D, h, w = p.shape
    lead, trail = n // 2 - 1, n // 2
    padded = torch.cat((torch.zeros(B, lead, h, w), p, torch.zeros(B, trail, h, w)), 1)
    window = sum(padded[:, j:j + D] for j in range(n))
    idx = depth_regression(p, torch.arange(D, dtype=torch.float32)).long().clamp(0, D - 1)
    return torch.gather(window, 1, idx.unsqueeze(1)).squeeze(1)


# ----------------------------------------------------------------------------------------------
# small nets
# ----------------------------------------------------------------------------------------------
def _bn_eval(x, sd, prefix):
    shape = [1, -1] + [1] * (x.dim() - 2)
    g, b = sd[prefix + ".weight"], sd[prefix + ".bias"]
    m, v = sd[prefix + ".running_mean"], sd[prefix + ".running_var"]
    return (x - m.reshape(shape)) / torch.sqrt(v.reshape(shape) + BN_EPS) * g.reshape(shape) + b.reshape(shape)


def instance_norm(x):
    m = x.mean(dim=(2, 3), keepdim=True)
    v = ((x - m) ** 2).mean(dim=(2, 3), keepdim=True)
    return (x - m) / torch.sqrt(v + IN_EPS)


def vis_net(x, sd, prefix):
    """2->16->16->16 (3x3 conv, BN, ReLU) -> 1x1 conv + bias -> sigmoid (model.py:14)."""
    for j in range(3):
        x = F.relu(_bn_eval(F.conv2d(x, sd[f"{prefix}.{j}.conv.weight"], padding=1), sd, f"{prefix}.{j}.bn"))
    return torch.sigmoid(F.conv2d(x, sd[f"{prefix}.3.weight"], sd[f"{prefix}.3.bias"]))


def dynamic_conv(x, sd, prefix, ksizes, epipole, temperature):
    """DynamicConv.forward (dynamic_conv.py:97-122) -> (blended [B,Co,H,W], norm_curv [B,1,H,W])."""
    B, _, H, W = x.shape
    ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    u = xs.reshape(1, 1, H, W) - epipole[:, 0].reshape(B, 1, 1, 1)
    v = ys.reshape(1, 1, H, W) - epipole[:, 1].reshape(B, 1, 1, 1)
    r = torch.sqrt(u * u + v * v) + 1e-6
    u, v = u / r, v / r
    quad = torch.cat((u * u, 2 * u * v, v * v), 1)
    curvs, branches = [], []
    for i, k in enumerate(ksizes):
        abc = F.conv2d(x, sd[f"{prefix}.att_convs.{i}.weight"], padding=(k - 1) // 2)
        curvs.append((abc * quad).sum(1, keepdim=True))
        branches.append(F.conv2d(x, sd[f"{prefix}.convs.{i}.weight"], sd.get(f"{prefix}.convs.{i}.bias"),
                                 padding=(k - 1) // 2))
    curv = torch.cat(curvs, 1)
    g = F.conv2d(curv, sd[f"{prefix}.att_weights.0.weight"])
    g = F.relu(_bn_eval(g, sd, f"{prefix}.att_weights.1"))
    g = F.conv2d(g, sd[f"{prefix}.att_weights.3.weight"])
    wgt = torch.softmax(g / temperature, dim=1)
    out = sum(branches[i] * wgt[:, i:i + 1] for i in range(len(ksizes)))
    return out, (curv * wgt).sum(1, keepdim=True)


def _dyn_block(x, sd, name, epipole, T):
    y, nc = dynamic_conv(x, sd, f"feature.{name}.conv", FEATURE_DYN_KSIZES[name], epipole, T)
    return F.leaky_relu(instance_norm(y), 0.1), nc


def _plain_block(x, sd, name, stride=1, padding=0):
    y = F.conv2d(x, sd[f"feature.{name}.conv.weight"], stride=stride, padding=padding)
    return F.leaky_relu(instance_norm(y), 0.1)


def _up2_nearest(x):
    return x.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)


def feature_net(img, sd, epipole, T):
    """FeatureNet.forward (module.py:236-267)."""
    e1, e2 = epipole / 2, epipole / 4
    c00, n00 = _dyn_block(img, sd, "conv00", epipole, T)
    c01, n01 = _dyn_block(c00, sd, "conv01", epipole, T)
    d1 = _plain_block(c01, sd, "downsample1", 2, 1)
    c10, n10 = _dyn_block(d1, sd, "conv10", e1, T)
    c11, n11 = _dyn_block(c10, sd, "conv11", e1, T)
    d2 = _plain_block(c11, sd, "downsample2", 2, 1)
    c20, n20 = _dyn_block(d2, sd, "conv20", e2, T)
    c21, n21 = _dyn_block(c20, sd, "conv21", e2, T)

    out = {}
    o1, n22 = dynamic_conv(c21, sd, "feature.out1", FEATURE_DYN_KSIZES["out1"], e2, T)
    o1 = torch.tanh(instance_norm(o1))
    out["stage1"] = (o1, (n20 ** 2 + n21 ** 2 + n22 ** 2) / 3, n22.abs())

    i1 = _plain_block(torch.cat((_up2_nearest(c21), c11), 1), sd, "inner1")
    o2, n12 = dynamic_conv(i1, sd, "feature.out2", FEATURE_DYN_KSIZES["out2"], e1, T)
    o2 = torch.tanh(instance_norm(o2))
    out["stage2"] = (o2, (n10 ** 2 + n11 ** 2 + n12 ** 2) / 3, n12.abs())

    i2 = _plain_block(torch.cat((_up2_nearest(o2), c01), 1), sd, "inner2")
    o3, n02 = dynamic_conv(i2, sd, "feature.out3", FEATURE_DYN_KSIZES["out3"], epipole, T)
    o3 = torch.tanh(instance_norm(o3))
    out["stage3"] = (o3, (n00 ** 2 + n01 ** 2 + n02 ** 2) / 3, n02.abs())
    return out


def _cbr3d(x, sd, p, stride=1):
    return F.relu(_bn_eval(F.conv3d(x, sd[p + ".conv.weight"], stride=stride, padding=1), sd, p + ".bn"))


def _dbr3d(x, sd, p):
    y = F.conv_transpose3d(x, sd[p + ".conv.weight"], stride=2, padding=1, output_padding=1)
    return F.relu(_bn_eval(y, sd, p + ".bn"))


def cost_reg_net(x, sd, prefix):
    """CostRegNet.forward (module.py:305-315): [B,C,D,h,w] -> [B,1,D,h,w]."""
    c0 = _cbr3d(x, sd, prefix + ".conv0")
    c2 = _cbr3d(_cbr3d(c0, sd, prefix + ".conv1", 2), sd, prefix + ".conv2")
    c4 = _cbr3d(_cbr3d(c2, sd, prefix + ".conv3", 2), sd, prefix + ".conv4")
    y = _cbr3d(_cbr3d(c4, sd, prefix + ".conv5", 2), sd, prefix + ".conv6")
    y = c4 + _dbr3d(y, sd, prefix + ".conv7")
    y = c2 + _dbr3d(y, sd, prefix + ".conv9")
    y = c0 + _dbr3d(y, sd, prefix + ".conv11")
    return F.conv3d(y, sd[prefix + ".prob.weight"], padding=1)


# ----------------------------------------------------------------------------------------------
# stage + cascade
# ----------------------------------------------------------------------------------------------
def similarity_entropy(ref_fea, warped):
    """entropy of softmax_D(sum_c ref*warped) (model.py:46-50) -> (in_prod, entropy [B,1,h,w])."""
    prod = ref_fea.unsqueeze(2) * warped
    sim = prod.sum(1)
    p = torch.softmax(sim, dim=1)
    return prod, -(p * torch.log(p)).sum(1, keepdim=True)


def stage_net(features, proj_matrices, depth_samples, sd, stage_idx, return_intermediates=False):
    """StageNet.forward, eval branch (model.py:16-94)."""
    cams = torch.unbind(proj_matrices, 1)
    ref_P = compose_projection(cams[0])
    vol, vis_sum, nc_sum = 0.0, 0.0, 0.0
    inter = {"entropy": [], "vis": []}
    for feat, cam in zip(features, cams[1:]):
        ref_fea, ref_nc_sum, ref_nc = feat["ref"]
        src_fea, src_nc_sum, _ = feat["src"]
        warped = homo_warp(src_fea, compose_projection(cam), ref_P, depth_samples)
        prod, ent = similarity_entropy(ref_fea, warped)
        vis = vis_net(torch.cat((ent, ref_nc), 1), sd, f"stage_net.vis.{stage_idx}")
        vol = vol + prod * vis.unsqueeze(1)
        vis_sum = vis_sum + vis
        nc_sum = nc_sum + (ref_nc_sum + src_nc_sum) / 2
        inter["entropy"].append(ent)
        inter["vis"].append(vis)
    volume = vol / (vis_sum.unsqueeze(1) + 1e-6)
    nc_mean = nc_sum / len(features)
    logits = cost_reg_net(volume, sd, f"cost_regularization.{stage_idx}").squeeze(1)
    p = torch.softmax(logits, dim=1)
    out = {"depth": depth_regression(p, depth_samples), "photometric_confidence": conf_regression(p),
           "norm_curv": nc_mean}
    if return_intermediates:
        inter.update(volume=volume, logits=logits, prob=p)
        out["_inter"] = inter
    return out


def _cbr2d(x, sd, p):
    """ConvBnReLU (module.py:169-198): 3x3 conv, no bias -> BatchNorm2d (eval) -> ReLU."""
    return F.relu(_bn_eval(F.conv2d(x, sd[p + ".conv.weight"], padding=1), sd, p + ".bn"))


def refinement(sd, img, depth_0, depth_min, depth_max, prefix="refine_network"):
    """Refinement.forward (module.py:337-370): img [B,3,H,W], depth_0 [B,1,H/2,W/2], depth_min/max [B] -> [B,1,H,W]."""
    B = depth_min.shape[0]
    lo, hi = depth_min.view(B, 1, 1, 1), depth_max.view(B, 1, 1, 1)
    depth = (depth_0 - lo) / (hi - lo) * 10
    conv0 = _cbr2d(img, sd, prefix + ".conv0")
    d = _cbr2d(_cbr2d(depth, sd, prefix + ".conv1"), sd, prefix + ".conv2")
    d = F.conv_transpose2d(d, sd[prefix + ".deconv.weight"], stride=2, padding=1, output_padding=1)
    d = F.relu(_bn_eval(d, sd, prefix + ".bn"))
    res = F.conv2d(_cbr2d(torch.cat((d, conv0), 1), sd, prefix + ".conv3"), sd[prefix + ".res.weight"], padding=1)
    depth = (F.interpolate(depth, scale_factor=2, mode="bilinear", align_corners=True) + res) / 10
    return depth * (hi - lo) + lo


def cdsmvsnet_forward(sd, imgs, proj_matrices, depth_values, ndepths, ratios, temperature=0.01,
                      return_intermediates=False, refine=False):
    """CDSMVSNet.forward with grad_method='detach', share_cr=False (model.py:140-223).  refine=True: the cascade works at half
    the image resolution on nearest-subsampled images (model.py:145-147,159-160) and the Refinement network lifts the last
    depth map back to full resolution (model.py:209-216)."""
    full_imgs = imgs
    if refine:
        imgs = imgs[..., ::2, ::2]      # F.interpolate(img, (H/2, W/2)) is nearest: picks pixel (2i, 2j)
    B, N, _, H, W = imgs.shape
    dmin, dmax = depth_values[:, 0], depth_values[:, -1]
    interval = depth_values[:, 1] - depth_values[:, 0]
    cams3 = torch.unbind(proj_matrices["stage3"], 1)
    feats = []
    for i in range(1, N):
        Fm = fundamental_matrix(cams3[0], cams3[i])
        e_ref, e_src = epipole_from_F(Fm), epipole_from_F(Fm.transpose(1, 2))
        feats.append({"ref": feature_net(imgs[:, 0], sd, e_ref, temperature),
                      "src": feature_net(imgs[:, i], sd, e_src, temperature)})
    outputs, depth = {}, None
    for s, D in enumerate(ndepths):
        name = f"stage{s + 1}"
        scale = (4, 2, 1)[s]
        cur = depth_values if depth is None else depth
        samples = depth_hypotheses(cur, D, ratios[s] * interval, H, W, dmin, dmax, scale)
        stage_feats = [{"ref": f["ref"][name], "src": f["src"][name]} for f in feats]
        o = stage_net(stage_feats, proj_matrices[name], samples, sd, s, return_intermediates)
        if return_intermediates:
            o["_inter"]["depth_samples"] = samples
        depth = o["depth"]
        outputs[name] = o
        outputs.update({k: v for k, v in o.items() if k != "_inter"})
    if refine:
        iv = interval.view(B, 1, 1)
        refined = refinement(sd, full_imgs[:, 0], (depth / iv).unsqueeze(1), dmin / interval, dmax / interval)
        outputs["refined_depth"] = refined.squeeze(1) * iv
    else:
        outputs["refined_depth"] = depth
    if return_intermediates:
        outputs["_features"] = feats
    return outputs


# ----------------------------------------------------------------------------------------------
# helpers shared by tests / bench
# ----------------------------------------------------------------------------------------------
# ----------------------------------------------------------------------------------------------
# training loss: models/losses.py:6-48
# ----------------------------------------------------------------------------------------------
def stage_loss(est, gt, mask, interval, curv=None):
    """(smooth-L1 mean of est/iv - gt/iv over mask > 0.5, masked mean of curv) -- losses.py:14-23, fp64 sums."""
    iv = interval.reshape(-1, 1, 1)
    on = (mask > 0.5)
    diff = (est / iv - gt / iv)
    a = diff.abs()
    per = torch.where(a < 1, 0.5 * diff * diff, a - 0.5)
    n = on.double().sum()
    depth_loss = ((per.double() * on).sum() / n).float()
    curv_mean = None if curv is None else ((curv.reshape(est.shape).double() * on).sum() / n).float()
    return depth_loss, curv_mean


def stage_loss_backward(est, gt, mask, interval, g_depth=1.0, g_curv=1.0):
    """(d/d est, d/d curv) of the two means above."""
    iv = interval.reshape(-1, 1, 1)
    on = (mask > 0.5)
    diff = (est / iv - gt / iv)
    n = on.float().sum()
    slope = torch.where(diff.abs() < 1, diff, torch.sign(diff))
    return g_depth * slope / iv / n * on, g_curv * on.float() / n


def feat_loss(feat_dis, target, mask):
    """losses.py:25-35: BCE with logits over mask > 0.5 repeated across the planes, positives weighted by neg / pos."""
    on = (mask > 0.5).unsqueeze(1).expand_as(feat_dis)
    n = on.double().sum()
    pos = (target.double() * on).sum()
    pw = ((n - pos) / pos).float()
    x, y = feat_dis, target
    softplus_neg = torch.clamp(-x, min=0) + torch.log1p(torch.exp(-x.abs()))
    per = (1 - y) * x + (1 + (pw - 1) * y) * softplus_neg
    return ((per.double() * on).sum() / n).float()


def feat_loss_backward(feat_dis, target, mask, g=1.0):
    on = (mask > 0.5).unsqueeze(1).expand_as(feat_dis)
    n = on.float().sum()
    pw = (n - (target * on).sum()) / (target * on).sum()
    return g * ((1 - target) - (1 + (pw - 1) * target) * torch.sigmoid(-feat_dis)) / n * on


def final_loss(inputs, depth_gt_ms, mask_ms, dlossw=None, depth_interval=None):
    total, depth_loss = torch.zeros(()), None
    for i, k in enumerate(("stage1", "stage2", "stage3")):
        depth_loss, curv = stage_loss(inputs[k]["depth"], depth_gt_ms[k], mask_ms[k], depth_interval, inputs[k]["norm_curv"])
        fl = feat_loss(inputs[k]["feat_distance"], inputs[k]["feat_target"], mask_ms[k]) if "feat_distance" in inputs[k] else 0.0
        total = total + (1.0 if dlossw is None else dlossw[i]) * (depth_loss + 5 * fl + 0.1 * curv)
    if "refined_depth" in inputs:
        depth_loss, _ = stage_loss(inputs["refined_depth"], depth_gt_ms["stage4"], mask_ms["stage4"], depth_interval)
        total = total + 2 * depth_loss
    return total, depth_loss


def rel_l1(a: torch.Tensor, b: torch.Tensor) -> float:
    """mean |a-b| / mean |b| -- the parity metric north_star quotes (1e-3 on depth)."""
    return float((a.double() - b.double()).abs().mean() / b.double().abs().mean().clamp_min(1e-30))


def strip_module_prefix(sd: dict) -> dict:
    return {(k[7:] if k.startswith("module.") else k): v for k, v in sd.items()}
