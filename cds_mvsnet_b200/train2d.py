"""DynamicConv in training mode: forward and backward with the branch convolutions on csrc/train2d.cu.

Reference: models/dynamic_conv.py:97-122.  ``DynamicConv.forward`` (modules.py) routes here when the module is in training
mode.  The FLOPs of the layer -- per kernel size k a (Cout + 3)-channel k x k convolution of the input (feature branch and the
three curvature coefficients together), its input gradient and its weight gradient -- run on this repository's CUDA kernels
through ``Conv2dFn``; the per-pixel remainder (direction field, quadratic form, the gate MLP with its BatchNorm2d in batch-
statistics mode, softmax over the branches, blend) is a handful of torch element-wise ops on [B, K, H, W] maps whose autograd
does the bookkeeping, and calls the module's own ``att_weights`` container so the running statistics move as the reference's do.
"""
from __future__ import annotations

import torch

from ._lib import call, ptr


def _tap(w):           # [Cout,Cin,k,k] -> [Cin][k*k][Cout]
    co, ci, k, _ = w.shape
    return w.permute(1, 2, 3, 0).reshape(ci, k * k, co).contiguous()


def _conv(x, w_tap, cout, k):
    B, ci, H, W = x.shape
    out = torch.empty(B, cout, H, W, device=x.device)
    call("cds_train_conv2d", ptr(x), ptr(w_tap), B, ci, cout, H, W, k, ptr(out))
    return out


class Conv2dFn(torch.autograd.Function):
    """conv2d(x, w) with w [Cout,Cin,k,k], stride 1, pad (k-1)/2, no bias; fp32."""

    @staticmethod
    def forward(ctx, x, w):
        x, w = x.detach().float().contiguous(), w.detach().float().contiguous()
        ctx.save_for_backward(x, w)
        return _conv(x, _tap(w), w.shape[0], w.shape[2])

    @staticmethod
    def backward(ctx, g):
        x, w = ctx.saved_tensors
        g = g.detach().float().contiguous()
        co, ci, k, _ = w.shape
        dx = dw = None
        if ctx.needs_input_grad[0]:      # the same convolution with flipped, transposed weights: [in = Cout][k*k][out = Cin]
            dx = _conv(g, w.flip(2, 3).permute(0, 2, 3, 1).reshape(co, k * k, ci).contiguous(), ci, k)
        if ctx.needs_input_grad[1]:
            B, _, H, W = x.shape
            d = torch.empty(ci, k * k, co, device=x.device)
            call("cds_train_conv2d_wgrad", ptr(x), ptr(g), B, ci, co, H, W, k, ptr(d))
            dw = d.reshape(ci, k, k, co).permute(3, 0, 1, 2).contiguous()
        return dx, dw


def dynconv_train_forward(module, x, epipole, temperature):
    """(blended features [B,Cout,H,W], norm_curv [B,1,H,W]) of ``modules.DynamicConv`` in training mode."""
    B, _, H, W = x.shape
    dev = x.device
    ys = torch.arange(H, dtype=torch.float32, device=dev).view(1, 1, H, 1)
    xs = torch.arange(W, dtype=torch.float32, device=dev).view(1, 1, 1, W)
    e = epipole.to(device=dev, dtype=torch.float32)
    u, v = xs - e[:, 0].view(B, 1, 1, 1), ys - e[:, 1].view(B, 1, 1, 1)          # direction from the epipole to the pixel
    r = torch.sqrt(u * u + v * v) + 1e-6
    u, v = u / r, v / r
    quad = torch.cat((u * u, 2 * u * v, v * v), dim=1)                            # [B,3,H,W]
    co = module.out_c
    feats, curvs = [], []
    for conv, att in zip(module.convs, module.att_convs):
        both = Conv2dFn.apply(x, torch.cat((conv.weight, att.weight), dim=0))    # feature branch + (a, b, c) in one convolution
        y = both[:, :co]
        if conv.bias is not None:
            y = y + conv.bias.view(1, co, 1, 1)
        feats.append(y)
        curvs.append((both[:, co:] * quad).sum(dim=1, keepdim=True))
    curv = torch.cat(curvs, dim=1)                                                # [B,K,H,W]
    gate = torch.softmax(module.att_weights(curv) / temperature, dim=1)
    out = sum(f * gate[:, i:i + 1] for i, f in enumerate(feats))
    return out, (curv * gate).sum(dim=1, keepdim=True)
