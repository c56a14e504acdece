"""CostRegNet in training mode: forward with BatchNorm3d batch statistics and the full backward, on csrc/train3d.cu.

Reference: models/module.py:80-166 (blocks), :270-315 (CostRegNet).  ``CostRegNet.forward`` (modules.py) routes here when the
module is in training mode; the result carries an autograd node whose backward produces the gradient of the input volume and
of every parameter (conv weights, BatchNorm affine), so the reference's training loop (loss.backward(), optimizer.step())
runs on it unchanged.  Running statistics are updated like ``nn.BatchNorm3d`` does (momentum, unbiased variance).
"""
from __future__ import annotations

import torch

from ._lib import call, ptr

BLOCKS = ("conv0", "conv1", "conv2", "conv3", "conv4", "conv5", "conv6", "conv7", "conv9", "conv11")
STRIDE = {"conv0": 1, "conv1": 2, "conv2": 1, "conv3": 2, "conv4": 1, "conv5": 2, "conv6": 1}      # the others are transposed
SKIP = {"conv7": "conv4", "conv9": "conv2", "conv11": "conv0"}                                        # models/module.py:310-312


def _f(t):
    return t.detach().to(torch.float32).contiguous()


def _tap_conv(w):      # conv weight [Cout,Cin,3,3,3] -> [Cin][27][Cout]
    return w.permute(1, 2, 3, 4, 0).reshape(w.shape[1], 27, w.shape[0]).contiguous()


def _tap_deconv(w):    # transposed-conv weight [Cin,Cout,3,3,3] -> [Cin][27][Cout]
    return w.permute(0, 2, 3, 4, 1).reshape(w.shape[0], 27, w.shape[1]).contiguous()


def conv3d(x, w_tap, cout, stride):
    B, ci, D, H, W = x.shape
    out = torch.empty(B, cout, -(-D // stride), -(-H // stride), -(-W // stride), device=x.device)
    call("cds_train_conv3d", ptr(x), ptr(w_tap), B, ci, cout, D, H, W, stride, ptr(out))
    return out


def deconv3d(x, w_tap, cout):
    B, ci, D, H, W = x.shape
    out = torch.empty(B, cout, 2 * D, 2 * H, 2 * W, device=x.device)
    call("cds_train_deconv3d", ptr(x), ptr(w_tap), B, ci, cout, D, H, W, ptr(out))
    return out


def wgrad(x, g, stride):
    """[Cin][27][Cout] = sum over voxels of g (x) shifted x."""
    B, ci, D, H, W = x.shape
    dw = torch.empty(ci, 27, g.shape[1], device=x.device)
    call("cds_train_conv3d_wgrad", ptr(x), ptr(g), B, ci, g.shape[1], D, H, W, stride, ptr(dw))
    return dw


def bn_forward(raw, gamma, beta, eps, relu, skip):
    B, C = raw.shape[:2]
    V = raw[0, 0].numel()
    sums = torch.empty(C, 2, dtype=torch.float64, device=raw.device)
    call("cds_train_bn_stats", ptr(raw), B, C, V, ptr(sums))
    n = B * V
    mean = sums[:, 0] / n
    var = (sums[:, 1] / n - mean * mean).clamp_min(0.0)                  # biased, as the normalisation uses it
    rstd = torch.rsqrt(var + eps)
    mean32, rstd32 = mean.float().contiguous(), rstd.float().contiguous()
    y = torch.empty_like(raw)
    call("cds_train_bn_apply", ptr(raw), ptr(mean32), ptr(rstd32), ptr(gamma), ptr(beta), ptr(skip), int(relu), B, C, V, ptr(y))
    return y, mean32, rstd32, mean, var * (n / max(n - 1, 1))             # + batch mean / UNBIASED variance for the running stats


def bn_backward(dy, raw, mean, rstd, gamma, beta, relu):
    B, C = raw.shape[:2]
    V = raw[0, 0].numel()
    sums = torch.empty(C, 2, dtype=torch.float64, device=raw.device)
    dx = torch.empty_like(raw)
    call("cds_train_bn_backward", ptr(dy), ptr(raw), ptr(mean), ptr(rstd), ptr(gamma), ptr(beta), int(relu), B, C, V, ptr(sums), ptr(dx))
    return dx, sums[:, 1].float(), sums[:, 0].float()                     # dx, dgamma, dbeta


class CostRegTrainFn(torch.autograd.Function):
    """forward(x, eps tuple, *params) with params = for every block (conv.weight, bn.weight, bn.bias), then prob.weight."""

    @staticmethod
    def forward(ctx, x, eps, *params):
        x = _f(x)
        if x.dim() != 5 or any(s % 8 for s in x.shape[2:]):
            raise RuntimeError(f"CostRegNet (training): D, H, W must be divisible by 8, got {tuple(x.shape)}")
        p = [_f(t) for t in params]
        saved, acts, stats = {}, {}, []
        cur = x
        for i, name in enumerate(BLOCKS):
            w, gamma, beta = p[3 * i:3 * i + 3]
            if name in STRIDE:
                raw = conv3d(cur, _tap_conv(w), w.shape[0], STRIDE[name])
            else:
                raw = deconv3d(cur, _tap_deconv(w), w.shape[1])
            skip = acts[SKIP[name]] if name in SKIP else None
            y, mean, rstd, bmean, bvar = bn_forward(raw, gamma, beta, eps[i], True, skip)
            saved[name] = (cur, raw, mean, rstd)
            stats.append((bmean, bvar))
            acts[name] = cur = y
        wp = p[-1]
        logits = conv3d(cur, _tap_conv(wp), 1, 1)
        ctx.saved = saved
        ctx.last = cur
        ctx.params = p
        flat = tuple(t.float() for bm_bv in stats for t in bm_bv)
        ctx.mark_non_differentiable(*flat)
        return (logits,) + flat

    @staticmethod
    def backward(ctx, g_logits, *_unused):
        p, saved = ctx.params, ctx.saved
        g = _f(g_logits)
        grads = [None] * len(p)
        wp = p[-1]
        x11 = ctx.last
        grads[-1] = wgrad(x11, g, 1).reshape(wp.shape[1], 3, 3, 3, wp.shape[0]).permute(4, 0, 1, 2, 3).contiguous()
        # input gradient of a stride-1 conv: the same conv with flipped, transposed weights
        dcur = conv3d(g, wp.flip(2, 3, 4).permute(0, 2, 3, 4, 1).reshape(wp.shape[0], 27, wp.shape[1]).contiguous(), wp.shape[1], 1)
        pending = {}                                                        # gradients flowing into skip sources
        for i in reversed(range(len(BLOCKS))):
            name = BLOCKS[i]
            w, gamma, beta = p[3 * i:3 * i + 3]
            xin, raw, mean, rstd = saved[name]
            if name in pending:
                dcur = dcur + pending.pop(name)
            if name in SKIP:                                                # y = relu(bn(.)) + skip: the skip source gets dy as is
                src = SKIP[name]
                pending[src] = pending[src] + dcur if src in pending else dcur
            draw, dgamma, dbeta = bn_backward(dcur, raw, mean, rstd, gamma, beta, True)
            if name in STRIDE:
                s = STRIDE[name]
                grads[3 * i] = wgrad(xin, draw, s).reshape(w.shape[1], 3, 3, 3, w.shape[0]).permute(4, 0, 1, 2, 3).contiguous()
                if s == 1:
                    dcur = conv3d(draw, w.flip(2, 3, 4).permute(0, 2, 3, 4, 1).reshape(w.shape[0], 27, w.shape[1]).contiguous(), w.shape[1], 1)
                else:     # stride-2 conv: its weight [Cout,Cin,k] IS a transposed-conv weight [in = Cout, out = Cin, k]
                    dcur = deconv3d(draw, w.permute(0, 2, 3, 4, 1).reshape(w.shape[0], 27, w.shape[1]).contiguous(), w.shape[1])
            else:         # transposed block [Cin_t, Cout_t, k]: swap the roles of input and gradient
                grads[3 * i] = wgrad(draw, xin, 2).reshape(w.shape[1], 3, 3, 3, w.shape[0]).permute(4, 0, 1, 2, 3).contiguous()
                dcur = conv3d(draw, w.permute(1, 2, 3, 4, 0).reshape(w.shape[1], 27, w.shape[0]).contiguous(), w.shape[0], 2)
            grads[3 * i + 1], grads[3 * i + 2] = dgamma, dbeta
        assert not pending
        return (dcur, None) + tuple(grads)


def costreg_train_forward(module, x):
    """Training-mode forward of ``modules.CostRegNet``: logits [B,1,D,H,W] with an autograd node; updates the running statistics."""
    blocks = [getattr(module, n) for n in BLOCKS]
    params = [t for b in blocks for t in (b.conv.weight, b.bn.weight, b.bn.bias)] + [module.prob.weight]
    eps = tuple(float(b.bn.eps) for b in blocks)
    out = CostRegTrainFn.apply(x, eps, *params)
    logits, stats = out[0], out[1:]
    with torch.no_grad():
        for i, b in enumerate(blocks):
            bn = b.bn
            if bn.track_running_stats and bn.running_mean is not None:
                bn.num_batches_tracked += 1
                m = bn.momentum if bn.momentum is not None else 1.0 / float(bn.num_batches_tracked)
                bn.running_mean.mul_(1 - m).add_(stats[2 * i].to(bn.running_mean.dtype), alpha=m)
                bn.running_var.mul_(1 - m).add_(stats[2 * i + 1].to(bn.running_var.dtype), alpha=m)
    return logits
