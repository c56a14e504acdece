"""Drop-in for the reference's training loss (models/losses.py:6-48), first slice of SURVEY.md 8f-3.

``final_loss(inputs, depth_gt_ms, mask_ms, dlossw=..., depth_interval=...)`` keeps the reference's signature, return pair
``(total_loss, depth_loss)`` and arithmetic: per stage the smooth-L1 (beta 1) mean of ``depth / interval`` against the
ground truth over ``mask > 0.5`` plus 0.1 x the masked mean of ``norm_curv``, weighted by ``dlossw[stage]``; plus
2 x the same depth term for ``refined_depth`` (stage4 maps).  Each stage is ONE reduction kernel forward and one
elementwise kernel backward (``csrc/train.cu``) instead of the reference's boolean-index gathers; the handful of scalar
combinations stay torch ops on the device (no host sync).

The ``feat_distance`` / ``feat_target`` binary-cross-entropy term (losses.py:25-35) is built too (count pass, weighted loss pass,
elementwise backward), so a training-mode output dict of the REFERENCE's StageNet can be fed; this package's own fused
StageNet stays inference-only and never produces those keys.
"""
from __future__ import annotations

import torch

from ._lib import call, ptr


def _f32c(t):
    return t.to(torch.float32).contiguous()


class _StageLossFn(torch.autograd.Function):
    """(masked smooth-L1 mean of est/iv - gt/iv, masked mean of curv) with gradients to ``est`` and ``curv``."""

    @staticmethod
    def forward(ctx, est, curv, gt, mask, interval):
        if not est.is_cuda:
            raise RuntimeError("cds_b200 ops run on a CUDA device (B200) only; got a CPU tensor. There is no CPU fallback.")
        B, H, Wd = est.shape
        e, g, m, iv = _f32c(est), _f32c(gt), _f32c(mask), _f32c(interval).reshape(-1)
        if iv.numel() == 1 and B > 1:
            iv = iv.expand(B).contiguous()
        if tuple(g.shape) != (B, H, Wd) or tuple(m.shape) != (B, H, Wd) or iv.numel() != B:
            raise RuntimeError(f"final_loss: depth {tuple(est.shape)}, gt {tuple(gt.shape)}, mask {tuple(mask.shape)}, "
                               f"interval {tuple(interval.shape)} do not agree")
        c = None if curv is None else _f32c(curv).reshape(B, H, Wd)
        sums = torch.zeros(3, dtype=torch.float64, device=est.device)
        call("cds_stage_loss_forward", ptr(e), ptr(g), ptr(m), ptr(iv), ptr(c), B, H, Wd, ptr(sums))
        ctx.save_for_backward(e, g, m, iv, sums)
        ctx.curv_shape = None if curv is None else tuple(curv.shape)
        ctx.dtypes = (est.dtype, None if curv is None else curv.dtype)
        depth_loss = (sums[0] / sums[1]).float()
        curv_mean = (sums[2] / sums[1]).float()
        return depth_loss, curv_mean

    @staticmethod
    def backward(ctx, g_depth, g_curv):
        e, g, m, iv, sums = ctx.saved_tensors
        B, H, Wd = e.shape
        need_e, need_c = ctx.needs_input_grad[0], ctx.needs_input_grad[1] and ctx.curv_shape is not None
        gd = _f32c(g_depth) if need_e else None
        gc = _f32c(g_curv) if need_c else None
        grad_e = torch.empty_like(e) if need_e else None
        grad_c = torch.empty_like(e) if need_c else None
        if need_e or need_c:
            call("cds_stage_loss_backward", ptr(e), ptr(g), ptr(m), ptr(iv), ptr(sums), ptr(gd), ptr(gc), B, H, Wd, ptr(grad_e),
                 ptr(grad_c))
        return (grad_e.to(ctx.dtypes[0]) if need_e else None,
                grad_c.reshape(ctx.curv_shape).to(ctx.dtypes[1]) if need_c else None, None, None, None)


class _FeatLossFn(torch.autograd.Function):
    """binary_cross_entropy_with_logits(feat_dis[mask], target[mask], mean, pos_weight = neg / pos), mask repeated over D."""

    @staticmethod
    def forward(ctx, feat_dis, target, mask):
        if not feat_dis.is_cuda:
            raise RuntimeError("cds_b200 ops run on a CUDA device (B200) only; got a CPU tensor. There is no CPU fallback.")
        B, D, H, Wd = feat_dis.shape
        x, y, m = _f32c(feat_dis), _f32c(target), _f32c(mask)
        if tuple(y.shape) != (B, D, H, Wd) or tuple(m.shape) != (B, H, Wd):
            raise RuntimeError(f"final_loss: feat_distance {tuple(feat_dis.shape)}, feat_target {tuple(target.shape)}, "
                               f"mask {tuple(mask.shape)} do not agree")
        sums = torch.zeros(3, dtype=torch.float64, device=x.device)
        call("cds_feat_loss_forward", ptr(x), ptr(y), ptr(m), B, D, H, Wd, ptr(sums))
        ctx.save_for_backward(x, y, m, sums)
        ctx.dtype = feat_dis.dtype
        return (sums[2] / sums[1]).float()

    @staticmethod
    def backward(ctx, g):
        x, y, m, sums = ctx.saved_tensors
        B, D, H, Wd = x.shape
        if not ctx.needs_input_grad[0]:
            return None, None, None
        gg = _f32c(g)
        grad = torch.empty_like(x)
        call("cds_feat_loss_backward", ptr(x), ptr(y), ptr(m), ptr(sums), ptr(gg), B, D, H, Wd, ptr(grad))
        return grad.to(ctx.dtype), None, None


def final_loss(inputs, depth_gt_ms, mask_ms, **kwargs):
    """models/losses.py:6-48 -> (total_loss, depth_loss of the last term)."""
    depth_loss_weights = kwargs.get("dlossw", None)
    depth_interval = kwargs.get("depth_interval", 1.0)
    dev = mask_ms["stage1"].device
    if not torch.is_tensor(depth_interval):
        depth_interval = torch.tensor([float(depth_interval)], dtype=torch.float32, device=dev)
    total_loss = torch.zeros((), dtype=torch.float32, device=dev)
    depth_loss = 0.0
    for stage_key in ("stage1", "stage2", "stage3"):
        stage_inputs = inputs[stage_key]
        depth_loss, curv_reg = _StageLossFn.apply(stage_inputs["depth"], stage_inputs["norm_curv"], depth_gt_ms[stage_key],
                                                  mask_ms[stage_key], depth_interval)
        feat_loss = 0.0
        if "feat_distance" in stage_inputs:
            feat_loss = _FeatLossFn.apply(stage_inputs["feat_distance"], stage_inputs["feat_target"], mask_ms[stage_key])
        w = 1.0 if depth_loss_weights is None else depth_loss_weights[int(stage_key.replace("stage", "")) - 1]
        total_loss = total_loss + w * (depth_loss + 5 * feat_loss + 0.1 * curv_reg)
    if "refined_depth" in inputs:
        depth_loss, _ = _StageLossFn.apply(inputs["refined_depth"], None, depth_gt_ms["stage4"], mask_ms["stage4"], depth_interval)
        total_loss = total_loss + 2 * depth_loss
    return total_loss, depth_loss


def temperature_for_epoch(epoch: int) -> float:
    """The DynamicConv softmax temperature of training epoch ``epoch`` (1-based), trainer/trainer.py:45-49: one decade every
    two epochs from 1.0 (epoch 1) down to 10^-1.5 (epoch 4), then 0.01."""
    return float(10.0 ** (-(epoch - 1) / 2.0)) if epoch <= 4 else 0.01
