"""Drop-in replacements for the reference's hot-path call surface.

Same names, constructor arguments, forward signatures, output keys, error behaviour and
state-dict keys/shapes as the reference, so they can be rebound inside an imported reference
``models.model`` / ``models.module`` (see ``patch``) or used on their own:

    homo_warping_3D(src_fea, src_proj, ref_proj, depth_values)          models/utils/warping.py:69
    depth_regression(p, depth_values) / conf_regression(p, n=4)          models/module.py:373,382
    DynamicConv(in_c, out_c, size_kernels, stride, bias, thresh_scale)   models/dynamic_conv.py:81
    FeatureNet(base_channels, num_stage, stride, arch_mode)              models/module.py:201
    CostRegNet(in_channels, base_channels, last_layer, full_res)         models/module.py:270
    StageNet(num_mvs_stages)                                             models/model.py:11
    CDSMVSNet(refine, ndepths, depth_interals_ratio, share_cr, ...)      models/model.py:97

The torch ``nn`` sub-modules inside these classes are parameter CONTAINERS only (they give the
state dict the reference's key names); they are never called.  Every forward runs the CUDA kernels
of libcds_b200.so through the C ABI and raises if the library or a B200 is missing.  Public tensors
are fp32 NCHW / NCDHW like the reference's; the fused paths keep channels-last fp16 internally.
``CDSMVSNet`` covers ``refine=False`` and ``refine=True`` (``Refinement``, models/module.py:318-370).  The op-level
``homo_warping_3D`` / ``depth_regression`` are differentiable (their backward kernels live in csrc/train.cu); ``CostRegNet``
and ``DynamicConv`` switch to training forms with ``.train()`` (batch-statistics BatchNorm, full backward: train3d.py /
train2d.py), which is what the reference's training loop needs at ``patch(level="leaf")``; the fused ``FeatureNet`` /
``StageNet`` / ``CDSMVSNet`` forwards are inference only (eval mode): in training mode they raise NotImplementedError.

``patch(models_model, models_module, level=...)`` rebinds these names inside an imported reference tree at four depths:
"leaf" swaps only the leaf operators (DynamicConv, CostRegNet, the warp and regression functions), "ops" also FeatureNet and keeps the reference's own ``CDSMVSNet.forward`` and ``StageNet.forward`` and swaps the operators they call,
"stage" also swaps ``StageNet`` (the fused cost-volume kernels), "model" swaps ``CDSMVSNet`` itself (one CUDA graph).
"""
from __future__ import annotations

import ctypes

import torch
import torch.nn as nn

from . import _lib, weights as W
from ._lib import ACT_NONE, ACT_TANH, call, ptr
from .engine import STAGE_CHANNELS, Buffers, CascadeEngine, FeatureExtractor, Regulariser

DEFAULT_STORAGE = torch.float16


def _dev(t: torch.Tensor) -> torch.device:
    if not t.is_cuda:
        raise RuntimeError("cds_b200 ops run on a CUDA device (B200) only; got a CPU tensor. There is no CPU fallback.")
    return t.device


def _f32c(t):
    return t.to(torch.float32).contiguous()


# ------------------------------------------------------------------------------------------------
# functions
# ------------------------------------------------------------------------------------------------
def _warp_forward(src_fea, src_proj, ref_proj, depth_values):
    dev = _dev(src_fea)
    B, C, H, Wd = src_fea.shape
    D = depth_values.shape[1]
    per_pixel = depth_values.dim() == 4
    if per_pixel and tuple(depth_values.shape) != (B, D, H, Wd):
        raise RuntimeError(f"depth_values {tuple(depth_values.shape)} does not match features {tuple(src_fea.shape)}")
    coef = torch.empty(B, 12, dtype=torch.float32, device=dev)
    sp, rp = _f32c(src_proj), _f32c(ref_proj)   # named: temporaries must outlive the launch's pointer use
    call("cds_warp_coeffs", ptr(sp), ptr(rp), B, ptr(coef))
    out = torch.empty(B, C, D, H, Wd, dtype=torch.float32, device=dev)
    src, dep = _f32c(src_fea), _f32c(depth_values)
    call("cds_homo_warp", ptr(src), ptr(coef), ptr(dep), int(per_pixel), B, C, D, H, Wd, ptr(out))
    return out, coef, dep


class _HomoWarpFn(torch.autograd.Function):
    """homo_warping_3D with its backward: the gradient reaches src_fea only, the grid is no_grad (warping.py:79)."""

    @staticmethod
    def forward(ctx, src_fea, src_proj, ref_proj, depth_values):
        out, coef, dep = _warp_forward(src_fea, src_proj, ref_proj, depth_values)
        ctx.save_for_backward(coef, dep)
        ctx.src_shape, ctx.src_dtype = tuple(src_fea.shape), src_fea.dtype
        return out

    @staticmethod
    def backward(ctx, grad_out):
        coef, dep = ctx.saved_tensors
        B, C, H, Wd = ctx.src_shape
        D = dep.shape[1]
        g = _f32c(grad_out)
        if C % 4 == 0:   # vector reductions into a channels-last workspace, then one layout pass (two kernels behind one call)
            grad_src = torch.empty(B, C, H, Wd, dtype=torch.float32, device=g.device)
            ws = torch.zeros(B, H, Wd, C, dtype=torch.float32, device=g.device)
        else:
            grad_src, ws = torch.zeros(B, C, H, Wd, dtype=torch.float32, device=g.device), None
        call("cds_homo_warp_backward", ptr(g), ptr(coef), ptr(dep), int(dep.dim() == 4), B, C, D, H, Wd, ptr(grad_src), ptr(ws))
        return grad_src.to(ctx.src_dtype), None, None, None


def homo_warping_3D(src_fea, src_proj, ref_proj, depth_values):
    """[B,C,H,W], [B,4,4], [B,4,4], [B,D] | [B,D,H,W] -> [B,C,D,H,W] (models/utils/warping.py:69-104).

    Differentiable in ``src_fea`` (the only gradient the reference propagates, warping.py:79)."""
    if torch.is_grad_enabled() and src_fea.requires_grad:
        return _HomoWarpFn.apply(src_fea, src_proj, ref_proj, depth_values)
    return _warp_forward(src_fea, src_proj, ref_proj, depth_values)[0]


def _regress_forward(p, depth_values):
    dev = _dev(p)
    B, D, H, Wd = p.shape
    out = torch.empty(B, H, Wd, dtype=torch.float32, device=dev)
    pc, dv = _f32c(p), _f32c(depth_values)
    if dv.dim() == 1:
        dv = dv.unsqueeze(0).expand(B, D).contiguous()
    call("cds_softmax_regress", ptr(pc), ptr(dv), int(dv.dim() == 4), 1, B, D, H, Wd, ptr(out), None, None)
    return out, pc, dv


class _DepthRegressFn(torch.autograd.Function):
    """depth_regression with its backward (d/dp = g * depth_d, d/d depth = g * p_d)."""

    @staticmethod
    def forward(ctx, p, depth_values):
        out, pc, dv = _regress_forward(p, depth_values)
        ctx.save_for_backward(pc, dv)
        ctx.dv_shape, ctx.dtypes = tuple(depth_values.shape), (p.dtype, depth_values.dtype)
        return out

    @staticmethod
    def backward(ctx, grad_depth):
        pc, dv = ctx.saved_tensors
        B, D, H, Wd = pc.shape
        g = _f32c(grad_depth)
        need_p, need_dv = ctx.needs_input_grad
        grad_p = torch.empty_like(pc) if need_p else None
        grad_dv = torch.empty_like(pc) if need_dv else None
        call("cds_depth_regress_backward", ptr(g), ptr(pc), ptr(dv), int(dv.dim() == 4), B, D, H, Wd, ptr(grad_p), ptr(grad_dv))
        if need_dv and len(ctx.dv_shape) != 4:   # plane depths shared by all pixels: fold the per-pixel products
            grad_dv = grad_dv.sum((2, 3))
            grad_dv = grad_dv.sum(0) if len(ctx.dv_shape) == 1 else grad_dv
        return (grad_p.to(ctx.dtypes[0]) if need_p else None), (grad_dv.to(ctx.dtypes[1]) if need_dv else None)


def depth_regression(p, depth_values):
    """sum_d p_d * depth_d (models/module.py:373-379); depth_values [B,D] or [B,D,H,W].  Differentiable."""
    if torch.is_grad_enabled() and (p.requires_grad or depth_values.requires_grad):
        return _DepthRegressFn.apply(p, depth_values)
    return _regress_forward(p, depth_values)[0]


def conf_regression(p, n=4):
    """p[d-1]+p[d]+p[d+1]+p[d+2] at d = clamp(trunc(sum_k k p_k)) (models/module.py:382-391)."""
    if n != 4:
        raise NotImplementedError("conf_regression: only the reference's n=4 window is implemented")
    dev = _dev(p)
    B, D, H, Wd = p.shape
    out = torch.empty(B, H, Wd, dtype=torch.float32, device=dev)
    pc = _f32c(p)
    call("cds_softmax_regress", ptr(pc), None, 0, 1, B, D, H, Wd, None, ptr(out), None)
    return out


# ------------------------------------------------------------------------------------------------
# weight-cache plumbing shared by the modules
# ------------------------------------------------------------------------------------------------
class _CachedModule(nn.Module):
    """Rebuilds the derived (folded / re-laid-out) weights when parameters may have changed.

    ``load_state_dict`` / ``.to()`` / ``.train()`` of THIS module drop the cache at once; everything else that can change a
    weight -- ``model.feature.load_state_dict(...)`` on a submodule, ``p.data.copy_()``, an optimizer step, an EMA swap -- is
    caught by a fingerprint checked at every forward: the (storage address, in-place version counter) of every parameter and
    buffer below this module.  The one thing the fingerprint cannot see is a write through ``p.data`` (it bypasses autograd's
    version counter): call ``invalidate_cache()`` after such an edit."""

    def invalidate_cache(self):
        """Drop the derived weights (folded BatchNorm, operand images, CUDA graph); they are rebuilt by the next forward."""
        self._invalidate()

    def _invalidate(self):
        object.__setattr__(self, "_cache", None)
        object.__setattr__(self, "_fp", None)
        object.__setattr__(self, "_fp_tensors", None)

    def _fingerprint(self):
        ts = self._fp_tensors
        if ts is None:
            ts = list(self.parameters()) + list(self.buffers())
            object.__setattr__(self, "_fp_tensors", ts)
        return tuple([(t.data_ptr(), t._version) for t in ts])

    def _check_cache(self):
        """Drop the derived weights if any parameter / buffer was re-allocated or written since they were built."""
        if self._cache is None:
            object.__setattr__(self, "_fp_tensors", None)   # the module tree may have changed since the last build
            object.__setattr__(self, "_fp", self._fingerprint())
            return
        fp = self._fingerprint()
        if fp != self._fp:
            object.__setattr__(self, "_cache", None)
            object.__setattr__(self, "_fp_tensors", None)
            object.__setattr__(self, "_fp", self._fingerprint())

    def __init__(self):
        super().__init__()
        self._invalidate()
        self._register_load_state_dict_pre_hook(lambda *a, **k: self._invalidate())

    def _apply(self, fn, *a, **k):
        self._invalidate()
        return super()._apply(fn, *a, **k)

    def train(self, mode: bool = True):
        self._invalidate()
        return super().train(mode)

    def _require_eval(self):
        self._check_cache()
        if self.training:
            raise NotImplementedError(f"{type(self).__name__}: this fused CUDA path is inference-only; call .eval() first.  "
                                      "To train, rebind the leaf operators inside the reference's own model classes: "
                                      "patch(models.model, models.module, level=\"leaf\") (INTEGRATION.md section 7)")

    def _sd(self, prefix=""):
        return {prefix + k: v for k, v in self.state_dict().items()}


def _nhwc(x, storage):
    """fp32 NCHW -> channels-last storage tensor via the library's converter."""
    B, C, H, Wd = x.shape
    out = torch.empty(B, H, Wd, C, dtype=storage, device=x.device)
    xc = _f32c(x)
    call("cds_nchw_to_nhwc", ptr(xc), B, C, H, Wd, _lib.dtype_code(storage), ptr(out))
    return out


def _nchw(x):
    B, H, Wd, C = x.shape
    out = torch.empty(B, C, H, Wd, dtype=torch.float32, device=x.device)
    call("cds_nhwc_to_nchw", ptr(x), B, C, H, Wd, _lib.dtype_code(x.dtype), ptr(out))
    return out


# ------------------------------------------------------------------------------------------------
# DynamicConv
# ------------------------------------------------------------------------------------------------
class DynamicConv(_CachedModule):
    def __init__(self, in_c, out_c, size_kernels=(3, 5, 7), stride=1, bias=True, thresh_scale=0.01, **kwargs):
        super().__init__()
        if stride != 1:
            raise NotImplementedError("DynamicConv: stride must be 1 (the reference's att_convs ignore stride, "
                                      "models/dynamic_conv.py:85-86, so any other value breaks there too)")
        self.size_kernels = tuple(size_kernels)
        if len(self.size_kernels) not in (2, 3) or any(k % 2 == 0 or k < 1 for k in self.size_kernels):
            raise NotImplementedError("DynamicConv: the CUDA kernels take 2 or 3 odd kernel sizes (every layer of the reference's "
                                      f"FeatureNet, models/module.py:211-234); got {self.size_kernels}")
        self.thresh_scale = thresh_scale
        self.in_c, self.out_c = in_c, out_c
        self.storage = kwargs.pop("storage", DEFAULT_STORAGE)
        self.use_tc = kwargs.pop("use_tc", True)
        self.att_convs = nn.ModuleList([nn.Conv2d(in_c, 3, k, padding=(k - 1) // 2, bias=False) for k in size_kernels])
        self.convs = nn.ModuleList([nn.Conv2d(in_c, out_c, k, padding=(k - 1) // 2, stride=stride, bias=bias)
                                    for k in size_kernels])
        hidden = kwargs.get("hidden_dim", 4)
        if hidden != 4:
            raise NotImplementedError("DynamicConv: hidden_dim must be 4")
        self.att_weights = nn.Sequential(nn.Conv2d(len(size_kernels), hidden, 1, bias=False), nn.BatchNorm2d(hidden),
                                         nn.ReLU(inplace=True), nn.Conv2d(hidden, len(size_kernels), 1, bias=False))
        for p in self.att_convs.parameters():
            torch.nn.init.normal_(p, std=0.1)

    def forward(self, feature_vol, epipole=None, temperature=0.001):
        """Training mode (``.train()``): the branch convolutions, their input and weight gradients run on csrc/train2d.cu
        (fp32) behind an autograd node; the gate's BatchNorm2d uses batch statistics (train2d.py)."""
        if epipole is None:
            raise TypeError("DynamicConv.forward needs the epipole (the reference dereferences it unconditionally)")
        if self.training:
            from .train2d import dynconv_train_forward
            _dev(feature_vol)
            with torch.cuda.device(feature_vol.device):
                return dynconv_train_forward(self, feature_vol, epipole, temperature)
        self._require_eval()
        dev = _dev(feature_vol)
        B, C, H, Wd = feature_vol.shape
        if self._cache is None:
            object.__setattr__(self, "_cache", W.pack_dynamic_conv(self._sd("x."), "x", self.in_c, self.out_c,
                                                                   self.size_kernels, dev))
        w = self._cache
        dt = _lib.dtype_code(self.storage)
        raw = torch.empty(B, H, Wd, self.out_c, dtype=self.storage, device=dev)
        nc = torch.empty(B, 1, H, Wd, dtype=torch.float32, device=dev)
        ks = (ctypes.c_int * len(w.ksizes))(*w.ksizes)
        epi = _f32c(epipole)
        cin_tc = max(8, self.in_c)
        # tensor-core path: 8-channel operand slabs -- the 3-channel image (padded by cds_image_to_nhwc8) or C % 8 == 0
        if (self.use_tc and self.storage == torch.float16 and (self.in_c == 3 or self.in_c % 8 == 0)
                and _lib.LIB.load().cds_dynamic_conv_tc_supported(cin_tc, self.out_c, H, Wd, len(w.ksizes), ks)):
            # tensor-core path (tcgen05): fp16 channels-last pixels, image zero-padded from 3 to 8 channels
            if w.tc is None:
                w.tc = W.pack_dynamic_conv_tc(w)
            if C == 3:
                x8 = torch.empty(B, H, Wd, 8, dtype=torch.float16, device=dev)
                xc = _f32c(feature_vol)
                call("cds_image_to_nhwc8", ptr(xc), B, H, Wd, ptr(x8))
            else:
                x8 = _nhwc(feature_vol, torch.float16)
            call("cds_dynamic_conv_tc", ptr(x8), B, None, None, ACT_NONE, ptr(epi), 1.0, ptr(w.tc), ptr(w.bias), ptr(w.gate),
                 B, cin_tc, self.out_c, H, Wd, len(w.ksizes), ks, float(temperature), 0, ptr(raw), None, None, ptr(nc), None, 0,
                 None)
            return _nchw(raw), nc
        if C == 3:
            x, mode = _f32c(feature_vol), 1
        else:
            x, mode = _nhwc(feature_vol, self.storage), 0
        call("cds_dynamic_conv", ptr(x), mode, None, None, ACT_NONE, ptr(epi), 1.0, ptr(w.w_att), ptr(w.w_conv),
             ptr(w.bias), ptr(w.gate), B, self.in_c, self.out_c, H, Wd, len(w.ksizes), ks, float(temperature), dt,
             ptr(raw), None, ptr(nc), None, 0, None)
        return _nchw(raw), nc


# ------------------------------------------------------------------------------------------------
# FeatureNet
# ------------------------------------------------------------------------------------------------
class _Conv2dHolder(nn.Module):
    """Parameter container with the reference Conv2d wrapper's key layout (``.conv.*``)."""

    def __init__(self, conv):
        super().__init__()
        self.conv = conv


class FeatureNet(_CachedModule):
    def __init__(self, base_channels, num_stage=3, stride=4, arch_mode="unet", storage=DEFAULT_STORAGE):
        super().__init__()
        assert arch_mode in ["unet", "fpn"], "mode must be in 'unet' or 'fpn', but get:{}".format(arch_mode)
        if base_channels != 8:
            raise NotImplementedError("FeatureNet: the CUDA kernels are instantiated for base_channels=8 (the only "
                                      "value the reference uses, models/model.py:127)")
        self.arch_mode, self.stride, self.base_channels, self.num_stage = arch_mode, stride, base_channels, num_stage
        self.storage = storage
        b = base_channels
        dyn = lambda ci, co, ks, bias: DynamicConv(ci, co, size_kernels=ks, bias=bias)
        self.conv00 = _Conv2dHolder(dyn(3, b, (3, 7, 11), False))
        self.conv01 = _Conv2dHolder(dyn(b, b, (3, 5, 7), False))
        self.downsample1 = _Conv2dHolder(nn.Conv2d(b, 2 * b, 3, stride=2, padding=1, bias=False))
        self.conv10 = _Conv2dHolder(dyn(2 * b, 2 * b, (3, 5), False))
        self.conv11 = _Conv2dHolder(dyn(2 * b, 2 * b, (3, 5), False))
        self.downsample2 = _Conv2dHolder(nn.Conv2d(2 * b, 4 * b, 3, stride=2, padding=1, bias=False))
        self.conv20 = _Conv2dHolder(dyn(4 * b, 4 * b, (1, 3), False))
        self.conv21 = _Conv2dHolder(dyn(4 * b, 4 * b, (1, 3), False))
        self.out1 = dyn(4 * b, 4 * b, (1, 3), True)
        self.inner1 = _Conv2dHolder(nn.Conv2d(6 * b, 2 * b, 1, bias=False))
        self.inner2 = _Conv2dHolder(nn.Conv2d(3 * b, b, 1, bias=False))
        self.out2 = dyn(2 * b, 2 * b, (1, 3), True)
        self.out3 = dyn(b, b, (1, 3), True)
        self.out_channels = [4 * b, 2 * b, b]

    def forward(self, x, epipole=None, temperature=0.001):
        self._require_eval()
        dev = _dev(x)
        if epipole is None:
            raise TypeError("FeatureNet.forward needs the epipole")
        B, C, H, Wd = x.shape
        if H % 4 or Wd % 4:
            raise RuntimeError(f"FeatureNet: H, W must be divisible by 4 (got {H}x{Wd})")
        if self._cache is None:
            object.__setattr__(self, "_cache", (FeatureExtractor(W.pack_feature(self._sd("feature."), dev), self.storage),
                                                Buffers(dev)))
        fx, buf = self._cache
        idx = torch.arange(B, dtype=torch.int32, device=dev)
        feats = fx.run(buf, _f32c(x), idx, _f32c(epipole), B, H, Wd, temperature)
        out = {}
        for s in range(3):
            fea, ncsq, ncabs = feats[s][:3]
            out[f"stage{s + 1}"] = (_nchw(fea), ncsq.unsqueeze(1).clone(), ncabs.unsqueeze(1).clone())
        return out


# ------------------------------------------------------------------------------------------------
# CostRegNet
# ------------------------------------------------------------------------------------------------
class _ConvBn3d(nn.Module):
    def __init__(self, conv, ch):
        super().__init__()
        self.conv = conv
        self.bn = nn.BatchNorm3d(ch, momentum=0.1)


class CostRegNet(_CachedModule):
    def __init__(self, in_channels, base_channels, last_layer=True, full_res=False, storage=DEFAULT_STORAGE):
        super().__init__()
        if full_res:
            raise NotImplementedError("CostRegNet(full_res=True) is never constructed by the reference model "
                                      "(models/model.py:130-134) and is not implemented")
        if not last_layer:
            raise NotImplementedError("CostRegNet(last_layer=False) is not implemented")
        if base_channels != 8 or in_channels not in (8, 16, 32):
            raise NotImplementedError("CostRegNet: kernels are instantiated for base_channels=8, in_channels in {8,16,32}")
        self.last_layer, self.in_channels, self.storage = last_layer, in_channels, storage
        b = base_channels
        c3 = lambda ci, co, s: _ConvBn3d(nn.Conv3d(ci, co, 3, stride=s, padding=1, bias=False), co)
        d3 = lambda ci, co: _ConvBn3d(nn.ConvTranspose3d(ci, co, 3, stride=2, padding=1, output_padding=1, bias=False), co)
        self.conv0 = c3(in_channels, b, 1)
        self.conv1, self.conv2 = c3(b, 2 * b, 2), c3(2 * b, 2 * b, 1)
        self.conv3, self.conv4 = c3(2 * b, 4 * b, 2), c3(4 * b, 4 * b, 1)
        self.conv5, self.conv6 = c3(4 * b, 8 * b, 2), c3(8 * b, 8 * b, 1)
        self.conv7, self.conv9, self.conv11 = d3(8 * b, 4 * b), d3(4 * b, 2 * b), d3(2 * b, b)
        self.prob = nn.Conv3d(b, 1, 3, stride=1, padding=1, bias=False)

    def _engine(self, dev):
        self._check_cache()
        if self._cache is None:
            object.__setattr__(self, "_cache", (Regulariser(W.pack_costreg(self._sd("cr."), "cr", dev), self.storage), Buffers(dev)))
        return self._cache

    def forward(self, x):
        """[B,C,D,H,W] fp32 -> [B,1,D,H,W] fp32 logits (models/module.py:305-315).

        Training mode (``.train()``): BatchNorm3d with batch statistics (running statistics updated) and an autograd node whose
        backward yields the gradients of the volume and of every parameter (train3d.py / csrc/train3d.cu, fp32)."""
        if self.training:
            from .train3d import costreg_train_forward
            _dev(x)
            with torch.cuda.device(x.device):
                return costreg_train_forward(self, x)
        self._require_eval()
        dev = _dev(x)
        B, C, D, H, Wd = x.shape
        reg, buf = self._engine(dev)
        # NCDHW -> channel-blocked [B, C/8, D, H, W, 8]: every (b, c8) group is an 8-channel "image" of D*H rows
        vol = torch.empty(B, C // 8, D, H, Wd, 8, dtype=self.storage, device=dev)
        xc = _f32c(x)
        call("cds_nchw_to_nhwc", ptr(xc), B * (C // 8), 8, D * H, Wd, _lib.dtype_code(self.storage), ptr(vol))
        logits = reg.run(buf, "cr", vol, B, D, H, Wd)
        return logits.unsqueeze(1).clone()


# ------------------------------------------------------------------------------------------------
# StageNet
# ------------------------------------------------------------------------------------------------
class _ConvBnReLU2d(nn.Module):
    def __init__(self, ci, co):
        super().__init__()
        self.conv = nn.Conv2d(ci, co, 3, stride=1, padding=1, bias=False)
        self.bn = nn.BatchNorm2d(co)


class Refinement(_CachedModule):
    """Drop-in for models/module.py:318-370 (same parameter names: conv0..conv3.{conv,bn}, deconv, bn, res)."""

    def __init__(self):
        super().__init__()
        self.conv0, self.conv1, self.conv2 = _ConvBnReLU2d(3, 8), _ConvBnReLU2d(1, 8), _ConvBnReLU2d(8, 8)
        self.deconv = nn.ConvTranspose2d(8, 8, kernel_size=3, padding=1, output_padding=1, stride=2, bias=False)
        self.bn = nn.BatchNorm2d(8)
        self.conv3 = _ConvBnReLU2d(16, 8)
        self.res = nn.Conv2d(8, 1, kernel_size=3, padding=1, bias=False)

    def forward(self, img, depth_0, depth_min, depth_max):
        """img [B,3,H,W], depth_0 [B,1,H/2,W/2], depth_min / depth_max [B] -> refined depth [B,1,H,W]."""
        self._require_eval()
        dev = _dev(img)
        from .engine import Refiner
        if self._cache is None:
            object.__setattr__(self, "_cache", (Refiner(W.pack_refinement(self._sd("r."), "r", dev)), Buffers(dev)))
        refiner, buf = self._cache
        B, _, H, Wd = img.shape
        if tuple(depth_0.shape) != (B, 1, H // 2, Wd // 2):
            raise RuntimeError(f"Refinement: depth_0 must be [B,1,H/2,W/2] = {(B, 1, H // 2, Wd // 2)}, got {tuple(depth_0.shape)}")
        out = refiner.run(buf, _f32c(img), _f32c(depth_0).reshape(B, H // 2, Wd // 2), _f32c(depth_min).reshape(B), _f32c(depth_max).reshape(B),
                          None, B, H, Wd)
        return out.unsqueeze(1).clone()


class StageNet(_CachedModule):
    def __init__(self, num_mvs_stages=3, storage=DEFAULT_STORAGE):
        super().__init__()
        self.storage = storage
        self.vis = nn.ModuleList([nn.Sequential(_ConvBnReLU2d(2, 16), _ConvBnReLU2d(16, 16), _ConvBnReLU2d(16, 16),
                                                nn.Conv2d(16, 1, 1), nn.Sigmoid()) for _ in range(num_mvs_stages)])

    def forward(self, features, proj_matrices, depth_values, num_depth, cost_regularization, prob_volume_init=None,
                stage_idx=0, gt_depth=None):
        """Reference signature (models/model.py:16-17).  features: list of N-1 dicts
        {"ref": (fea[B,C,h,w], nc_sum[B,1,h,w], nc_abs[B,1,h,w]), "src": (...)}; proj_matrices [B,N,2,4,4];
        depth_values [B,D,h,w].  Returns {"depth", "photometric_confidence", "norm_curv"}."""
        self._require_eval()
        if prob_volume_init is not None or gt_depth is not None:
            raise NotImplementedError("StageNet: prob_volume_init / gt_depth belong to the training path (out of scope)")
        assert len(features) == proj_matrices.shape[1] - 1, "Different number of images and projection matrices"
        assert depth_values.shape[1] == num_depth, "depth_values.shape[1]:{}  num_depth:{}".format(depth_values.shape[1], num_depth)
        if not isinstance(cost_regularization, CostRegNet):
            raise TypeError("StageNet: cost_regularization must be the cds_mvsnet_b200 CostRegNet")
        ref0 = features[0]["ref"][0]
        dev = _dev(ref0)
        B, C, h, w = ref0.shape
        V, D = len(features), num_depth
        dt = _lib.dtype_code(self.storage)
        if self._cache is None:
            sd = self._sd("stage_net.")
            object.__setattr__(self, "_cache", ([W.pack_visnet(sd, f"stage_net.vis.{s}", dev) for s in range(len(self.vis))],
                                                Buffers(dev)))
        vis_w, buf = self._cache
        f32 = torch.float32
        ref_fea = torch.stack([_nhwc(f["ref"][0], self.storage) for f in features])     # [V,B,h,w,C]
        src_fea = torch.stack([_nhwc(f["src"][0], self.storage) for f in features])
        ref_ncsq = torch.stack([_f32c(f["ref"][1]).reshape(B, h, w) for f in features])
        src_ncsq = torch.stack([_f32c(f["src"][1]).reshape(B, h, w) for f in features])
        ref_ncabs = torch.stack([_f32c(f["ref"][2]).reshape(B, h, w) for f in features])
        pm = _f32c(proj_matrices)
        coef = torch.empty(1, B, V, 12, dtype=f32, device=dev)
        arr = (ctypes.c_void_p * 1)(pm.data_ptr())
        call("cds_camera_setup", arr, 1, 0, B, V + 1, ptr(coef), None)
        samples = _f32c(depth_values)
        entropy = torch.empty(V, B, h, w, dtype=f32, device=dev)
        call("cds_costvol_entropy", ptr(ref_fea), ptr(src_fea), ptr(coef), ptr(samples), V, B, C, D, h, w, dt, ptr(entropy))
        vis = torch.empty(V, B, h, w, dtype=f32, device=dev)
        call("cds_visnet", ptr(entropy), ptr(ref_ncabs), ptr(vis_w[stage_idx]), V * B, h, w, ptr(vis))
        volume = torch.empty(B, C // 8, D, h, w, 8, dtype=self.storage, device=dev)
        call("cds_costvol_aggregate", ptr(ref_fea), ptr(src_fea), ptr(coef), ptr(samples), ptr(vis), V, B, C, D, h, w, dt, ptr(volume))
        nc = torch.empty(B, 1, h, w, dtype=f32, device=dev)
        call("cds_nc_mean", ptr(ref_ncsq), ptr(src_ncsq), V, B * h * w, ptr(nc))
        reg, rbuf = cost_regularization._engine(dev)
        logits = reg.run(rbuf, "cr", volume, B, D, h, w)
        depth = torch.empty(B, h, w, dtype=f32, device=dev)
        conf = torch.empty(B, h, w, dtype=f32, device=dev)
        call("cds_softmax_regress", ptr(logits), ptr(samples), 1, 0, B, D, h, w, ptr(depth), ptr(conf), None)
        return {"depth": depth, "photometric_confidence": conf, "norm_curv": nc}


# ------------------------------------------------------------------------------------------------
# CDSMVSNet
# ------------------------------------------------------------------------------------------------
class CDSMVSNet(_CachedModule):
    def __init__(self, refine=False, ndepths=(48, 32, 8), depth_interals_ratio=(4, 2, 1), share_cr=False,
                 grad_method="detach", arch_mode="fpn", cr_base_chs=(8, 8, 8), storage=DEFAULT_STORAGE):
        super().__init__()
        assert len(ndepths) == len(depth_interals_ratio)
        if len(ndepths) > 3:
            raise NotImplementedError("at most 3 stages (the reference defines scales for stage1..3 only)")
        self.refine, self.share_cr, self.ndepths = refine, share_cr, tuple(ndepths)
        self.depth_interals_ratio, self.grad_method = tuple(depth_interals_ratio), grad_method
        self.arch_mode, self.cr_base_chs, self.num_stage = arch_mode, tuple(cr_base_chs), len(ndepths)
        self.storage = storage
        self.stage_infos = {"stage1": {"scale": 4.0}, "stage2": {"scale": 2.0}, "stage3": {"scale": 1.0}}
        self.feature = FeatureNet(base_channels=8, arch_mode=arch_mode, storage=storage)
        self.stage_net = StageNet(num_mvs_stages=len(ndepths), storage=storage)
        if share_cr:
            raise NotImplementedError("share_cr=True cannot work in the reference either: one CostRegNet cannot take "
                                      "32-, 16- and 8-channel volumes (models/model.py:130)")
        self.cost_regularization = nn.ModuleList([CostRegNet(in_channels=self.feature.out_channels[i],
                                                             base_channels=self.cr_base_chs[i], storage=storage)
                                                  for i in range(self.num_stage)])
        if refine:
            self.refine_network = Refinement()

    def engine(self, dev) -> CascadeEngine:
        self._check_cache()
        if self._cache is None:
            mw = W.pack_model(self.state_dict(), self.num_stage, dev)
            rw = W.pack_refinement({k: v.detach() for k, v in self.state_dict().items()}, "refine_network", dev) if self.refine else None
            object.__setattr__(self, "_cache", CascadeEngine(mw, self.ndepths, self.depth_interals_ratio, self.storage, dev,
                                                             refine_weights=rw))
        return self._cache

    def forward(self, imgs, proj_matrices, depth_values, gt_depths=None, temperature=0.001):
        """imgs [B,N,3,H,W] (fp32 in [0,1], or uint8: divided by 255 on the device), proj_matrices {"stageK": [B,N,2,4,4]},
        depth_values [B,Dtot] -> the reference's output dict (models/model.py:140-223): per-stage dicts + top-level copies of
        the last stage + refined_depth.

        The ~70 launches of a forward are replayed as ONE CUDA graph, captured the first time an input signature (shapes,
        temperature) is seen; ``CDS_GRAPH=0`` or a caller that is itself capturing falls back to launch-by-launch.  The
        returned tensors are the caller's own (one device copy of the packed result maps), never the engine's buffers."""
        self._require_eval()
        if gt_depths is not None:
            raise NotImplementedError("gt_depths is a training input (out of scope)")
        dev = _dev(imgs)
        eng = self.engine(dev)
        with torch.cuda.device(dev):
            graphed = _USE_GRAPH and not torch.cuda.is_current_stream_capturing()
            out = (eng.forward_graph if graphed else eng.forward)(imgs, proj_matrices, depth_values, temperature)
            refined = out["refined_depth"].clone() if self.refine else None
            return eng.outputs_from(eng._out_pack.clone(), refined)


_USE_GRAPH = __import__("os").environ.get("CDS_GRAPH", "1") != "0"


# ------------------------------------------------------------------------------------------------
PATCH_LEVELS = ("leaf", "ops", "stage", "model")


def patch(models_model=None, models_module=None, level="model"):
    """Rebind the hot-path names inside an imported reference ``models.model`` / ``models.module`` so that the reference's
    own code constructs and calls the CUDA implementations (the names are those models/model.py:4-6 imports).

    level="leaf":  only the leaf operators: ``DynamicConv`` (inside the reference's own ``Conv2d`` / ``FeatureNet``),
                   ``CostRegNet``, ``Refinement``, ``homo_warping_3D``, ``depth_regression``, ``conf_regression``.
    level="ops":   additionally ``FeatureNet``: the reference's own ``CDSMVSNet.forward`` (models/model.py:140-223) and ``StageNet.forward`` (:16-94) keep
                   driving the cascade; ``FeatureNet``, ``CostRegNet``, ``Refinement``, ``DynamicConv``, ``homo_warping_3D``,
                   ``depth_regression`` and ``conf_regression`` are the CUDA ones (state-dict keys unchanged).
    level="stage": additionally ``StageNet`` = the fused plane-sweep / visibility / aggregation kernels.
    level="model": additionally ``CDSMVSNet`` itself (feature batching across views, the whole cascade as one CUDA graph).
    Returns {module: {name: original}} so that ``unpatch`` can restore the reference."""
    if level not in PATCH_LEVELS:
        raise ValueError(f"patch level must be one of {PATCH_LEVELS}, got {level!r}")
    ops = [("homo_warping_3D", homo_warping_3D), ("depth_regression", depth_regression), ("conf_regression", conf_regression),
           ("CostRegNet", CostRegNet), ("Refinement", Refinement), ("DynamicConv", DynamicConv)]
    if level != "leaf":
        ops.append(("FeatureNet", FeatureNet))
    if level in ("stage", "model"):
        ops.append(("StageNet", StageNet))
    if level == "model":
        ops.append(("CDSMVSNet", CDSMVSNet))
    saved = {}
    for mod in (models_model, models_module):
        if mod is None:
            continue
        saved[mod] = {}
        for name, obj in ops:
            if hasattr(mod, name):   # only names the reference module itself binds
                saved[mod][name] = getattr(mod, name)
                setattr(mod, name, obj)
    return saved


def unpatch(saved):
    """Undo ``patch`` (its return value)."""
    for mod, names in saved.items():
        for name, obj in names.items():
            setattr(mod, name, obj)
