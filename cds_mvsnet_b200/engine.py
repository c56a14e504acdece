"""Host-side driver of the fused depth-inference cascade (the CUDA path of ``CDSMVSNet.forward``).

Everything numerical happens in libcds_b200.so; this module owns buffers and launch order only.
Reference flow being replaced: models/model.py:140-223 (cascade), :16-94 (stage), models/module.py:
236-267 (feature extractor), :305-315 (regulariser).  Layout of the feature-extractor batch: item
(side, v, b) with side 0 = the reference image seen with pair v's epipole, side 1 = source image v
(the reference recomputes the ref features per pair, models/model.py:154-161); so the first V*B
items are exactly the ``ref_fea [V,B,...]`` the cost-volume kernels want and the rest ``src_fea``.
"""
from __future__ import annotations

import ctypes
import os

import torch

from . import _lib
from ._lib import ACT_LRELU, ACT_NONE, ACT_TANH, call, ptr
from .weights import ModelWeights

STAGE_SCALE = (4, 2, 1)   # models/model.py:115-125
STAGE_CHANNELS = (32, 16, 8)


class Buffers:
    """Named persistent device buffers (static addresses => CUDA-graph friendly)."""

    def __init__(self, device):
        self.device = device
        self._t = {}

    def get(self, name, shape, dtype):
        shape = tuple(int(s) for s in shape)
        t = self._t.get(name)
        if t is None or t.shape != shape or t.dtype != dtype:
            t = torch.empty(shape, dtype=dtype, device=self.device)
            self._t[name] = t
        return t

    def nbytes(self):
        return sum(t.numel() * t.element_size() for t in self._t.values())


def _ksizes(ks):
    return (ctypes.c_int * len(ks))(*ks)


def kcall(tag, flops, nbytes, name, *args):
    """Launch `name` with a profiling tag and its ALGORITHMIC cost (flops, bytes) -- see DESIGN.md section 5."""
    _lib.set_tag(tag, (float(flops), float(nbytes)))
    call(name, *args)


def _esize(dt):
    return 2 if dt == torch.float16 else 4


class FeatureExtractor:
    """FeatureNet (models/module.py:201-267) for a batch of n images in ~17 launches."""

    N_STATS = 13

    def __init__(self, fw, storage=torch.float16, use_tc=True):
        self.fw = fw
        self.storage = storage
        self.dt = _lib.dtype_code(storage)
        self.use_tc = use_tc and os.environ.get("CDS_USE_TC", "1") != "0" and os.environ.get("CDS_TC_DYN", "1") != "0"
        # hi/lo fp16 activation planes (~22-bit activations) into conv10/conv11, the layers the depth output is most sensitive
        # to: without it the worst-case "noise" input exceeds north_star's 1e-3 on some seeds (seed 4
        # 2.6e-3 -> 9e-4).  CDS_SPLIT=0 turns it off (diagnostics only).
        self.split_precision = os.environ.get("CDS_SPLIT", "1") != "0"
        self.use_tc2d = use_tc and os.environ.get("CDS_USE_TC", "1") != "0" and os.environ.get("CDS_TC_CONV2D", "1") != "0"
        self.use_u8 = os.environ.get("CDS_U8_CONV00", "1") != "0"   # 8-bit images straight into conv00 (pixel-pair operand slots)
        self.use_rows = os.environ.get("CDS_S2_ROWS", "1") != "0"   # stride-2 layers on the row-streaming kernel (conv2d_s2rows.cu)
        # inner1/inner2 (1x1 conv over the concatenation): the 2x2-block CUDA-core form (fp32 math, 0.154 / 0.133 ms at cfg2) beats the
        # gather-form tensor-core kernel (0.289 / 0.156 ms) on this 24- / 48-deep contraction; CDS_TC_INNER=1 selects the latter
        self.tc_inner = os.environ.get("CDS_TC_INNER", "0") != "0"
        self._buf = None
        self.pairs = None   # (V, B) when the batch is the cascade's (side, v, b) pair batch
        self.share_ref = os.environ.get("CDS_SHARE_REF", "1") != "0"
        # trunk layers on the row-folded persistent kernel (csrc/dynconv_kh.cu); CDS_DYN_KH=0 keeps them on the tap GEMM
        self.use_kh = os.environ.get("CDS_DYN_KH", "1") != "0"

    def _dyn(self, name, x, in_mode, img_index, in_stats, in_act, epi, epi_scale, n, H, W, T, out, out_stats, nc_sq,
             nc_mode, nc_abs, norm_curv=None, split_in=False, out_lo=None):
        w = self.fw.dyn[name]
        e = _esize(self.storage)
        px = n * H * W
        flops = 2.0 * sum(k * k for k in w.ksizes) * w.cin * (w.cout + 3) * px
        nbytes = px * (w.cin * ((1 if x.dtype == torch.uint8 else 4) if in_mode == 1 else e) + w.cout * e + 8)
        _lib.set_tag("feat." + name, (flops, float(nbytes)))
        ks = _ksizes(w.ksizes)
        if in_mode == 1 and x.dtype == torch.uint8:
            # 8-bit images: pixel-pair operand slots, four taps per MMA (csrc/dynconv_kh.cu, cds_dynamic_conv_kh_u8)
            n_images = x.shape[0]
            pad = _lib.LIB.load().cds_dynamic_conv_kh_u8_pad()
            px2 = self._buf.get("f.img_px2", (n_images, H, W + 2 * pad, 8), torch.float16)
            call("cds_image_u8_to_px2", ptr(x), n_images, H, W, ptr(px2))
            pv, pb = self.pairs if (self.pairs is not None and self.share_ref) else (0, 0)
            call("cds_dynamic_conv_kh_u8", ptr(px2), n_images, ptr(img_index), ptr(epi), float(epi_scale), ptr(w.kh_u8), ptr(w.gate),
                 n, w.cout, H, W, len(w.ksizes), ks, float(T), ptr(out), ptr(out_lo), ptr(out_stats), ptr(norm_curv), ptr(nc_sq),
                 nc_mode, ptr(nc_abs), pv, pb)
            return
        if (self.use_tc and w.tc is not None and self.storage == torch.float16
                and _lib.LIB.load().cds_dynamic_conv_tc_supported(max(8, w.cin), w.cout, H, W, len(w.ksizes), ks)):
            n_images = n
            if in_mode == 1:   # planar fp32 images -> fp16 [*,H,W,8] once per forward
                n_images = x.shape[0]
                img8 = self._buf.get("f.img8", (n_images, H, W, 8), torch.float16)
                call("cds_image_to_nhwc8", ptr(x), n_images, H, W, ptr(img8))
                x = img8
            if (self.use_kh and w.kh is not None
                    and _lib.LIB.load().cds_dynamic_conv_kh_supported(max(8, w.cin), w.cout, H, W, len(w.ksizes), ks)):
                pv, pb = self.pairs if (in_mode == 1 and self.pairs is not None and self.share_ref) else (0, 0)
                call("cds_dynamic_conv_kh", ptr(x), n_images, ptr(img_index), ptr(in_stats), in_act, ptr(epi), float(epi_scale),
                     ptr(w.kh), ptr(w.bias), ptr(w.gate), n, max(8, w.cin), w.cout, H, W, len(w.ksizes), ks, float(T),
                     int(split_in), ptr(out), ptr(out_lo), ptr(out_stats), ptr(norm_curv), ptr(nc_sq), nc_mode, ptr(nc_abs), pv, pb)
                return
            if in_mode == 1 and self.pairs is not None and self.share_ref and not split_in:
                V, B = self.pairs   # the reference image's convolutions are shared by its V pairs
                call("cds_dynamic_conv_tc_pairs", ptr(x), n_images, ptr(img_index), ptr(epi), float(epi_scale), ptr(w.tc), ptr(w.bias),
                     ptr(w.gate), V, B, max(8, w.cin), w.cout, H, W, len(w.ksizes), ks, float(T), ptr(out), ptr(out_lo), ptr(out_stats),
                     ptr(norm_curv), ptr(nc_sq), nc_mode, ptr(nc_abs))
                return
            call("cds_dynamic_conv_tc", ptr(x), n_images, ptr(img_index), ptr(in_stats), in_act, ptr(epi), float(epi_scale),
                 ptr(w.tc), ptr(w.bias), ptr(w.gate), n, max(8, w.cin), w.cout, H, W, len(w.ksizes), ks, float(T),
                 int(split_in), ptr(out), ptr(out_lo), ptr(out_stats),
                 ptr(norm_curv), ptr(nc_sq), nc_mode, ptr(nc_abs))
            return
        assert not split_in and out_lo is None, "split-precision activations exist on the tensor-core path only"
        call("cds_dynamic_conv", ptr(x), in_mode, ptr(img_index), ptr(in_stats), in_act, ptr(epi), float(epi_scale),
             ptr(w.w_att), ptr(w.w_conv), ptr(w.bias), ptr(w.gate), n, w.cin, w.cout, H, W, len(w.ksizes),
             _ksizes(w.ksizes), float(T), self.dt, ptr(out), ptr(out_stats), ptr(norm_curv), ptr(nc_sq), nc_mode,
             ptr(nc_abs))

    def _down(self, name, x, x_stats, w32, n, cin, cout, H, W, out, out_stats, split, x_lo=None):
        """3x3 stride-2 conv of FeatureNet.downsample1/2 (out: [1 or 2 planes, n, H/2, W/2, cout]; x_lo: residual plane of x)."""
        e = _esize(self.storage)
        Ho, Wo = (H + 1) // 2, (W + 1) // 2
        _lib.set_tag("feat." + name, (2.0 * 9 * cin * cout * n * Ho * Wo, float(n * (H * W * cin + Ho * Wo * cout) * e)))
        if (self.use_tc2d and self.use_rows and self.storage == torch.float16 and name in self.fw.rows
                and _lib.LIB.load().cds_conv2d_3x3s2_rows_supported(cin, cout, H, W)):
            call("cds_conv2d_3x3s2_rows", ptr(x), ptr(x_lo), ptr(x_stats), ACT_LRELU, ptr(self.fw.rows[name]), n, cin, cout, H, W,
                 ptr(out[0]), ptr(out[1]) if split else None, ptr(out_stats))
            return
        if (self.use_tc2d and self.storage == torch.float16 and name in self.fw.tc
                and _lib.LIB.load().cds_conv2d_3x3s2_tc_supported(cin, cout)):
            call("cds_conv2d_3x3s2_tc", ptr(x), ptr(x_lo), ptr(x_stats), ACT_LRELU, ptr(self.fw.tc[name]), n, cin, cout, H, W, ptr(out[0]),
                 ptr(out[1]) if split else None, ptr(out_stats))
            return
        assert x_lo is None, "split-precision activations exist on the tensor-core path only"
        call("cds_conv2d_3x3s2", ptr(x), ptr(x_stats), ACT_LRELU, ptr(w32), n, cin, cout, H, W, self.dt, ptr(out[0]),
             ptr(out[1]) if split else None, ptr(out_stats))

    def _inner(self, name, a, a_stats, a_act, b, b_stats, w32, n, ca, cb, cout, H, W, out, out_stats):
        """1x1 conv over cat(nearest-up2(a), b) of FeatureNet.inner1/2."""
        e = _esize(self.storage)
        _lib.set_tag("feat." + name, (2.0 * (ca + cb) * cout * n * H * W, float(n * ((H // 2) * (W // 2) * ca + H * W * (cb + cout)) * e)))
        if (self.use_tc2d and self.tc_inner and self.storage == torch.float16 and name in self.fw.tc
                and _lib.LIB.load().cds_conv2d_1x1_cat_tc_supported(ca, cb, cout)):
            call("cds_conv2d_1x1_cat_tc", ptr(a), ptr(a_stats), a_act, ptr(b), ptr(b_stats), ACT_LRELU, ptr(self.fw.tc[name]),
                 n, ca, cb, cout, H, W, ptr(out), ptr(out_stats))
            return
        call("cds_conv2d_1x1_cat", ptr(a), ptr(a_stats), a_act, ptr(b), ptr(b_stats), ACT_LRELU, ptr(w32), n, ca, cb, cout, H, W,
             self.dt, ptr(out), ptr(out_stats))

    def _split_ok(self, H, W):
        """The precise trunk (split-precision storage + operands in every layer between the image and the stage-1 feature) needs
        the tensor-core kernels of all of them."""
        if not (self.use_tc and self.use_tc2d and self.split_precision and self.storage == torch.float16):
            return False
        lib = _lib.LIB.load()
        for name, (h, w) in (("conv01", (H, W)), ("conv10", (H // 2, W // 2)), ("conv20", (H // 4, W // 4))):
            d = self.fw.dyn[name]
            if d.tc is None or not lib.cds_dynamic_conv_tc_supported(max(8, d.cin), d.cout, h, w, len(d.ksizes), _ksizes(d.ksizes)):
                return False
        return True

    def u8_ok(self, H, W) -> bool:
        """conv00 can take the 8-bit images as they are (pixel-pair slots): the row-folded tensor-core path with its operand image."""
        w = self.fw.dyn["conv00"]
        return (self.use_u8 and self.use_tc and self.use_kh and self.storage == torch.float16 and w.kh_u8 is not None
                and bool(_lib.LIB.load().cds_dynamic_conv_kh_supported(8, w.cout, H, W, len(w.ksizes), _ksizes(w.ksizes))))

    def run(self, buf: Buffers, imgs, img_index, epipoles, n, H, W, temperature, pairs=None, after_stage1=None):
        """imgs: planar fp32 [*,3,H,W] (or uint8 when ``u8_ok``); img_index int32 [n]; epipoles fp32 [n,2].
        Returns {stage: (fea [n,h,w,C] storage dtype, nc_sq [n,h,w] fp32, nc_abs [n,h,w] fp32)}."""
        st, dt = self.storage, self.dt
        self._buf = buf
        self.pairs = pairs
        imgs = imgs.reshape(-1, 3, H, W)
        H2, W2, H4, W4 = H // 2, W // 2, H // 4, W // 4
        e = _esize(st)
        stats = buf.get("f.stats", (self.N_STATS, n, 32, 2), torch.float64)
        stats.zero_()
        # the kernels index statistics densely as [n][C][2]: give each layer its own dense view
        sviews = {}

        def sv(i, c):
            if i not in sviews:
                sviews[i] = stats[i].reshape(-1)[: n * c * 2].view(n, c, 2)
            return sviews[i]

        f32 = torch.float32
        # Precise trunk: every activation between the image and the stage-1 feature is kept as two fp16 planes (value + rounding
        # residual = ~22 bits) and fed to the tensor cores as twice as many K slabs, and every weight carries its residual in
        # extra N columns.  The stage-1 depth of the chaotic "noise" input amplifies each rounding of this chain ~9x into the
        # final depth; the heads of stages 2 / 3 (inner1/2, out2/3) are not on it and stay single-plane.
        split = self._split_ok(H, W)
        P2 = 2 if split else 1
        lo = lambda t: t[1] if split else None
        raw00 = buf.get("f.raw00", (P2, n, H, W, 8), st)
        raw01 = buf.get("f.raw01", (P2, n, H, W, 8), st)
        rawd1 = buf.get("f.rawd1", (P2, n, H2, W2, 16), st)
        raw10 = buf.get("f.raw10", (P2, n, H2, W2, 16), st)
        raw11 = buf.get("f.raw11", (P2, n, H2, W2, 16), st)
        rawd2 = buf.get("f.rawd2", (P2, n, H4, W4, 32), st)
        raw20 = buf.get("f.raw20", (P2, n, H4, W4, 32), st)
        raw21 = buf.get("f.raw21", (P2, n, H4, W4, 32), st)
        rawo1 = buf.get("f.rawo1", (P2, n, H4, W4, 32), st)
        fea1 = buf.get("f.fea1", (n, H4, W4, 32), f32 if split else st)   # precise: fp32 stage-1 feature
        rawi1 = buf.get("f.rawi1", (n, H2, W2, 16), st)
        rawo2 = buf.get("f.rawo2", (n, H2, W2, 16), st)
        fea2 = buf.get("f.fea2", (n, H2, W2, 16), st)
        rawi2 = buf.get("f.rawi2", (n, H, W, 8), st)
        rawo3 = buf.get("f.rawo3", (n, H, W, 8), st)
        fea3 = buf.get("f.fea3", (n, H, W, 8), st)
        ncsq = [buf.get(f"f.ncsq{i}", (n, h, w), f32) for i, (h, w) in enumerate(((H4, W4), (H2, W2), (H, W)))]
        ncab = [buf.get(f"f.ncabs{i}", (n, h, w), f32) for i, (h, w) in enumerate(((H4, W4), (H2, W2), (H, W)))]
        T = temperature
        fw = self.fw
        # full resolution
        self._dyn("conv00", imgs, 1, img_index, None, ACT_NONE, epipoles, 1.0, n, H, W, T, raw00[0], sv(0, 8), ncsq[2], 0, None,
                  out_lo=lo(raw00))
        self._dyn("conv01", raw00, 0, None, sv(0, 8), ACT_LRELU, epipoles, 1.0, n, H, W, T, raw01[0], sv(1, 8), ncsq[2], 1, None,
                  split_in=split, out_lo=lo(raw01))
        # 1/2 resolution
        self._down("downsample1", raw01[0], sv(1, 8), fw.downsample1, n, 8, 16, H, W, rawd1, sv(2, 16), split, x_lo=lo(raw01))
        self._dyn("conv10", rawd1, 0, None, sv(2, 16), ACT_LRELU, epipoles, 0.5, n, H2, W2, T, raw10[0], sv(3, 16), ncsq[1], 0, None,
                  split_in=split, out_lo=lo(raw10))
        self._dyn("conv11", raw10, 0, None, sv(3, 16), ACT_LRELU, epipoles, 0.5, n, H2, W2, T, raw11[0], sv(4, 16), ncsq[1], 1, None,
                  split_in=split, out_lo=lo(raw11))
        # 1/4 resolution
        self._down("downsample2", raw11[0], sv(4, 16), fw.downsample2, n, 16, 32, H2, W2, rawd2, sv(5, 32), split, x_lo=lo(raw11))
        self._dyn("conv20", rawd2, 0, None, sv(5, 32), ACT_LRELU, epipoles, 0.25, n, H4, W4, T, raw20[0], sv(6, 32), ncsq[0], 0, None,
                  split_in=split, out_lo=lo(raw20))
        self._dyn("conv21", raw20, 0, None, sv(6, 32), ACT_LRELU, epipoles, 0.25, n, H4, W4, T, raw21[0], sv(7, 32), ncsq[0], 1, None,
                  split_in=split, out_lo=lo(raw21))
        # stage-1 output
        self._dyn("out1", raw21, 0, None, sv(7, 32), ACT_LRELU, epipoles, 0.25, n, H4, W4, T, rawo1[0], sv(8, 32), ncsq[0], 2, ncab[0],
                  split_in=split, out_lo=lo(rawo1))
        fea1_16 = None
        if split:
            # the similarity-entropy sweep of stage 1 only feeds the visibility net: it reads an fp16 copy of the feature (half the
            # gather bytes; measured no change of the depth error at cfg2: 8.27e-4 vs 8.31e-4).  CDS_S0_ENTROPY_F16=0: fp32.
            if os.environ.get("CDS_S0_ENTROPY_F16", "1") != "0":
                fea1_16 = buf.get("f.fea1_16", (n, H4, W4, 32), st)
            kcall("feat.act1", 0, n * H4 * W4 * 32 * (2 * e + 4), "cds_instnorm_act_split_f32", ptr(rawo1[0]), ptr(rawo1[1]), ptr(sv(8, 32)),
                  ACT_TANH, n, 32, H4, W4, ptr(fea1), ptr(fea1_16))
        else:
            kcall("feat.act1", 0, 2 * n * H4 * W4 * 32 * e, "cds_instnorm_act", ptr(rawo1[0]), ptr(sv(8, 32)), ACT_TANH, n, 32, H4, W4, dt, ptr(fea1))
        if after_stage1 is not None:   # the stage-1 feature is complete: the caller may start stage 1 on another stream
            after_stage1((fea1, ncsq[0], ncab[0], fea1_16))
        # stage-2 output: inner1 over cat(up2(conv21), conv11)
        self._inner("inner1", raw21[0], sv(7, 32), ACT_LRELU, raw11[0], sv(4, 16), fw.inner1, n, 32, 16, 16, H2, W2, rawi1, sv(9, 16))
        self._dyn("out2", rawi1, 0, None, sv(9, 16), ACT_LRELU, epipoles, 0.5, n, H2, W2, T, rawo2, sv(10, 16), ncsq[1], 2, ncab[1])
        kcall("feat.act2", 0, 2 * n * H2 * W2 * 16 * e, "cds_instnorm_act", ptr(rawo2), ptr(sv(10, 16)), ACT_TANH, n, 16, H2, W2, dt, ptr(fea2))
        # stage-3 output: inner2 over cat(up2(stage-2 feature), conv01)
        self._inner("inner2", fea2, None, ACT_NONE, raw01[0], sv(1, 8), fw.inner2, n, 16, 8, 8, H, W, rawi2, sv(11, 8))
        self._dyn("out3", rawi2, 0, None, sv(11, 8), ACT_LRELU, epipoles, 1.0, n, H, W, T, rawo3, sv(12, 8), ncsq[2], 2, ncab[2])
        kcall("feat.act3", 0, 2 * n * H * W * 8 * e, "cds_instnorm_act", ptr(rawo3), ptr(sv(12, 8)), ACT_TANH, n, 8, H, W, dt, ptr(fea3))
        return {0: (fea1, ncsq[0], ncab[0], fea1_16), 1: (fea2, ncsq[1], ncab[1]), 2: (fea3, ncsq[2], ncab[2])}


class Regulariser:
    """CostRegNet (models/module.py:270-315) + prob head on a channels-last volume."""

    def __init__(self, cw, storage=torch.float16, use_tc=True):
        self.cw = cw
        self.storage = storage
        self.dt = _lib.dtype_code(storage)
        self.use_tc = use_tc and os.environ.get("CDS_USE_TC", "1") != "0" and os.environ.get("CDS_TC_CONV3D", "1") != "0"
        self.use_gtc = use_tc and os.environ.get("CDS_USE_TC", "1") != "0" and os.environ.get("CDS_TC_GATHER", "1") != "0"
        self.use_roll = use_tc and os.environ.get("CDS_USE_TC", "1") != "0" and os.environ.get("CDS_TC_ROLL", "1") != "0"
        self.tag = "cr"

    def _conv(self, name, x, B, D, H, W, stride, out):
        l = self.cw.layers[name]
        e = _esize(self.storage)
        m_in, m_out = B * D * H * W, out.numel() // l.cout
        _lib.set_tag(f"{self.tag}.{name}", (2.0 * 27 * l.cin * l.cout * m_out, float((l.cin * m_in + l.cout * m_out) * e)))
        if (self.use_roll and stride == 1 and self.storage == torch.float16 and "roll" in l.extra
                and _lib.LIB.load().cds_conv3d_k3_roll_supported(l.cin, l.cout, D, H, W)):
            call("cds_conv3d_k3_roll", ptr(x), ptr(l.extra["roll"]), ptr(l.bias), B, l.cin, l.cout, D, H, W, 1, ptr(out))
            return
        if (self.use_tc and stride == 1 and self.storage == torch.float16 and "tc" in l.extra
                and _lib.LIB.load().cds_conv3d_k3_tc_supported(l.cin, l.cout, D, H, W, stride)):
            call("cds_conv3d_k3_tc", ptr(x), ptr(l.extra["tc"]), ptr(l.bias), B, l.cin, l.cout, D, H, W, 1, ptr(out))
            return
        if (self.use_gtc and self.storage == torch.float16 and "gtc" in l.extra
                and _lib.LIB.load().cds_conv3d_k3_gtc_supported(l.cin, l.cout, stride)):
            call("cds_conv3d_k3_gtc", ptr(x), ptr(l.extra["gtc"]), ptr(l.bias), B, l.cin, l.cout, D, H, W, stride, 1, ptr(out))
            return
        call("cds_conv3d_k3", ptr(x), ptr(l.w), ptr(l.bias), B, l.cin, l.cout, D, H, W, stride, 1, self.dt, ptr(out))

    def _deconv(self, name, x, skip, B, D, H, W, out):
        l = self.cw.layers[name]
        e = _esize(self.storage)
        m_in = B * D * H * W
        _lib.set_tag(f"{self.tag}.{name}", (2.0 * 27 * l.cin * l.cout * m_in, float((l.cin * m_in + 2 * l.cout * 8 * m_in) * e)))
        if (self.use_tc and self.storage == torch.float16 and "tc" in l.extra
                and _lib.LIB.load().cds_deconv3d_k3s2_tc_supported(l.cin, l.cout, D, H, W)):
            call("cds_deconv3d_k3s2_tc", ptr(x), ptr(l.extra["tc"]), ptr(l.bias), ptr(skip), B, l.cin, l.cout, D, H, W, ptr(out))
            return
        if (self.use_gtc and self.storage == torch.float16 and "gtc" in l.extra
                and _lib.LIB.load().cds_deconv3d_k3s2_gtc_supported(l.cin, l.cout)):
            call("cds_deconv3d_k3s2_gtc", ptr(x), ptr(l.extra["gtc"]), ptr(l.bias), ptr(skip), B, l.cin, l.cout, D, H, W, ptr(out))
            return
        call("cds_deconv3d_k3s2", ptr(x), ptr(l.w), ptr(l.bias), ptr(skip), B, l.cin, l.cout, D, H, W, self.dt, ptr(out))

    def run(self, buf: Buffers, tag, volume, B, D, H, W, samples=None, depth=None, conf=None):
        """volume [B,C/8,D,H,W,8] (channel-blocked) -> fp32 logits [B,D,H,W].  With samples / depth / conf given and the rolling
        prob head available, the tail (softmax over D, depth_regression, conf_regression: models/model.py:85-92) runs fused in
        the prob head's epilogue and ``self.fused_tail`` is set."""
        self.fused_tail = False
        if D % 8 or H % 8 or W % 8:
            raise RuntimeError(f"CostRegNet needs D, H, W divisible by 8 (got {D}x{H}x{W}); the reference fails the "
                               "same way at its skip additions (models/module.py:310-312)")
        st = self.storage
        self.tag = tag
        b = self.cw.layers["conv0"].cout
        D2, H2, W2, D4, H4, W4, D8, H8, W8 = D // 2, H // 2, W // 2, D // 4, H // 4, W // 4, D // 8, H // 8, W // 8
        g = lambda n, s: buf.get(f"{tag}.{n}", (s[0], s[4] // 8) + tuple(s[1:4]) + (8,), st)   # [B, C/8, D, H, W, 8]
        c0 = g("c0", (B, D, H, W, b))
        c1 = g("c1", (B, D2, H2, W2, 2 * b))
        c2 = g("c2", (B, D2, H2, W2, 2 * b))
        c3 = g("c3", (B, D4, H4, W4, 4 * b))
        c4 = g("c4", (B, D4, H4, W4, 4 * b))
        c5 = g("c5", (B, D8, H8, W8, 8 * b))
        c6 = g("c6", (B, D8, H8, W8, 8 * b))
        u7 = g("u7", (B, D4, H4, W4, 4 * b))
        u9 = g("u9", (B, D2, H2, W2, 2 * b))
        u11 = g("u11", (B, D, H, W, b))
        logits = buf.get(f"{tag}.logits", (B, D, H, W), torch.float32)
        self._conv("conv0", volume, B, D, H, W, 1, c0)
        self._conv("conv1", c0, B, D, H, W, 2, c1)
        self._conv("conv2", c1, B, D2, H2, W2, 1, c2)
        self._conv("conv3", c2, B, D2, H2, W2, 2, c3)
        self._conv("conv4", c3, B, D4, H4, W4, 1, c4)
        self._conv("conv5", c4, B, D4, H4, W4, 2, c5)
        self._conv("conv6", c5, B, D8, H8, W8, 1, c6)
        self._deconv("conv7", c6, c4, B, D8, H8, W8, u7)
        self._deconv("conv9", u7, c2, B, D4, H4, W4, u9)
        self._deconv("conv11", u9, c0, B, D2, H2, W2, u11)
        m = B * D * H * W
        if (self.use_roll and st == torch.float16 and self.cw.prob_roll is not None
                and _lib.LIB.load().cds_conv3d_k3_roll_supported(b, 1, D, H, W)):
            # CDS_FUSED_TAIL=1: softmax + regression in the prob head's epilogue (cds_prob_head_regress).  Measured at cfg2: the
            # three fused launches cost 0.09 ms MORE than prob head + cds_softmax_regress (the prob head is bound by its
            # epilogue warps, which the exp / hypothesis loads lengthen), so the separate tail stays the default.
            if samples is not None and depth is not None and conf is not None and os.environ.get("CDS_FUSED_TAIL", "0") == "1":
                kcall(f"{tag}.prob_regress_tail", 2.0 * 27 * b * m, m * (b * _esize(st) + 8) + 8 * B * H * W, "cds_prob_head_regress",
                      ptr(u11), ptr(self.cw.prob_roll), ptr(samples), B, D, H, W, ptr(logits), ptr(depth), ptr(conf))
                self.fused_tail = True
            else:
                kcall(f"{tag}.prob", 2.0 * 27 * b * m, m * (b * _esize(st) + 4), "cds_conv3d_k3_roll", ptr(u11), ptr(self.cw.prob_roll), None,
                      B, b, 1, D, H, W, 0, ptr(logits))
        elif (self.use_tc and st == torch.float16 and self.cw.prob_tc is not None
                and _lib.LIB.load().cds_conv3d_k3_tc_supported(b, 1, D, H, W, 1)):
            kcall(f"{tag}.prob", 2.0 * 27 * b * m, m * (b * _esize(st) + 4), "cds_conv3d_k3_tc", ptr(u11), ptr(self.cw.prob_tc), None,
                  B, b, 1, D, H, W, 0, ptr(logits))
        else:
            kcall(f"{tag}.prob", 2.0 * 27 * b * m, m * (b * _esize(st) + 4), "cds_prob_conv", ptr(u11), ptr(self.cw.prob), B, b, D,
                  H, W, self.dt, ptr(logits))
        return logits


class Refiner:
    """Refinement network (models/module.py:318-370) on csrc/refine.cu: seven launches, fp32 planar."""

    def __init__(self, rw):
        self.rw = rw

    def run(self, buf: Buffers, img, depth0, lo, hi, post, B, H, W):
        """img [B,3,H,W] fp32, depth0 [B,H/2,W/2], lo / hi / post [B] (post may be None) -> refined [B,H,W] (a Buffers tensor)."""
        if H % 2 or W % 2:
            raise RuntimeError(f"Refinement needs even H, W (got {H}x{W})")
        h, w, f32, rw = H // 2, W // 2, torch.float32, self.rw
        dn = buf.get("rf.depth_n", (B, h, w), f32)
        c0 = buf.get("rf.conv0", (B, 8, H, W), f32)
        c1 = buf.get("rf.conv1", (B, 8, h, w), f32)
        c2 = buf.get("rf.conv2", (B, 8, h, w), f32)
        dc = buf.get("rf.deconv", (B, 8, H, W), f32)
        c3 = buf.get("rf.conv3", (B, 8, H, W), f32)
        out = buf.get("rf.refined", (B, H, W), f32)
        P, p = B * H * W, B * h * w
        kcall("refine.prescale", 0, 8 * p, "cds_refine_prescale", ptr(depth0), ptr(lo), ptr(hi), B, h, w, ptr(dn))
        kcall("refine.conv0", 2.0 * 27 * 8 * P, 4 * P * 11, "cds_conv2d_3x3_f32", ptr(img), None, ptr(rw["conv0"][0]), ptr(rw["conv0"][1]),
              B, 3, 0, 8, H, W, 1, ptr(c0))
        kcall("refine.conv1", 2.0 * 9 * 8 * p, 4 * p * 9, "cds_conv2d_3x3_f32", ptr(dn), None, ptr(rw["conv1"][0]), ptr(rw["conv1"][1]),
              B, 1, 0, 8, h, w, 1, ptr(c1))
        kcall("refine.conv2", 2.0 * 72 * 8 * p, 4 * p * 16, "cds_conv2d_3x3_f32", ptr(c1), None, ptr(rw["conv2"][0]), ptr(rw["conv2"][1]),
              B, 8, 0, 8, h, w, 1, ptr(c2))
        kcall("refine.deconv", 2.0 * 18 * 8 * P, 4 * (p * 8 + P * 8), "cds_deconv2d_k3s2_f32", ptr(c2), ptr(rw["deconv"][0]),
              ptr(rw["deconv"][1]), B, 8, h, w, ptr(dc))
        kcall("refine.conv3", 2.0 * 144 * 8 * P, 4 * P * 24, "cds_conv2d_3x3_f32", ptr(dc), ptr(c0), ptr(rw["conv3"][0]), ptr(rw["conv3"][1]),
              B, 8, 8, 8, H, W, 1, ptr(c3))
        kcall("refine.final", 2.0 * 72 * P, 4 * P * 9 + 4 * p, "cds_refine_final", ptr(c3), ptr(rw["res"]), ptr(dn), ptr(lo), ptr(hi),
              ptr(post), B, h, w, ptr(out))
        return out


class CascadeEngine:
    """The whole CDSMVSNet.forward (refine=False, eval) on the CUDA kernels."""

    def __init__(self, weights: ModelWeights, ndepths, ratios, storage=torch.float16, device=None, refine_weights=None):
        self.w = weights
        self.refiner = Refiner(refine_weights) if refine_weights is not None else None   # refine=True (models/model.py:209-216)
        self.ndepths = tuple(int(d) for d in ndepths)
        self.ratios = tuple(float(r) for r in ratios)
        self.storage = storage
        self.dt = _lib.dtype_code(storage)
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.buf = Buffers(self.device)
        self.features = FeatureExtractor(weights.feature, storage)
        self.regs = [Regulariser(cw, storage) for cw in weights.costreg]
        self.use_tc = os.environ.get("CDS_USE_TC", "1") != "0" and os.environ.get("CDS_TC_VIS", "1") != "0"
        self.overlap = os.environ.get("CDS_OVERLAP", "1") != "0"   # stage 1 on a side stream under the feature heads
        self._side = torch.cuda.Stream(device=self.device) if self.device.type == "cuda" else None
        self.launches = 0

    def clone(self) -> "CascadeEngine":
        """A second cascade over the SAME packed weights with buffers, side stream and CUDA graph of its own: lets two depth maps
        be in flight on two streams (the tails and small grids of one map's kernels leave SMs idle that the other fills:
        measured +6 % maps/s at cfg2, streaming.DepthMapStream)."""
        return CascadeEngine(self.w, self.ndepths, self.ratios, self.storage, self.device,
                             refine_weights=self.refiner.rw if self.refiner is not None else None)

    # -- pieces -------------------------------------------------------------------------------
    def camera_setup(self, proj_matrices, B, N):
        n_st = len(self.ndepths)
        keys = [f"stage{s + 1}" for s in range(n_st)]
        epi_key = "stage3"   # models/model.py:152 always uses the stage-3 cameras for the epipoles
        mats = [proj_matrices[k] for k in keys]
        if epi_key in keys:
            epi_idx = keys.index(epi_key)
        else:
            mats.append(proj_matrices[epi_key])
            epi_idx = len(mats) - 1
        mats = [m.to(device=self.device, dtype=torch.float32).contiguous() for m in mats]
        for m in mats:
            if tuple(m.shape) != (B, N, 2, 4, 4):
                raise AssertionError(f"proj_matrices entries must be [B,N,2,4,4], got {tuple(m.shape)}")
        V = N - 1
        coef = self.buf.get("cam.coef", (len(mats), B, V, 12), torch.float32)
        epi = self.buf.get("cam.epi", (2, V, B, 2), torch.float32)
        arr = (ctypes.c_void_p * len(mats))(*[m.data_ptr() for m in mats])
        kcall("camera_setup", 0, 0, "cds_camera_setup", arr, len(mats), epi_idx, B, N, ptr(coef), ptr(epi))
        self._keep = mats
        return coef, epi

    def stage(self, s, feats_s, coef_s, depth_values, prev_depth, B, V, H, W):
        """One StageNet.forward (models/model.py:16-94) at stage index s."""
        D, scale, C = self.ndepths[s], STAGE_SCALE[s], STAGE_CHANNELS[s]
        h, w = H // scale, W // scale
        buf, f32 = self.buf, torch.float32
        fea, ncsq, ncabs = feats_s[:3]
        fea16 = feats_s[3] if len(feats_s) > 3 else None   # optional fp16 copy of a precise (fp32) stage-1 feature
        VB = V * B
        samples = buf.get(f"s{s}.samples", (B, D, h, w), f32)
        hp, wp = (prev_depth.shape[1], prev_depth.shape[2]) if prev_depth is not None else (0, 0)
        e, P = _esize(self.storage), B * h * w
        kcall(f"s{s}.hypotheses", 0, 4 * D * P + 4 * B * hp * wp, "cds_depth_hypotheses", ptr(depth_values), depth_values.shape[1],
              ptr(prev_depth), hp, wp, B, D, self.ratios[s], H, W, scale, ptr(samples))
        ref_fea, src_fea = fea[:VB], fea[VB:]
        # precise stage 1: its feature arrives in fp32 (FeatureExtractor.run) and both plane sweeps run their fp32-feature forms;
        # the volume goes to the regulariser as fp16 value (+ residual) planes
        fea_f32 = fea.dtype == torch.float32 and self.storage == torch.float16
        fdt, fe = (_lib.CDS_F32, 4) if fea_f32 else (self.dt, e)
        entropy = buf.get(f"s{s}.entropy", (V, B, h, w), f32)
        # the entropy only feeds the visibility net: fp16 features take the packed-half blend (cds_costvol_entropy_fast: entropy
        # to 5e-4 instead of 2e-4; depth error at cfg2 6.90e-4 -> 6.96e-4, 0.317 + 0.457 -> 0.278 + 0.411 ms)
        if fea16 is not None:
            kcall(f"s{s}.costvol_entropy", 2.0 * 9 * C * D * P * V, V * P * (2 * C * e + 4) + 4 * D * P, "cds_costvol_entropy_fast",
                  ptr(fea16[:VB]), ptr(fea16[VB:]), ptr(coef_s), ptr(samples), V, B, C, D, h, w, self.dt, ptr(entropy))
        else:
            kcall(f"s{s}.costvol_entropy", 2.0 * 9 * C * D * P * V, V * P * (2 * C * fe + 4) + 4 * D * P, "cds_costvol_entropy_fast",
                  ptr(ref_fea), ptr(src_fea), ptr(coef_s), ptr(samples), V, B, C, D, h, w, fdt, ptr(entropy))
        vis = buf.get(f"s{s}.vis", (V, B, h, w), f32)
        # precise stage 1: the fp32 CUDA-core visibility net (its fp16 tensor-core form costs 1.5e-4 of final depth error on the
        # noise input at cfg2 through the stage-to-stage amplification; at quarter resolution the fp32 form is 0.1 ms)
        vis_precise = fea_f32 and os.environ.get("CDS_S0_VIS_F32", "1") != "0"
        if (self.use_tc and not vis_precise and self.w.vis_tc and self.storage == torch.float16
                and _lib.LIB.load().cds_visnet_tc_supported(h, w)):
            wgt, fp = self.w.vis_tc[s]
            kcall(f"s{s}.visnet", 9824.0 * P * V, 12 * P * V, "cds_visnet_tc", ptr(entropy), ptr(ncabs[:VB]), ptr(wgt), ptr(fp),
                  VB, h, w, ptr(vis))
        else:
            kcall(f"s{s}.visnet", 9824.0 * P * V, 12 * P * V, "cds_visnet", ptr(entropy), ptr(ncabs[:VB]), ptr(self.w.vis[s]),
                  VB, h, w, ptr(vis))
        volume = buf.get(f"s{s}.volume", (B, C // 8, D, h, w, 8), self.storage)
        if fea_f32:
            kcall(f"s{s}.costvol_aggregate", 2.0 * 10 * C * D * P * V, V * P * (2 * C * fe + 4) + 4 * D * P + C * D * P * e,
                  "cds_costvol_aggregate_split", ptr(ref_fea), ptr(src_fea), ptr(coef_s), ptr(samples), ptr(vis), V, B, C, D, h, w,
                  ptr(volume), None)
        else:
            kcall(f"s{s}.costvol_aggregate", 2.0 * 10 * C * D * P * V, V * P * (2 * C * e + 4) + 4 * D * P + C * D * P * e,
                  "cds_costvol_aggregate", ptr(ref_fea), ptr(src_fea), ptr(coef_s), ptr(samples), ptr(vis), V, B, C, D, h, w, self.dt,
                  ptr(volume))
        nc = self._out_views[s][2]
        kcall(f"s{s}.nc_mean", 0, 4 * P * (2 * V + 1), "cds_nc_mean", ptr(ncsq[:VB]), ptr(ncsq[VB:]), V, B * h * w, ptr(nc))
        depth, conf = self._out_views[s][0], self._out_views[s][1]
        logits = self.regs[s].run(buf, f"s{s}.cr", volume, B, D, h, w, samples=samples, depth=depth, conf=conf)
        if not self.regs[s].fused_tail:
            kcall(f"s{s}.softmax_regress", 0, 8 * D * P + 8 * P, "cds_softmax_regress", ptr(logits), ptr(samples), 1, 0, B, D, h, w,
                  ptr(depth), ptr(conf), None)
        return {"depth": depth, "photometric_confidence": conf, "norm_curv": nc}

    def _alloc_outputs(self, B, H, W):
        """The per-stage result maps (depth, photometric_confidence [B,h,w]; norm_curv [B,1,h,w]) are views of ONE buffer, so a
        caller that needs private copies (``CDSMVSNet.forward``) makes one device copy instead of nine."""
        sizes = [B * (H // STAGE_SCALE[s]) * (W // STAGE_SCALE[s]) for s in range(len(self.ndepths))]
        pack = self._out_pack = self.buf.get("out.pack", (3 * sum(sizes),), torch.float32)
        self._out_views, o = [], 0
        for s, n in enumerate(sizes):
            h, w = H // STAGE_SCALE[s], W // STAGE_SCALE[s]
            self._out_views.append((pack[o:o + n].view(B, h, w), pack[o + n:o + 2 * n].view(B, h, w),
                                    pack[o + 2 * n:o + 3 * n].view(B, 1, h, w)))
            o += 3 * n
        return pack

    def outputs_from(self, pack, refined=None):
        """The reference's output dict (models/model.py:199-223) over the maps stored in ``pack`` (a copy of ``out.pack``)."""
        out, o = {}, 0
        for s, (d, c, n) in enumerate(self._out_views):
            k = d.numel()
            st = {"depth": pack[o:o + k].view(d.shape), "photometric_confidence": pack[o + k:o + 2 * k].view(c.shape),
                  "norm_curv": pack[o + 2 * k:o + 3 * k].view(n.shape)}
            o += 3 * k
            out[f"stage{s + 1}"] = st
            out.update(st)
        out["refined_depth"] = refined if refined is not None else out["depth"]
        return out

    # -- whole forward as one CUDA graph ----------------------------------------------------------
    def forward_graph(self, imgs, proj_matrices, depth_values, temperature=0.001):
        """Same as ``forward`` for DEVICE inputs, replayed from a CUDA graph: the ~70 launches of a forward (all on static
        buffers, every argument fixed by the input shapes) are captured once per input signature and re-issued as one graph
        launch; the inputs are copied into the graph's static input tensors first.  Returns the engine's output buffers."""
        key = (tuple(imgs.shape), imgs.dtype, tuple(sorted((k, tuple(v.shape)) for k, v in proj_matrices.items())),
               tuple(depth_values.shape), float(temperature))
        if getattr(self, "_graph_key", None) != key:
            dev, f32 = self.device, torch.float32
            self._graph = self._graph_key = None   # the buffers below may move: the old graph must never be replayed again
            self._g_imgs = torch.empty(imgs.shape, dtype=imgs.dtype if imgs.dtype == torch.uint8 else f32, device=dev)
            self._g_proj = {k: torch.empty(v.shape, dtype=f32, device=dev) for k, v in proj_matrices.items()}
            self._g_dv = torch.empty(depth_values.shape, dtype=f32, device=dev)
            self._copy_inputs(imgs, proj_matrices, depth_values)
            self.forward(self._g_imgs, self._g_proj, self._g_dv, temperature)   # warm-up: allocates every buffer, loads modules
            torch.cuda.current_stream(dev).synchronize()
            before = _lib.LAUNCHES
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                self._g_out = self.forward(self._g_imgs, self._g_proj, self._g_dv, temperature)
            self._graph_launches = _lib.LAUNCHES - before
            self._graph, self._graph_key = graph, key
        self._copy_inputs(imgs, proj_matrices, depth_values)
        self._graph.replay()
        _lib.LAUNCHES += self._graph_launches
        return self._g_out

    def _copy_inputs(self, imgs, proj_matrices, depth_values):
        self._g_imgs.copy_(imgs, non_blocking=True)
        for k, v in proj_matrices.items():
            self._g_proj[k].copy_(v, non_blocking=True)
        self._g_dv.copy_(depth_values, non_blocking=True)

    # -- whole forward --------------------------------------------------------------------------
    def forward(self, imgs, proj_matrices, depth_values, temperature=0.001):
        if imgs.dim() != 5 or imgs.shape[2] != 3:
            raise AssertionError(f"imgs must be [B,N,3,H,W], got {tuple(imgs.shape)}")
        B, N, _, H, W = imgs.shape
        if N < 2:
            raise AssertionError("need at least one source view")
        V = N - 1
        dev = self.device
        u8_work = None
        if imgs.dtype == torch.uint8:
            # 8-bit images as the data layer reads them (datasets/general_eval.py:74: np.float32(img) / 255.): uploaded as bytes
            # (a quarter of the PCIe traffic).  conv00 takes them as they are (k / 256 in pixel-pair operand slots, 256 / 255 in
            # its weights); only the Refinement network (refine=True) needs the fp32 image, divided here -- the IEEE fp32
            # quotient is bit-identical to the host's.
            u8 = imgs.to(device=dev).contiguous()
            self._keep_u8 = u8
            Hw, Ww = (H // 2, W // 2) if self.refiner is not None else (H, W)
            if self.features.u8_ok(Hw, Ww):
                u8_work = u8
            if u8_work is None or self.refiner is not None:
                imgs = self.buf.get("in.imgs_f32", tuple(u8.shape), torch.float32)
                kcall("image_u8_to_f32", 0, 5 * u8.numel(), "cds_image_u8_to_f32", ptr(u8), u8.numel(), ptr(imgs))
        if imgs.dtype != torch.uint8:
            imgs = imgs.to(device=dev, dtype=torch.float32).contiguous()
        full_imgs, Hf, Wf = imgs, H, W
        if self.refiner is not None:
            # refine=True: the cascade works at half resolution on nearest-subsampled images (models/model.py:145-147: the
            # default F.interpolate picks pixel (2i, 2j)); the Refinement network restores the full resolution at the end
            if H % 64 or W % 64:
                raise RuntimeError(f"refine=True needs H and W divisible by 64 (got {H}x{W}); see SURVEY.md 8c fixture 6")
            H, W = H // 2, W // 2
            if u8_work is not None:
                half8 = self.buf.get("rf.imgs_half_u8", (B, N, 3, H, W), torch.uint8)
                half8.copy_(u8_work[..., ::2, ::2])
                u8_work = half8
            else:
                half = self.buf.get("rf.imgs_half", (B, N, 3, H, W), torch.float32)
                half.copy_(imgs[..., ::2, ::2])
                imgs = half
        if H % 32 or W % 32:
            raise RuntimeError(f"H and W must be divisible by 32 (got {H}x{W}); see SURVEY.md 8c fixture 6")
        depth_values = depth_values.to(device=dev, dtype=torch.float32).contiguous()
        coef, epi = self.camera_setup(proj_matrices, B, N)
        self._alloc_outputs(B, H, W)
        n = 2 * V * B
        key = ("imgidx", B, N)
        if getattr(self, "_imgidx_key", None) != key:
            idx = torch.empty(2, V, B, dtype=torch.int32)
            for v in range(V):
                for b in range(B):
                    idx[0, v, b] = b * N
                    idx[1, v, b] = b * N + v + 1
            self._imgidx = idx.reshape(-1).to(dev)
            self._imgidx_key = key
        outputs, stage1 = {}, {}
        overlap = self.overlap and len(self.ndepths) > 1

        def start_stage1(feat1):
            # Stage 1 only needs the quarter-resolution feature: it runs on a side stream while this stream finishes the
            # half / full resolution heads of the feature extractor (disjoint buffers; joined before stage 2).  Its deep
            # regulariser layers are small grids, so the two streams fill each other's gaps.
            main = torch.cuda.current_stream(dev)
            self._side.wait_stream(main)
            with torch.cuda.stream(self._side):
                stage1["out"] = self.stage(0, feat1, coef[0], depth_values, None, B, V, H, W)
            stage1["main"] = main

        feats = self.features.run(self.buf, u8_work if u8_work is not None else imgs, self._imgidx, epi, n, H, W, temperature, pairs=(V, B),
                                  after_stage1=start_stage1 if overlap else None)
        depth = None
        for s in range(len(self.ndepths)):
            if s == 0 and overlap:
                stage1["main"].wait_stream(self._side)
                o = stage1["out"]
            else:
                o = self.stage(s, feats[s], coef[s], depth_values, depth, B, V, H, W)
            depth = o["depth"]
            outputs[f"stage{s + 1}"] = o
            outputs.update(o)
        if self.refiner is not None:
            # models/model.py:210-216: depths in units of the plane interval go through the net, the result is scaled back
            iv = self.buf.get("rf.interval", (3, B), torch.float32)          # rows: interval, dmin / interval, dmax / interval
            torch.sub(depth_values[:, 1], depth_values[:, 0], out=iv[0])
            torch.div(depth_values[:, 0], iv[0], out=iv[1])
            torch.div(depth_values[:, -1], iv[0], out=iv[2])
            d0 = self.buf.get("rf.depth0", (B, H, W), torch.float32)
            torch.div(depth, iv[0].view(B, 1, 1), out=d0)
            ref_img = self.buf.get("rf.ref_img", (B, 3, Hf, Wf), torch.float32)
            ref_img.copy_(full_imgs[:, 0])
            outputs["refined_depth"] = self.refiner.run(self.buf, ref_img, d0, iv[1], iv[2], iv[0], B, Hf, Wf)
        else:
            outputs["refined_depth"] = depth
        return outputs
