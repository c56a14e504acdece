"""Derived weight caches for the CUDA kernels.

The stored parameters keep the reference's names/shapes (so its checkpoints load unchanged); what
the kernels consume -- BatchNorm folded into conv weight/bias, tap-major channels-last weight
layouts, packed gate MLPs -- is derived here from a state dict and must be rebuilt whenever the
parameters change (``load_state_dict`` / ``.to()``).  All folding is done in fp64, stored fp32.

Reference layouts being re-laid-out:
  nn.Conv2d weight [Cout,Cin,k,k], nn.Conv3d [Cout,Cin,3,3,3], nn.ConvTranspose3d [Cin,Cout,3,3,3]
  (models/module.py:102,146 ), BatchNorm eval fold (models/module.py:104,148,194).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import torch

BN_EPS = 1e-5

# FeatureNet layer table (models/module.py:211-234): name -> (Cin, Cout, kernel sizes, state-dict prefix)
DYN_LAYERS = {
    "conv00": (3, 8, (3, 7, 11), "feature.conv00.conv"), "conv01": (8, 8, (3, 5, 7), "feature.conv01.conv"),
    "conv10": (16, 16, (3, 5), "feature.conv10.conv"), "conv11": (16, 16, (3, 5), "feature.conv11.conv"),
    "conv20": (32, 32, (1, 3), "feature.conv20.conv"), "conv21": (32, 32, (1, 3), "feature.conv21.conv"),
    "out1": (32, 32, (1, 3), "feature.out1"), "out2": (16, 16, (1, 3), "feature.out2"),
    "out3": (8, 8, (1, 3), "feature.out3"),
}
# DynamicConv layers that run on csrc/dynconv_kh.cu: the trunk image -> stage-1 feature.  The kernel also covers the stage-2/3
# heads (out2, out3), but its eight epilogue warps per SM are the bound of such light layers: measured 0.245 / 0.505 ms against
# 0.181 / 0.318 ms of csrc/dynconv_tc.cu (several small CTAs per SM).  CDS_KH_LAYERS overrides (diagnostics)
KH_LAYERS = tuple(n for n in __import__("os").environ.get("CDS_KH_LAYERS", "conv00,conv01,conv10,conv11,conv20,conv21,out1").split(",") if n)
COSTREG_CONVS = ("conv0", "conv1", "conv2", "conv3", "conv4", "conv5", "conv6")
COSTREG_DECONVS = ("conv7", "conv9", "conv11")


def _host(sd):
    """The state dict on the HOST: every fold / re-layout below is host arithmetic (fp64), only the finished operand images are
    uploaded -- a model that already sits on the GPU must not turn weight packing into hundreds of tiny device kernels."""
    return {k: v.detach().cpu() for k, v in sd.items()}


def _up(t, dtype, device):
    """Convert and make contiguous on the HOST, then upload (one memcpy, no device kernel)."""
    return t.detach().cpu().to(dtype).contiguous().to(device)


def _bn_fold(sd, prefix):
    g, b = sd[prefix + ".weight"].double(), sd[prefix + ".bias"].double()
    m, v = sd[prefix + ".running_mean"].double(), sd[prefix + ".running_var"].double()
    scale = g / torch.sqrt(v + BN_EPS)
    return scale, b - m * scale


@dataclass
class DynWeights:
    cin: int
    cout: int
    ksizes: tuple
    w_att: torch.Tensor    # [sum k*k, Cin, 4]
    w_conv: torch.Tensor   # [sum k*k, Cin, Cout]
    bias: torch.Tensor | None  # [K, Cout]
    gate: torch.Tensor     # 4K + 4 + 4K floats
    tc: torch.Tensor | None = None   # fp16 tensor-core operand image of csrc/dynconv_tc.cu (tap GEMM)
    kh: torch.Tensor | None = None   # fp16 operand images of csrc/dynconv_kh.cu (kernel rows folded into N; trunk layers)
    kh_u8: torch.Tensor | None = None   # conv00 only: the pixel-pair form for 8-bit images (cds_dynamic_conv_kh_u8)


def pack_dynamic_conv(sd, prefix, cin, cout, ksizes, device) -> DynWeights:
    sd = _host({k: v for k, v in sd.items() if k.startswith(prefix + ".")})
    att, conv, bias = [], [], []
    for i, k in enumerate(ksizes):
        a = sd[f"{prefix}.att_convs.{i}.weight"].double()          # [3,Cin,k,k]
        a = a.permute(2, 3, 1, 0).reshape(k * k, cin, 3)
        att.append(torch.cat((a, torch.zeros(k * k, cin, 1, dtype=torch.float64, device=a.device)), 2))
        w = sd[f"{prefix}.convs.{i}.weight"].double()               # [Cout,Cin,k,k]
        conv.append(w.permute(2, 3, 1, 0).reshape(k * k, cin, cout))
        bk = f"{prefix}.convs.{i}.bias"
        if bk in sd:
            bias.append(sd[bk].double())
    K = len(ksizes)
    scale, shift = _bn_fold(sd, f"{prefix}.att_weights.1")
    w1 = sd[f"{prefix}.att_weights.0.weight"].double().reshape(4, K) * scale.reshape(4, 1)
    w2 = sd[f"{prefix}.att_weights.3.weight"].double().reshape(K, 4)
    gate = torch.cat((w1.reshape(-1), shift.reshape(-1), w2.reshape(-1)))
    f32 = dict(dtype=torch.float32, device=device)
    return DynWeights(cin, cout, tuple(ksizes), torch.cat(att).to(torch.float32).contiguous().to(device), torch.cat(conv).to(torch.float32).contiguous().to(device),
                      torch.stack(bias).to(torch.float32).contiguous().to(device) if bias else None, gate.to(torch.float32).contiguous().to(device))


def pack_dynamic_conv_tc(w: DynWeights) -> torch.Tensor:
    """fp16 B-operand image for csrc/dynconv_tc.cu (Cin 3 is zero-padded to 8).

    Every branch is embedded in the kmax x kmax tap grid.  Branch b owns N columns [b*NPAD, (b+1)*NPAD) with
    NPAD = roundup16(F + 6): F = 2*Cout feature columns (fp16-rounded weights followed by their rounding residuals), then
    (a, b, c) rounded to fp16, then the rounding residuals of (a, b, c).
    K = 16 per MMA = two 8-channel slabs: for Cin <= 8 two consecutive taps ([tap0, zero pad], [tap1, tap2], ...), for
    Cin > 8 two channel chunks of one tap.  First the MMAs of the inner taps (support of the second-largest kernel;
    all branches, N = K*NPAD), then those of the outer ring (largest kernel only, N = NPAD); per MMA
    [k-chunk 2][N/8][8 n][8 k]."""
    cin, cout = w.cin, w.cout
    c8 = max(1, cin // 8)
    att, conv = w.w_att.detach().cpu().double(), w.w_conv.detach().cpu().double()
    K, kmax = len(w.ksizes), max(w.ksizes)
    wlo = True                                    # every layer carries the feature weights' fp16 rounding residual
    fcols = (2 if wlo else 1) * cout               # feature columns per branch; the curvature columns follow
    npad = (fcols + 6 + 15) // 16 * 16
    ntap = kmax * kmax
    full = torch.zeros(ntap, c8 * 8, npad * K, dtype=torch.float64)           # [tap, k (cin), n]
    t0 = 0
    for b, k in enumerate(w.ksizes):
        o = (kmax - k) // 2
        for ky in range(k):
            for kx in range(k):
                t = (ky + o) * kmax + (kx + o)
                src = t0 + ky * k + kx
                if wlo:
                    chi = conv[src].to(torch.float16).to(torch.float64)
                    full[t, :cin, npad * b:npad * b + cout] = chi
                    full[t, :cin, npad * b + cout:npad * b + 2 * cout] = conv[src] - chi
                else:
                    full[t, :cin, npad * b:npad * b + cout] = conv[src]
                a = att[src][:, :3]
                hi = a.to(torch.float16).to(torch.float64)
                full[t, :cin, npad * b + fcols:npad * b + fcols + 3] = hi
                full[t, :cin, npad * b + fcols + 3:npad * b + fcols + 6] = a - hi
        t0 += k * k
    if cin == 3:   # image layer: channels 3..5 of the operand carry the image's fp16 rounding residual (cds_image_to_nhwc8)
        full[:, 3:6, :] = full[:, 0:3, :]
    kin = max(k for k in w.ksizes if k < kmax)
    lo, hi_ = (kmax - kin) // 2, (kmax - kin) // 2 + kin
    inside = lambda t: lo <= t // kmax < hi_ and lo <= t % kmax < hi_
    inner = [t for t in range(ntap) if inside(t)]
    ring = [t for t in range(ntap) if not inside(t)]

    def slabs_of(taps, pad_first):
        if c8 == 1:
            return ([(taps[0], 0), None] + [(t, 0) for t in taps[1:]]) if pad_first else [(t, 0) for t in taps]
        return [(t, c) for t in taps for c in range(c8)]

    def image(slabs, col0, ncol):
        assert len(slabs) % 2 == 0
        img = torch.zeros(len(slabs) // 2, 2, ncol // 8, 8, 8, dtype=torch.float64)
        for s, sl in enumerate(slabs):
            if sl is not None:
                t, c = sl
                img[s // 2, s % 2] = full[t, c * 8:(c + 1) * 8, col0:col0 + ncol].t().reshape(ncol // 8, 8, 8)
        return img.reshape(-1)

    for t in ring:
        assert full[t, :, :npad * (K - 1)].abs().max() == 0     # only the largest kernel reaches the ring
    out = torch.cat((image(slabs_of(inner, True), 0, npad * K), image(slabs_of(ring, False), npad * (K - 1), npad)))
    return out.to(torch.float16).contiguous().to(w.w_conv.device)


def kh_layout(cout: int, k: int) -> tuple[int, int]:
    """(columns of one kernel-row group, columns of one image of a k x k branch) of csrc/dynconv_kh.cu -- asked of the library,
    which owns the layout."""
    from . import _lib
    lib = _lib.LIB.load()
    return int(lib.cds_dynamic_conv_kh_group_cols(cout)), int(lib.cds_dynamic_conv_kh_image_cols(cout, k))


def pack_dynamic_conv_kh(w: DynWeights, px2: bool = False) -> torch.Tensor:
    """fp16 B-operand images for csrc/dynconv_kh.cu (kernel rows folded into N; Cin 3 is padded to 8 with the image's
    residual channels).

    Per branch b (kernel k, in the order of ``ksizes``) two images -- the fp16-rounded weights, then their fp16 rounding
    residuals -- each ``nj`` steps of [k-chunk 2][NCOLS/8][8 n][8 k].  Column n = g*NPAD + c with column group g <-> kernel
    row dy = k-1-g (so that group g feeds output row y = R - h + g of input row R), c < Cout the feature weights, c in
    [Cout, Cout+3) the curvature weights (a, b, c), the rest -- and the columns from k*NPAD to NCOLS -- zero; NPAD and NCOLS
    come from ``kh_layout``.  K: Cin <= 8: step j = horizontal taps (2j, 2j+1) x 8 channels (zero past the kernel); Cin > 8:
    tap (2j)//C8, channel chunks (2j)%C8 and +1.

    ``px2`` (the image layer on 8-bit images, cds_dynamic_conv_kh_u8): an operand slot holds (RGB of a pixel, RGB of its right
    neighbour, 0, 0) as k / 256 for the byte k (exact in fp16), so step j covers the FOUR taps 4j .. 4j+3 -- k-chunk q = taps
    (4j+2q, 4j+2q+1), K rows 0-2 / 3-5 their RGB weights -- and the factor 256 / 255 that turns k / 256 into the loader's
    k / 255 is folded into the weights before the hi / lo split."""
    if px2 and w.cin != 3:
        raise ValueError("pixel-pair operand slots exist for the 3-channel image layer only")
    cin, cout = w.cin, w.cout
    c8 = max(1, cin // 8)
    att, conv = w.w_att.detach().cpu().double(), w.w_conv.detach().cpu().double()
    parts, t0 = [], 0
    for k in w.ksizes:
        npad, ncols = kh_layout(cout, k)
        full = torch.zeros(k, k, c8 * 8, npad, dtype=torch.float64)           # [dy, dx, cin, n]
        for dy in range(k):
            for dx in range(k):
                src = t0 + dy * k + dx
                full[dy, dx, :cin, :cout] = conv[src]
                full[dy, dx, :cin, cout:cout + 3] = att[src][:, :3]
        if px2:
            full = full * (256.0 / 255.0)   # the slots hold k / 256 (exact in fp16, weights stay in fp16's normal range)
        elif cin == 3:   # channels 3..5 of the operand carry the image's fp16 rounding residual (cds_image_to_nhwc8)
            full[:, :, 3:6, :] = full[:, :, 0:3, :]
        t0 += k * k
        hi = full.to(torch.float16).to(torch.float64)
        nj = (k + 3) // 4 if px2 else ((k + 1) // 2 if c8 == 1 else k * c8 // 2)
        for img_w in (hi, full - hi):
            img = torch.zeros(nj, 2, ncols, 8, dtype=torch.float64)             # [step, k-chunk, column n, 8 k]
            for j in range(nj):
                for q in range(2):
                    if px2:
                        for half, dx in enumerate((4 * j + 2 * q, 4 * j + 2 * q + 1)):
                            if dx < k:
                                for g in range(k):
                                    img[j, q, g * npad:(g + 1) * npad, 3 * half:3 * half + 3] = img_w[k - 1 - g, dx, 0:3, :].t()
                        continue
                    if c8 == 1:
                        dx, ch = 2 * j + q, 0
                    else:
                        dx, ch = (2 * j) // c8, (2 * j) % c8 + q
                    if dx >= k:
                        continue
                    for g in range(k):
                        img[j, q, g * npad:(g + 1) * npad] = img_w[k - 1 - g, dx, ch * 8:(ch + 1) * 8, :].t()
            parts.append(img.reshape(-1))
    return torch.cat(parts).to(torch.float16).contiguous().to(w.w_conv.device)


def pack_conv2d(sd, key, device) -> torch.Tensor:
    w = sd[key].detach().cpu().double()                                            # [Cout,Cin,k,k]
    co, ci, k, _ = w.shape
    return w.permute(2, 3, 1, 0).reshape(k * k, ci, co).to(torch.float32).contiguous().to(device)


def pack_visnet(sd, prefix, device) -> torch.Tensor:
    sd = _host({k: v for k, v in sd.items() if k.startswith(prefix + ".")})
    parts = []
    for j in range(3):
        scale, shift = _bn_fold(sd, f"{prefix}.{j}.bn")
        w = sd[f"{prefix}.{j}.conv.weight"].double() * scale.reshape(-1, 1, 1, 1)   # [16,Cin,3,3]
        parts += [w.permute(2, 3, 1, 0).reshape(-1), shift.reshape(-1)]
    parts += [sd[f"{prefix}.3.weight"].double().reshape(-1), sd[f"{prefix}.3.bias"].double().reshape(-1)]
    return torch.cat(parts).to(torch.float32).contiguous().to(device)


def pack_visnet_tc(sd, prefix, device):
    """Operands for csrc/visnet_tc.cu: (fp16 image [L1 2 steps | L2 3 | L3 3] x [k-chunk 2][6][8 n][8 k], fp32 b1 b2 b3 w4 b4).

    The kernel rows are folded into N: column n = g*16 + cout, group g <-> kernel row 2-g (group g of the MMA on input row i
    feeds output row i-2+g).  Layer 1 sees (entropy_hi, curv_hi, entropy_lo, curv_lo, 0, 0, 0, 0) per pixel, so its weights are
    duplicated for the "lo" channels; its steps pair neighbouring pixels: step 0 = horizontal taps (0, 1), step 1 = (1, 2)
    with zeros for tap 1.  Layers 2/3: step = horizontal tap, the two k-chunks are the two 8-channel halves of the input."""
    ws, bs = [], []
    for j in range(3):
        scale, shift = _bn_fold(sd, f"{prefix}.{j}.bn")
        ws.append((sd[f"{prefix}.{j}.conv.weight"].double() * scale.reshape(-1, 1, 1, 1)).cpu())   # [16, Cin, 3, 3]
        bs.append(shift.cpu())
    parts = []
    img = torch.zeros(2, 2, 48, 8, dtype=torch.float64)                       # [step, k-chunk, column, k]
    for step, taps in enumerate(((0, 1), (None, 2))):
        for q, kw in enumerate(taps):
            if kw is None:
                continue
            for g in range(3):
                wt = ws[0][:, :, 2 - g, kw]                                    # [16 cout, 2]
                img[step, q, g * 16:(g + 1) * 16, 0] = wt[:, 0]
                img[step, q, g * 16:(g + 1) * 16, 1] = wt[:, 1]
                img[step, q, g * 16:(g + 1) * 16, 2] = wt[:, 0]
                img[step, q, g * 16:(g + 1) * 16, 3] = wt[:, 1]
    parts.append(img.reshape(-1))
    for j in (1, 2):
        img = torch.zeros(3, 2, 48, 8, dtype=torch.float64)
        for kw in range(3):
            for q in range(2):
                for g in range(3):
                    img[kw, q, g * 16:(g + 1) * 16] = ws[j][:, q * 8:(q + 1) * 8, 2 - g, kw]   # [16 cout, 8 cin]
        parts.append(img.reshape(-1))
    wgt = torch.cat(parts).to(torch.float16).contiguous().to(device)
    fp = torch.cat((bs[0], bs[1], bs[2], sd[f"{prefix}.3.weight"].double().reshape(-1).cpu(),
                    sd[f"{prefix}.3.bias"].double().reshape(-1).cpu())).to(torch.float32).contiguous().to(device)
    return wgt, fp


@dataclass
class Conv3dWeights:
    cin: int
    cout: int
    w: torch.Tensor      # [27, Cin, Cout] fp32, BN scale folded
    bias: torch.Tensor   # [Cout]
    extra: dict = field(default_factory=dict)   # tensor-core re-layouts live here


def pack_conv3d(sd, prefix, transposed, device) -> Conv3dWeights:
    sd = _host({k: v for k, v in sd.items() if k.startswith(prefix + ".")})
    scale, shift = _bn_fold(sd, prefix + ".bn")
    w = sd[prefix + ".conv.weight"].double()
    if transposed:   # [Cin,Cout,3,3,3]
        ci, co = w.shape[:2]
        w = (w * scale.reshape(1, -1, 1, 1, 1)).permute(2, 3, 4, 0, 1)
    else:            # [Cout,Cin,3,3,3]
        co, ci = w.shape[:2]
        w = (w * scale.reshape(-1, 1, 1, 1, 1)).permute(2, 3, 4, 1, 0)
    f32 = dict(dtype=torch.float32, device=device)
    return Conv3dWeights(ci, co, w.reshape(27, ci, co).to(torch.float32).contiguous().to(device), shift.to(torch.float32).contiguous().to(device))


def pack_conv3d_tc(l: Conv3dWeights) -> torch.Tensor:
    """fp16 B-operand image for csrc/conv3d_tc.cu: [mma][k-chunk 2][Npad/8][8 n][8 k] (K-major, no swizzle).

    K order = 8-channel slabs.  Cin == 8: [tap0, zero pad], [tap1, tap2], ..., [tap25, tap26];
    Cin > 8: slabs tap-major / channel-chunk minor, consecutive pairs."""
    ci, co = l.cin, l.cout
    c8 = ci // 8
    npad = max(16, co)
    w = l.w.detach().cpu().to(torch.float64)                       # [27, Cin, Cout]
    if c8 == 1:
        slabs = [(0, 0), None] + [(t, 0) for t in range(1, 27)]
    else:
        slabs = [(t, c) for t in range(27) for c in range(c8)]
    assert len(slabs) % 2 == 0
    img = torch.zeros(len(slabs) // 2, 2, npad // 8, 8, 8, dtype=torch.float64)
    for s, sl in enumerate(slabs):
        if sl is None:
            continue
        t, c = sl
        blk = w[t, c * 8:(c + 1) * 8, :]                            # [8 k, Cout]
        full = torch.zeros(8, npad, dtype=torch.float64)
        full[:, :co] = blk
        if co in (1, 8):   # N is padded to 16 anyway: the spare columns carry the fp16 rounding residual of the weights
            hi = blk.to(torch.float16).to(torch.float64)
            full[:, :co] = hi
            full[:, co:2 * co] = blk - hi
        img[s // 2, s % 2] = full.t().reshape(npad // 8, 8, 8)      # [n-group, n row, k]
    return img.to(torch.float16).contiguous().to(l.w.device)


def pack_deconv3d_tc(l: Conv3dWeights) -> torch.Tensor:
    """fp16 B-operand image for the transposed-conv kernel in csrc/conv3d_tc.cu.

    MMA j = (neighbour offset s = (sd,sh,sw) in {0,1}^3, channel-chunk pair q); N = 8*Cout, column n = parity*Cout + co
    with parity = (pd,ph,pw).  Per axis: parity 0 uses tap k=1 from offset 0 only; parity 1 uses k=2 from offset 0 and
    k=0 from offset 1 (out[2i-1+k] += in[i] w[k]).  Layout [mma][k-chunk 2][N/8][8 n][8 k]."""
    ci, co = l.cin, l.cout
    c8, N = ci // 8, 8 * l.cout
    w = l.w.detach().cpu().to(torch.float64)                       # [27, Cin, Cout] folded

    def tap(par, off):
        if par == 0:
            return 1 if off == 0 else None
        return 2 if off == 0 else 0

    img = torch.zeros(8 * c8 // 2, 2, N // 8, 8, 8, dtype=torch.float64)
    for s in range(8):
        sd, sh, sw = s >> 2, (s >> 1) & 1, s & 1
        blockw = torch.zeros(ci, N, dtype=torch.float64)            # [k (cin), n]
        for par in range(8):
            pd, ph, pw = par >> 2, (par >> 1) & 1, par & 1
            kd, kh, kw = tap(pd, sd), tap(ph, sh), tap(pw, sw)
            if kd is None or kh is None or kw is None:
                continue
            blockw[:, par * co:(par + 1) * co] = w[(kd * 3 + kh) * 3 + kw]
        for q in range(c8 // 2):
            j = s * (c8 // 2) + q
            for kc in range(2):
                c = 2 * q + kc
                img[j, kc] = blockw[c * 8:(c + 1) * 8].t().reshape(N // 8, 8, 8)
    return img.to(torch.float16).contiguous().to(l.w.device)


def _hi_lo_columns(blk: torch.Tensor) -> torch.Tensor:
    """[8 k, Cout] fp64 -> [8 k, 2*Cout]: fp16-rounded weights, then their rounding residual (summed in the epilogue)."""
    hi = blk.to(torch.float16).to(torch.float64)
    return torch.cat((hi, blk - hi), dim=1)


def pack_conv3d_gtc(l: Conv3dWeights) -> torch.Tensor:
    """fp16 B-operand image for csrc/conv3d_gtc.cu (Conv3d stride 1|2): [mma][k-chunk 2][N/8][8 n][8 k], N = 2*Cout
    (weights + residual columns).  Slab order as pack_conv3d_tc: Cin == 8: [tap0, zero, tap1, ..., tap26]; Cin > 8:
    tap-major / channel-chunk minor; an MMA takes two consecutive slabs."""
    ci, co = l.cin, l.cout
    c8, N = ci // 8, 2 * l.cout
    w = l.w.detach().cpu().to(torch.float64)                       # [27, Cin, Cout]
    slabs = [(0, 0), None] + [(t, 0) for t in range(1, 27)] if c8 == 1 else [(t, c) for t in range(27) for c in range(c8)]
    assert len(slabs) % 2 == 0
    img = torch.zeros(len(slabs) // 2, 2, N // 8, 8, 8, dtype=torch.float64)
    for s, sl in enumerate(slabs):
        if sl is None:
            continue
        t, c = sl
        img[s // 2, s % 2] = _hi_lo_columns(w[t, c * 8:(c + 1) * 8, :]).t().reshape(N // 8, 8, 8)
    return img.to(torch.float16).contiguous().to(l.w.device)


def pack_conv3d_roll(l: Conv3dWeights) -> torch.Tensor:
    """fp16 B-operand image for csrc/conv3d_roll.cu (Conv3d k3 s1, Cout 8 or 1, kw folded into N):
    [kd][mma][k-chunk 2][N/8][8 n][8 k]; column n = kw*2*Cout + j holds the fp16-rounded weight (j < Cout) or its rounding
    residual (j >= Cout).  K slabs per kd: Cin == 8: [kh0, kh1], [kh2, zero]; Cin > 8: kh-major, channel-chunk pairs."""
    ci, co = l.cin, l.cout
    c8, cw = ci // 8, 2 * l.cout
    npad = (3 * cw + 15) // 16 * 16
    w = l.w.detach().cpu().to(torch.float64)                       # [27, Cin, Cout]
    slabs = [(0, 0), (1, 0), (2, 0), None] if c8 == 1 else [(kh, c) for kh in range(3) for c in range(c8)]
    mma_kd = len(slabs) // 2
    img = torch.zeros(3 * mma_kd, 2, npad // 8, 8, 8, dtype=torch.float64)
    for kd in range(3):
        for s, sl in enumerate(slabs):
            if sl is None:
                continue
            kh, c = sl
            full = torch.zeros(8, npad, dtype=torch.float64)
            for kw in range(3):
                full[:, kw * cw:(kw + 1) * cw] = _hi_lo_columns(w[(kd * 3 + kh) * 3 + kw, c * 8:(c + 1) * 8, :])
            img[kd * mma_kd + s // 2, s % 2] = full.t().reshape(npad // 8, 8, 8)
    return img.to(torch.float16).contiguous().to(l.w.device)


def pack_deconv3d_gtc(l: Conv3dWeights) -> torch.Tensor:
    """fp16 B-operand image for the transposed conv in csrc/conv3d_gtc.cu: the 8 output-parity classes (pd,ph,pw) one after
    the other, each a dense conv over its input neighbours (sd,sh,sw) in {0..pd}x{0..ph}x{0..pw}; per axis parity 0 takes
    tap k=1 (neighbour 0), parity 1 takes k=2 from neighbour 0 and k=0 from neighbour 1 (out[2i-1+k] += in[i] w[k]).
    Layout per class [tap][chunk pair][k-chunk 2][N/8][8 n][8 k], N = 2*Cout (weights + residual columns)."""
    ci, co = l.cin, l.cout
    c8, N = ci // 8, 2 * l.cout
    w = l.w.detach().cpu().to(torch.float64)                       # [27, Cin, Cout] folded

    def tap(par, off):
        return 1 if par == 0 else (2 if off == 0 else 0)

    mmas = []
    for cls in range(8):
        pd, ph, pw = cls >> 2, (cls >> 1) & 1, cls & 1
        for sd in range(pd + 1):
            for sh in range(ph + 1):
                for sw in range(pw + 1):
                    wt = w[(tap(pd, sd) * 3 + tap(ph, sh)) * 3 + tap(pw, sw)]           # [Cin, Cout]
                    for q in range(c8 // 2):
                        m = torch.zeros(2, N // 8, 8, 8, dtype=torch.float64)
                        for kc in range(2):
                            c = 2 * q + kc
                            m[kc] = _hi_lo_columns(wt[c * 8:(c + 1) * 8]).t().reshape(N // 8, 8, 8)
                        mmas.append(m)
    return torch.stack(mmas).to(torch.float16).contiguous().to(l.w.device)


@dataclass
class CostRegWeights:
    layers: dict          # name -> Conv3dWeights
    prob: torch.Tensor    # [27, 8]
    prob_tc: torch.Tensor | None = None
    prob_roll: torch.Tensor | None = None


def pack_costreg(sd, prefix, device) -> CostRegWeights:
    layers = {n: pack_conv3d(sd, f"{prefix}.{n}", False, device) for n in COSTREG_CONVS}
    layers.update({n: pack_conv3d(sd, f"{prefix}.{n}", True, device) for n in COSTREG_DECONVS})
    for n in ("conv0", "conv2", "conv4"):
        layers[n].extra["tc"] = pack_conv3d_tc(layers[n])
    layers["conv0"].extra["roll"] = pack_conv3d_roll(layers["conv0"])
    for n in ("conv9", "conv11"):
        layers[n].extra["tc"] = pack_deconv3d_tc(layers[n])
    for n in ("conv1", "conv2", "conv3", "conv4", "conv5", "conv6"):   # gather-form kernel (stride 2, small deep layers)
        layers[n].extra["gtc"] = pack_conv3d_gtc(layers[n])
    for n in ("conv7", "conv9"):
        layers[n].extra["gtc"] = pack_deconv3d_gtc(layers[n])
    p = sd[prefix + ".prob.weight"].detach().cpu().double()                        # [1,8,3,3,3]
    prob = p.permute(2, 3, 4, 1, 0).reshape(27, p.shape[1]).to(torch.float32).contiguous().to(device)
    cw = CostRegWeights(layers, prob)
    if p.shape[1] == 8:   # tensor-core image of the prob head (8 -> 1)
        zero = torch.zeros(1).to(device)
        cw.prob_tc = pack_conv3d_tc(Conv3dWeights(8, 1, prob.reshape(27, 8, 1), zero))
        cw.prob_roll = pack_conv3d_roll(Conv3dWeights(8, 1, prob.reshape(27, 8, 1), zero))
    return cw


def pack_conv2d_gtc(w: torch.Tensor) -> torch.Tensor:
    """fp16 B-operand image for csrc/conv2d_gtc.cu from tap-major weights [taps, Cin, Cout] (3x3 s2: 9 taps; 1x1: 1 tap):
    [mma][k-chunk 2][N/8][8 n][8 k], N = 2*Cout (weights + fp16 rounding residual columns); K slabs tap-major / 8-channel
    chunk minor, zero-padded to an even count; an MMA takes two consecutive slabs."""
    taps, ci, co = w.shape
    c8, N = ci // 8, 2 * co
    w = w.detach().cpu().to(torch.float64)
    slabs = [(t, c) for t in range(taps) for c in range(c8)]
    if len(slabs) % 2:
        slabs.append(None)
    img = torch.zeros(len(slabs) // 2, 2, N // 8, 8, 8, dtype=torch.float64)
    for s, sl in enumerate(slabs):
        if sl is None:
            continue
        t, c = sl
        img[s // 2, s % 2] = _hi_lo_columns(w[t, c * 8:(c + 1) * 8, :]).t().reshape(N // 8, 8, 8)
    return img.to(dtype=torch.float16).contiguous()


def pack_conv2d_s2rows(w: torch.Tensor) -> torch.Tensor:
    """fp16 B-operand images for csrc/conv2d_s2rows.cu from tap-major weights [9, Cin, Cout] of a 3x3 stride-2 conv:
    [kernel row dy][image][k-chunk 2][N/8][8 n][8 k], N = 2*Cout (weights | fp16 rounding residuals).  Cin 8: image 0 = taps
    (dx 0, dx 2) -- the odd-phase pixels left and right --, image 1 = (dx 1, dx 1): the even pixel's value and residual slabs
    share one weight block; Cin 16: image dx = the tap's channel chunks (0, 1)."""
    taps, ci, co = w.shape
    assert taps == 9 and ci in (8, 16)
    N = 2 * co
    w = w.detach().cpu().to(torch.float64)
    blk = lambda t, c: _hi_lo_columns(w[t, c * 8:(c + 1) * 8, :]).t().reshape(N // 8, 8, 8)
    nimg = 2 if ci == 8 else 3
    img = torch.zeros(3, nimg, 2, N // 8, 8, 8, dtype=torch.float64)
    for dy in range(3):
        if ci == 8:
            img[dy, 0, 0], img[dy, 0, 1] = blk(dy * 3 + 0, 0), blk(dy * 3 + 2, 0)
            img[dy, 1, 0], img[dy, 1, 1] = blk(dy * 3 + 1, 0), blk(dy * 3 + 1, 0)
        else:
            for dx in range(3):
                img[dy, dx, 0], img[dy, dx, 1] = blk(dy * 3 + dx, 0), blk(dy * 3 + dx, 1)
    return img.to(dtype=torch.float16).contiguous()


@dataclass
class FeatureWeights:
    dyn: dict             # name -> DynWeights
    downsample1: torch.Tensor
    downsample2: torch.Tensor
    inner1: torch.Tensor  # [48,16]
    inner2: torch.Tensor  # [24,8]
    tc: dict = field(default_factory=dict)   # name -> fp16 tensor-core operand image (csrc/conv2d_gtc.cu)
    rows: dict = field(default_factory=dict)   # stride-2 layers: operand images of csrc/conv2d_s2rows.cu


def pack_feature(sd, device) -> FeatureWeights:
    dyn = {n: pack_dynamic_conv(sd, pre, ci, co, ks, device) for n, (ci, co, ks, pre) in DYN_LAYERS.items()}
    for n in dyn:
        dyn[n].tc = pack_dynamic_conv_tc(dyn[n])
        if n in KH_LAYERS:
            dyn[n].kh = pack_dynamic_conv_kh(dyn[n])
            if dyn[n].cin == 3:
                dyn[n].kh_u8 = pack_dynamic_conv_kh(dyn[n], px2=True)
    fw = FeatureWeights(dyn, pack_conv2d(sd, "feature.downsample1.conv.weight", device),
                        pack_conv2d(sd, "feature.downsample2.conv.weight", device),
                        pack_conv2d(sd, "feature.inner1.conv.weight", device).reshape(48, 16),
                        pack_conv2d(sd, "feature.inner2.conv.weight", device).reshape(24, 8))
    fw.rows = {"downsample1": pack_conv2d_s2rows(fw.downsample1).to(device), "downsample2": pack_conv2d_s2rows(fw.downsample2).to(device)}
    fw.tc = {"downsample1": pack_conv2d_gtc(fw.downsample1).to(device), "downsample2": pack_conv2d_gtc(fw.downsample2).to(device),
             "inner1": pack_conv2d_gtc(fw.inner1.reshape(1, 48, 16)).to(device),
             "inner2": pack_conv2d_gtc(fw.inner2.reshape(1, 24, 8)).to(device)}
    return fw


@dataclass
class ModelWeights:
    feature: FeatureWeights
    vis: list
    costreg: list
    vis_tc: list = field(default_factory=list)   # per stage (fp16 operand image, fp32 biases) for csrc/visnet_tc.cu


def pack_refinement(sd, prefix, device) -> dict:
    """Folded fp32 weights of the Refinement network (models/module.py:318-335) for csrc/refine.cu: per ConvBnReLU
    (w [Cout][Cin][3][3] * bn scale, bias = bn shift); the transposed conv keeps torch's [Cin][Cout][3][3] with the
    following BatchNorm folded over Cout; res.weight as is."""
    sd = _host({k: v for k, v in sd.items() if k.startswith(prefix + ".")})
    f32 = dict(dtype=torch.float32, device=device)
    out = {}
    for n in ("conv0", "conv1", "conv2", "conv3"):
        scale, shift = _bn_fold(sd, f"{prefix}.{n}.bn")
        w = sd[f"{prefix}.{n}.conv.weight"].double() * scale.reshape(-1, 1, 1, 1)
        out[n] = (w.to(torch.float32).contiguous().to(device), shift.to(torch.float32).contiguous().to(device))
    scale, shift = _bn_fold(sd, f"{prefix}.bn")
    w = sd[f"{prefix}.deconv.weight"].double() * scale.reshape(1, -1, 1, 1)                    # [Cin, Cout, 3, 3]
    out["deconv"] = (w.to(torch.float32).contiguous().to(device), shift.to(torch.float32).contiguous().to(device))
    out["res"] = sd[f"{prefix}.res.weight"].double().reshape(8, 9).to(torch.float32).contiguous().to(device)      # [1, 8, 3, 3]
    return out


def pack_model(sd, n_stages: int, device, share_cr: bool = False) -> ModelWeights:
    sd = _host(sd)
    vis = [pack_visnet(sd, f"stage_net.vis.{s}", device) for s in range(n_stages)]
    if share_cr:
        cr = [pack_costreg(sd, "cost_regularization", device)] * n_stages
    else:
        cr = [pack_costreg(sd, f"cost_regularization.{s}", device) for s in range(n_stages)]
    vis_tc = [pack_visnet_tc(sd, f"stage_net.vis.{s}", device) for s in range(n_stages)]
    return ModelWeights(pack_feature(sd, device), vis, cr, vis_tc)
