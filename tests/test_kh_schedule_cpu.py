"""CPU model of the MMA schedule of csrc/dynconv_kh.cu (kernel rows folded into N): the packed operand images of
``weights.pack_dynamic_conv_kh`` are pushed through exactly the index arithmetic the kernel's issuers use -- row streaming,
column-group ranges clipped to the tile, the step's A-slab offsets, accumulation into zeroed slots -- and the accumulator
tile must equal the branch convolutions / curvature convolutions of the oracle.  Host logic only: no kernel runs here."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from cds_mvsnet_b200 import weights as W

TX = 128


def emulate(layer, sd, x, TY):
    """x [Cin_pad, H, W] fp64 (the staged operand) -> per branch [H, W, NPAD] accumulators, via the kernel's schedule."""
    cin, cout, ks, pre = W.DYN_LAYERS[layer]
    w = W.pack_dynamic_conv(sd, pre, cin, cout, ks, "cpu")
    img = W.pack_dynamic_conv_kh(w).double().numpy()
    c8 = max(1, cin // 8)
    npad = W.kh_layout(cout, ks[0])[0]
    tight = npad % 16 != 0
    slots = TY + 1 if tight else TY                    # tight column groups: one spare slot per branch behind the tile
    assert len(ks) * slots * npad <= 512, "accumulator tile exceeds TMEM"
    hmax = (max(ks) - 1) // 2
    txo = TX - 2 * hmax
    _, H, Wd = x.shape
    # operand images per branch: [t (hi, lo)][j][q][n][kk]
    imgs, o = [], 0
    for k in ks:
        nj = (k + 1) // 2 if c8 == 1 else k * c8 // 2
        ncols = W.kh_layout(cout, k)[1]
        sz = nj * 2 * ncols * 8
        per = []
        for _t in range(2):
            per.append(img[o:o + sz].reshape(nj, 2, ncols, 8))
            o += sz
        imgs.append(per)
    assert o == img.size
    out = [np.full((H, Wd, npad), np.nan) for _ in ks]
    xt, yt = -(-Wd // txo), -(-H // TY)
    # TMEM, persistent over the CTA's tiles: per branch [slot columns][128 lanes]; starts zeroed
    acc = [np.zeros((slots * npad, TX)) for _ in ks]
    for ty in range(yt):
        for tx in range(xt):
            y0, x0 = ty * TY, max(0, min(tx * txo, Wd - txo))
            for b in range(len(ks)):   # every slot of the tile was handed over zeroed by the epilogue
                assert not acc[b][:TY * npad].any()
            for R in range(max(y0 - hmax, 0), min(y0 + TY - 1 + hmax, H - 1) + 1):
                # the staged row segment: pixels x0-hmax .. +128 (+ spill), zero outside the image, [px][c8][8]
                row = np.zeros((TX + 2 * hmax + 2, c8 * 8))
                for px in range(TX + 2 * hmax + 2):
                    gx = x0 - hmax + px
                    if 0 <= gx < Wd and px < TX:
                        row[px] = x[:, R, gx]
                    elif px >= TX:
                        row[px] = 7.0      # spill past the segment: finite garbage that only reaches discarded rows / zero weights
                for b, k in enumerate(ks):
                    hb = (k - 1) // 2
                    ylo, yhi = max(y0, R - hb), min(min(y0 + TY - 1, H - 1), R + hb)
                    if ylo > yhi:
                        continue
                    g0 = (ylo - (R - hb)) * npad                       # first weight column
                    n = -(-((yhi - ylo + 1) * npad) // 16) * 16        # an M = 128 MMA needs N % 16 == 0
                    d0 = (ylo - y0) * npad                             # first accumulator column (of the branch's region)
                    assert d0 + n <= slots * npad, "an MMA leaves its branch's accumulator region"
                    nj = (k + 1) // 2 if c8 == 1 else k * c8 // 2
                    for prod in range(2):          # (A, W_hi), (A, W_lo); the A_lo product reuses W_hi with another plane
                        for j in range(nj):
                            A = np.zeros((TX, 16))
                            for q in range(2):
                                if c8 == 1:
                                    off, ch = hmax - hb + 2 * j + q, 0
                                else:
                                    off, ch = hmax - hb + (2 * j) // c8, (2 * j) % c8 + q
                                A[:, q * 8:(q + 1) * 8] = row[off:off + TX, ch * 8:(ch + 1) * 8]
                            Bm = imgs[b][prod][j]                                           # [q][n][kk]
                            Bfull = np.concatenate((Bm[0], Bm[1]), axis=1)                  # [n, 16]
                            assert g0 + n <= Bfull.shape[0], "an MMA reads past its weight image"
                            acc[b][d0:d0 + n] += Bfull[g0:g0 + n] @ A.T                     # every MMA accumulates
            for b in range(len(ks)):
                for s in range(TY):
                    y = y0 + s
                    if y < H:
                        for r in range(txo):
                            gx = x0 + r
                            if gx < Wd and gx >= tx * txo:
                                out[b][y, gx] = acc[b][s * npad:(s + 1) * npad, r]
                    # the epilogue hands every slot back zeroed (an unused one may have caught rounding columns)
                    acc[b][s * npad:(s + 1) * npad] = 0
    return out, (cin, cout, ks, pre)


def _ty(layer):
    cout = W.DYN_LAYERS[layer][1]
    tight = W.kh_layout(cout, 3)[0] % 16 != 0
    return {8: 13 if tight else 10, 16: 11 if tight else 8, 32: 6 if tight else 5}[cout]


@pytest.mark.parametrize("layer,hw", [("conv01", (31, 140)), ("conv00", (13, 131)), ("conv00", (29, 20)), ("conv10", (25, 30)),
                                      ("conv20", (14, 130))])
def test_kh_schedule_reproduces_branch_convolutions(pretrained_sd, layer, hw):
    TY = _ty(layer)
    torch.manual_seed(0)
    cin, cout, ks, pre = W.DYN_LAYERS[layer]
    H, Wd = hw
    x = torch.randn(max(8, cin), H, Wd, dtype=torch.float64)
    if cin == 3:   # the image layer: channels 3..5 carry the residual of 0..2, 6..7 are zero
        x[3:6] = x[0:3] * 1e-3
        x[6:] = 0
    out, _ = emulate(layer, pretrained_sd, x.numpy(), TY)
    xin = (x[0:3] + x[3:6]) if cin == 3 else x[:cin]
    for b, k in enumerate(ks):
        ref_f = F.conv2d(xin[None], pretrained_sd[f"{pre}.convs.{b}.weight"].double(), padding=(k - 1) // 2)[0]
        ref_a = F.conv2d(xin[None], pretrained_sd[f"{pre}.att_convs.{b}.weight"].double(), padding=(k - 1) // 2)[0]
        got = torch.from_numpy(out[b])
        assert not torch.isnan(got[..., :cout + 3]).any(), "a pixel was never written"
        scale = ref_f.abs().max()
        assert (got[..., :cout].permute(2, 0, 1) - ref_f).abs().max() < 1e-5 * scale + 1e-7
        assert (got[..., cout:cout + 3].permute(2, 0, 1) - ref_a).abs().max() < 1e-5 * ref_a.abs().max() + 1e-7
        assert got[..., cout + 3:].abs().max() == 0


def test_kh_schedule_pixel_pair_slots_for_8bit_images(pretrained_sd):
    """The image layer on 8-bit images (cds_dynamic_conv_kh_u8): operand slots hold (RGB of a pixel, RGB of its right neighbour,
    0, 0) as byte / 256 on rows padded by the library's pad, one K = 16 step covers the four taps 4j .. 4j+3 (k-chunk q at slot
    offset 4j + 2q), and the packed images carry 256 / 255: pushed through the kernel's index arithmetic, the accumulators must
    equal the convolutions of float32(byte) / 255."""
    from cds_mvsnet_b200 import _lib
    layer = "conv00"
    cin, cout, ks, pre = W.DYN_LAYERS[layer]
    w = W.pack_dynamic_conv(pretrained_sd, pre, cin, cout, ks, "cpu")
    img = W.pack_dynamic_conv_kh(w, px2=True).double().numpy()
    pad = _lib.LIB.load().cds_dynamic_conv_kh_u8_pad()
    TY, (H, Wd) = _ty(layer), (17, 140)
    rng = np.random.default_rng(3)
    u8 = rng.integers(0, 256, size=(3, H, Wd))
    hmax = (max(ks) - 1) // 2
    txo = TX - 2 * hmax
    assert pad >= hmax + 1
    # padded pixel-pair slots [H][Wd + 2 pad][8]
    slots = np.zeros((H, Wd + 2 * pad, 8))
    for xp in range(Wd + 2 * pad):
        for q in range(2):
            x = xp - pad + q
            if 0 <= x < Wd:
                slots[:, xp, 3 * q:3 * q + 3] = u8[:, :, x].T / 256.0
    npad = W.kh_layout(cout, ks[0])[0]
    slots_per = TY + 1
    imgs, o = [], 0
    for k in ks:
        nj, ncols = (k + 3) // 4, W.kh_layout(cout, k)[1]
        sz = nj * 2 * ncols * 8
        imgs.append([img[o + t * sz:o + (t + 1) * sz].reshape(nj, 2, ncols, 8) for t in range(2)])
        o += 2 * sz
    assert o == img.size
    out = [np.full((H, Wd, npad), np.nan) for _ in ks]
    acc = [np.zeros((slots_per * npad, TX)) for _ in ks]
    for ty in range(-(-H // TY)):
        for tx in range(-(-Wd // txo)):
            y0, x0 = ty * TY, max(0, min(tx * txo, Wd - txo))
            for R in range(max(y0 - hmax, 0), min(y0 + TY - 1 + hmax, H - 1) + 1):
                # the row segment the TMA box delivers: padded columns x0 - hmax + pad .. + 128 (+ the spill behind it)
                seg = np.zeros((TX + 16, 8))
                lo = x0 - hmax + pad
                hi = min(lo + TX + 16, Wd + 2 * pad)
                seg[:hi - lo] = slots[R, lo:hi]
                for b, k in enumerate(ks):
                    hb = (k - 1) // 2
                    ylo, yhi = max(y0, R - hb), min(min(y0 + TY - 1, H - 1), R + hb)
                    if ylo > yhi:
                        continue
                    g0, n = (ylo - (R - hb)) * npad, -(-((yhi - ylo + 1) * npad) // 16) * 16
                    d0 = (ylo - y0) * npad
                    for prod in range(2):
                        for j in range((k + 3) // 4):
                            A = np.concatenate([seg[hmax - hb + 4 * j + 2 * q:hmax - hb + 4 * j + 2 * q + TX] for q in range(2)], axis=1)   # [128, 16]
                            Bm = imgs[b][prod][j]
                            Bfull = np.concatenate((Bm[0], Bm[1]), axis=1)
                            acc[b][d0:d0 + n] += Bfull[g0:g0 + n] @ A.T
            for b in range(len(ks)):
                for s in range(TY):
                    y = y0 + s
                    if y < H:
                        for r in range(txo):
                            gx = x0 + r
                            if gx < Wd and gx >= tx * txo:
                                out[b][y, gx] = acc[b][s * npad:(s + 1) * npad, r]
                    acc[b][s * npad:(s + 1) * npad] = 0
    xin = torch.from_numpy(u8.astype(np.float64) / 255.0)
    for b, k in enumerate(ks):
        ref_f = F.conv2d(xin[None], pretrained_sd[f"{pre}.convs.{b}.weight"].double(), padding=(k - 1) // 2)[0]
        ref_a = F.conv2d(xin[None], pretrained_sd[f"{pre}.att_convs.{b}.weight"].double(), padding=(k - 1) // 2)[0]
        got = torch.from_numpy(out[b])
        assert not torch.isnan(got[..., :cout + 3]).any()
        assert (got[..., :cout].permute(2, 0, 1) - ref_f).abs().max() < 1e-5 * ref_f.abs().max() + 1e-7
        assert (got[..., cout:cout + 3].permute(2, 0, 1) - ref_a).abs().max() < 1e-5 * ref_a.abs().max() + 1e-7
