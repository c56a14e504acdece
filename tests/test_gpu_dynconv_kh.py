"""csrc/dynconv_kh.cu (trunk DynamicConv layers, kernel rows folded into N, persistent row streaming) against the oracle:
every layer shape, single-plane and split-precision input, the pair batch of the image layer, statistics and the curvature
accumulator modes, image sizes that exercise partial tiles / strips and several tiles per CTA."""
import ctypes

import pytest
import torch

from cds_mvsnet_b200 import _lib, weights as W
from cds_mvsnet_b200._lib import call, ptr
from oracle import oracle as O

pytestmark = pytest.mark.gpu
T = 0.01
DEV = "cuda"
torch.set_grad_enabled(False)
cu = lambda t: t.to(DEV)


def _weights(sd, name):
    cin, cout, ks, pre = W.DYN_LAYERS[name]
    w = W.pack_dynamic_conv(sd, pre, cin, cout, ks, DEV)
    w.kh = W.pack_dynamic_conv_kh(w)
    return w, cin, cout, ks, pre


def _stats(x):
    return torch.stack((x.double().sum((2, 3)), (x.double() ** 2).sum((2, 3))), -1).contiguous()


def _run_kh(w, planes, n_images, img_index, stats, in_act, epi, n, cin_k, cout, hw, ks, split, pv=0, pb=0, nc_mode=0, ncsq=None, bias=None):
    out = torch.full((n, *hw, cout), float("nan"), device=DEV, dtype=torch.float16)
    out_lo = torch.full_like(out, float("nan"))
    ostats = torch.zeros(n, cout, 2, device=DEV, dtype=torch.float64)
    nc = torch.full((n, *hw), float("nan"), device=DEV)
    ncsq = torch.full((n, *hw), float("nan"), device=DEV) if ncsq is None else ncsq
    ncabs = torch.full((n, *hw), float("nan"), device=DEV)
    kz = (ctypes.c_int * len(ks))(*ks)
    call("cds_dynamic_conv_kh", ptr(planes), n_images, ptr(img_index), ptr(stats), in_act, ptr(epi), 1.0, ptr(w.kh), ptr(bias), ptr(w.gate),
         n, cin_k, cout, hw[0], hw[1], len(ks), kz, T, int(split), ptr(out), ptr(out_lo), ptr(ostats), ptr(nc), ptr(ncsq), nc_mode,
         ptr(ncabs), pv, pb)
    torch.cuda.synchronize()
    return out, out_lo, ostats, nc, ncsq, ncabs


@pytest.mark.parametrize("name", ["conv01", "conv10", "conv20"])
@pytest.mark.parametrize("hw", [(24, 40), (37, 150), (64, 300), (7, 128), (150, 700)], ids=lambda hw: f"{hw[0]}x{hw[1]}")   # the last: several tiles per persistent CTA
@pytest.mark.parametrize("split", [1, 0])
def test_kh_trunk_layer_vs_oracle(pretrained_sd, name, hw, split):
    w, cin, cout, ks, pre = _weights(pretrained_sd, name)
    torch.manual_seed(hw[1] + len(name))
    n = 3
    x = 1.7 + 0.8 * torch.randn(n, cin, *hw)                       # raw pre-norm activations with a mean offset
    epi = torch.tensor([[hw[1] * 1.7, -hw[0] * 0.6], [-30.0, hw[0] / 2.0], [hw[1] / 3.0, hw[0] * 2.0]])
    nhwc = x.permute(0, 2, 3, 1).contiguous()
    hi = nhwc.half()
    lo = (nhwc - hi.float()).half()
    stored = (hi.float() + lo.float()) if split else hi.float()      # what the kernel is given
    xs = stored.permute(0, 3, 1, 2)
    xin = torch.nn.functional.leaky_relu(O.instance_norm(xs), 0.1)
    ref_y, ref_nc = O.dynamic_conv(xin, pretrained_sd, pre, ks, epi, T)
    planes = cu(torch.stack((hi, lo)) if split else hi.unsqueeze(0)).contiguous()
    out, out_lo, ostats, nc, ncsq, ncabs = _run_kh(w, planes, n, None, cu(_stats(xs)), 1, cu(epi), n, cin, cout, hw, ks, split)
    y = (out.float() + out_lo.float()).cpu().permute(0, 3, 1, 2)
    ey, en = O.rel_l1(y, ref_y), O.rel_l1(nc.cpu().unsqueeze(1), ref_nc)
    print(f"kh {name} {hw} split={split}: out {ey:.2e} curv {en:.2e}")
    assert not torch.isnan(y).any() and not torch.isnan(nc).any()
    tol = 2e-4 if split else 4e-3      # split: only the product terms lo x lo are dropped; single plane: fp16 operand rounding
    assert ey < tol and en < tol
    # the residual plane is what fp16 rounding of the value plane dropped: below half an ulp of the value
    assert (out_lo.float().abs() <= out.float().abs() * 2.0 ** -11 + 1e-7).all()
    # statistics of what was written (fp32 accumulators), curvature maps
    got = ostats.cpu()
    want = _stats(y)
    torch.testing.assert_close(got, want, rtol=2e-3, atol=2e-2 * hw[0] * hw[1] ** 0.5)
    torch.testing.assert_close(ncsq.cpu(), nc.cpu() ** 2, rtol=1e-6, atol=1e-9)
    torch.testing.assert_close(ncabs.cpu(), nc.cpu().abs(), rtol=0, atol=0)


def test_kh_curvature_accumulator_modes(pretrained_sd):
    w, cin, cout, ks, pre = _weights(pretrained_sd, "conv10")
    hw, n = (20, 140), 2
    torch.manual_seed(9)
    x = torch.randn(n, cin, *hw)
    epi = torch.tensor([[10.0, 5.0], [300.0, -40.0]])
    hi = x.permute(0, 2, 3, 1).contiguous().half()
    st = cu(_stats(hi.float().permute(0, 3, 1, 2)))
    planes = cu(hi.unsqueeze(0)).contiguous()
    base = torch.rand(n, *hw)
    _, _, _, nc, _, _ = _run_kh(w, planes, n, None, st, 1, cu(epi), n, cin, cout, hw, ks, 0)
    for mode, want in ((1, lambda b, c: b + c * c), (2, lambda b, c: (b + c * c) / 3.0)):
        acc = cu(base.clone())
        _run_kh(w, planes, n, None, st, 1, cu(epi), n, cin, cout, hw, ks, 0, nc_mode=mode, ncsq=acc)
        torch.testing.assert_close(acc.cpu(), want(base, nc.cpu()), rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("V,B,hw", [(4, 1, (24, 150)), (2, 2, (33, 130)), (5, 1, (8, 40)), (1, 1, (9, 128)), (4, 1, (200, 900)),
                                    (4, 1, (105, 1300))],   # partial last row block FOLLOWED by more tiles on the same CTA, slow 4-pair epilogue
                         ids=lambda v: f"{v[0]}x{v[1]}" if isinstance(v, tuple) else str(v))
def test_kh_image_layer_pair_batch(pretrained_sd, V, B, hw):
    """conv00 over the cascade's (side, v, b) pair batch (the reference image's MMAs shared by its pairs) and as a plain batch:
    both against the oracle, value + residual planes."""
    w, cin, cout, ks, pre = _weights(pretrained_sd, "conv00")
    torch.manual_seed(V * 10 + B)
    N = V + 1
    imgs = torch.rand(B * N, 3, *hw)
    n = 2 * V * B
    idx = torch.empty(2, V, B, dtype=torch.int32)
    for v in range(V):
        for b in range(B):
            idx[0, v, b], idx[1, v, b] = b * N, b * N + v + 1
    idx = idx.reshape(-1)
    epi = torch.randn(n, 2) * hw[1]
    img8 = torch.empty(B * N, *hw, 8, device=DEV, dtype=torch.float16)
    ic = cu(imgs)
    call("cds_image_to_nhwc8", ptr(ic), B * N, hw[0], hw[1], ptr(img8))
    x_items = torch.stack([imgs[int(i)] for i in idx])
    ref_y, ref_nc = O.dynamic_conv(x_items, pretrained_sd, pre, ks, epi, T)
    res = []
    for pv, pb in ((V, B), (0, 0)):
        out, out_lo, ostats, nc, ncsq, _ = _run_kh(w, img8, B * N, cu(idx), None, 0, cu(epi), n, 8, cout, hw, ks, 0, pv=pv, pb=pb)
        y = (out.float() + out_lo.float()).cpu().permute(0, 3, 1, 2)
        assert not torch.isnan(y).any() and not torch.isnan(nc).any()
        ey, en = O.rel_l1(y, ref_y), O.rel_l1(nc.cpu().unsqueeze(1), ref_nc)
        print(f"kh conv00 pairs={pv > 0} {hw}: out {ey:.2e} curv {en:.2e}")
        assert ey < 5e-5 and en < 5e-5          # image and weights both carry their residuals: fp32-level accuracy
        torch.testing.assert_close(ostats.cpu(), _stats(y), rtol=2e-3, atol=2e-2 * hw[0] * hw[1] ** 0.5)
        res.append(y)
    assert O.rel_l1(res[0], res[1]) < 1e-6


@pytest.mark.parametrize("name", ["out1", "out2", "out3"])
@pytest.mark.parametrize("hw", [(24, 40), (45, 300), (150, 700)], ids=lambda hw: f"{hw[0]}x{hw[1]}")
def test_kh_output_heads_vs_oracle(pretrained_sd, name, hw):
    """The three output heads (1x1 + 3x3 branches, with bias, curvature accumulator in its closing mode)."""
    w, cin, cout, ks, pre = _weights(pretrained_sd, name)
    torch.manual_seed(hw[0] + len(name))
    n = 2
    x = 0.3 + torch.randn(n, cin, *hw)
    epi = torch.tensor([[hw[1] * 0.7, hw[0] * 1.6], [-300.0, hw[0] / 2.0]])
    hi = x.permute(0, 2, 3, 1).contiguous().half()
    xs = hi.float().permute(0, 3, 1, 2)
    xin = torch.nn.functional.leaky_relu(O.instance_norm(xs), 0.1)
    ref_y, ref_nc = O.dynamic_conv(xin, pretrained_sd, pre, ks, epi, T)
    base = torch.rand(n, *hw)
    acc = cu(base.clone())
    assert w.bias is not None
    out, out_lo, ostats, nc, ncsq, ncabs = _run_kh(w, cu(hi.unsqueeze(0)).contiguous(), n, None, cu(_stats(xs)), 1, cu(epi), n, cin, cout, hw,
                                                   ks, 0, nc_mode=2, ncsq=acc, bias=w.bias)
    y = (out.float() + out_lo.float()).cpu().permute(0, 3, 1, 2)
    ey, en = O.rel_l1(y, ref_y), O.rel_l1(nc.cpu().unsqueeze(1), ref_nc)
    print(f"kh {name} {hw}: out {ey:.2e} curv {en:.2e}")
    assert not torch.isnan(y).any() and not torch.isnan(nc).any()
    assert ey < 4e-3 and en < 4e-3
    torch.testing.assert_close(ostats.cpu(), _stats(y), rtol=2e-3, atol=2e-2 * hw[0] * hw[1] ** 0.5)
    torch.testing.assert_close(acc.cpu(), (base + nc.cpu() ** 2) / 3.0, rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(ncabs.cpu(), nc.cpu().abs(), rtol=0, atol=0)


def test_kh_rejects_unsupported():
    z = torch.zeros(64, device=DEV)
    kz = (ctypes.c_int * 2)(3, 3)
    with pytest.raises(RuntimeError):   # 8 -> 8 (3,3) is not a layer of the feature extractor
        call("cds_dynamic_conv_kh", ptr(z), 1, None, None, 0, ptr(z), 1.0, ptr(z), None, ptr(z), 1, 8, 8, 16, 16, 2, kz, T, 0, ptr(z),
             None, None, None, None, 0, None, 0, 0)
    assert _lib.LIB.load().cds_dynamic_conv_kh_supported(8, 8, 16, 16, 2, kz) == 0


@pytest.mark.parametrize("V,B,hw", [(4, 1, (24, 150)), (2, 2, (33, 130)), (1, 1, (9, 128)), (3, 1, (40, 8)), (4, 1, (105, 1300))],
                         ids=lambda v: f"{v[0]}x{v[1]}" if isinstance(v, tuple) else str(v))
def test_kh_image_layer_on_8bit_images(pretrained_sd, V, B, hw):
    """conv00 on 8-bit images: pixel-pair operand slots (byte / 256, four horizontal taps per MMA, 256 / 255 folded into the
    weights) against the oracle on float32(byte) / 255 -- the pair batch and the plain batch, value + residual planes -- and
    against the fp32-image path of the same kernel."""
    w, cin, cout, ks, pre = _weights(pretrained_sd, "conv00")
    w.kh_u8 = W.pack_dynamic_conv_kh(w, px2=True)
    lib = _lib.LIB.load()
    kz = (ctypes.c_int * len(ks))(*ks)
    assert w.kh_u8.numel() == lib.cds_dynamic_conv_kh_u8_weight_halfs(cout, len(ks), kz)
    torch.manual_seed(V * 10 + B + hw[1])
    N = V + 1
    u8 = torch.randint(0, 256, (B * N, 3, *hw), dtype=torch.uint8)
    imgs = u8.float() / 255.0
    n = 2 * V * B
    idx = torch.empty(2, V, B, dtype=torch.int32)
    for v in range(V):
        for b in range(B):
            idx[0, v, b], idx[1, v, b] = b * N, b * N + v + 1
    idx = idx.reshape(-1)
    epi = torch.randn(n, 2) * hw[1]
    x_items = torch.stack([imgs[int(i)] for i in idx])
    ref_y, ref_nc = O.dynamic_conv(x_items, pretrained_sd, pre, ks, epi, T)
    pad = lib.cds_dynamic_conv_kh_u8_pad()
    px2 = torch.full((B * N, hw[0], hw[1] + 2 * pad, 8), float("nan"), device=DEV, dtype=torch.float16)
    uc = cu(u8)
    call("cds_image_u8_to_px2", ptr(uc), B * N, hw[0], hw[1], ptr(px2))
    torch.cuda.synchronize()
    assert not torch.isnan(px2).any()
    got = px2[:, :, pad:pad + hw[1], :3].float().cpu() * 256.0
    assert torch.equal(got, u8.permute(0, 2, 3, 1).float())                                  # exact bytes, own pixel
    assert torch.equal(px2[:, :, pad - 1, 3:6].float().cpu() * 256.0, u8[:, :, :, 0].permute(0, 2, 1).float())   # left pad slot: its right neighbour
    assert px2[:, :, pad + hw[1] - 1, 3:6].abs().max() == 0 and px2[..., 6:].abs().max() == 0
    res = []
    for pv, pb in ((V, B), (0, 0)):
        out = torch.full((n, *hw, cout), float("nan"), device=DEV, dtype=torch.float16)
        out_lo = torch.full_like(out, float("nan"))
        ostats = torch.zeros(n, cout, 2, device=DEV, dtype=torch.float64)
        nc = torch.full((n, *hw), float("nan"), device=DEV)
        ncsq = torch.full((n, *hw), float("nan"), device=DEV)
        ic, ec = cu(idx), cu(epi)
        call("cds_dynamic_conv_kh_u8", ptr(px2), B * N, ptr(ic), ptr(ec), 1.0, ptr(w.kh_u8), ptr(w.gate), n, cout, hw[0], hw[1], len(ks), kz, T,
             ptr(out), ptr(out_lo), ptr(ostats), ptr(nc), ptr(ncsq), 0, None, pv, pb)
        torch.cuda.synchronize()
        y = (out.float() + out_lo.float()).cpu().permute(0, 3, 1, 2)
        assert not torch.isnan(y).any() and not torch.isnan(nc).any()
        ey, en = O.rel_l1(y, ref_y), O.rel_l1(nc.cpu().unsqueeze(1), ref_nc)
        print(f"kh conv00 u8 pairs={pv > 0} {hw}: out {ey:.2e} curv {en:.2e}")
        assert ey < 5e-5 and en < 5e-5
        torch.testing.assert_close(ostats.cpu(), _stats(y), rtol=2e-3, atol=2e-2 * hw[0] * hw[1] ** 0.5)
        res.append(y)
    assert O.rel_l1(res[0], res[1]) < 1e-6
    # the fp32-image path of the same layer
    img8 = torch.empty(B * N, *hw, 8, device=DEV, dtype=torch.float16)
    fc = cu(imgs)
    call("cds_image_to_nhwc8", ptr(fc), B * N, hw[0], hw[1], ptr(img8))
    out, out_lo, _, _, _, _ = _run_kh(w, img8, B * N, cu(idx), None, 0, cu(epi), n, 8, cout, hw, ks, 0, pv=V, pb=B)
    assert O.rel_l1((out.float() + out_lo.float()).cpu().permute(0, 3, 1, 2), res[0]) < 5e-5
