"""The torch operators the oracle delegates to == their published definitions (plain-C restatement)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
CDIR = os.path.join(os.path.dirname(HERE), "oracle", "c")
FP = ctypes.POINTER(ctypes.c_float)


@pytest.fixture(scope="module")
def clib():
    subprocess.run(["make", "-s", "-C", CDIR], check=True)
    return ctypes.CDLL(os.path.join(CDIR, "libcds_oracle.so"))


def fp(a):
    return a.ctypes.data_as(FP)


def arr(t):
    return np.ascontiguousarray(t.numpy().astype(np.float32))


def test_conv2d(clib):
    torch.manual_seed(0)
    for (ci, co, k, s, p) in ((3, 5, 7, 1, 3), (8, 4, 3, 2, 1), (6, 3, 1, 1, 0)):
        x, w, b = torch.randn(2, ci, 9, 11), torch.randn(co, ci, k, k), torch.randn(co)
        ref = F.conv2d(x, w, b, stride=s, padding=p)
        y = np.zeros(tuple(ref.shape), np.float32)
        xa, wa, ba = arr(x), arr(w), arr(b)
        clib.cds_c_conv2d(fp(xa), fp(wa), fp(ba), 2, ci, 9, 11, co, k, s, p, fp(y))
        np.testing.assert_allclose(y, ref.numpy(), atol=2e-5, rtol=1e-5)


@pytest.mark.parametrize("stride", [1, 2])
def test_conv3d(clib, stride):
    torch.manual_seed(1)
    x, w = torch.randn(1, 4, 5, 6, 7), torch.randn(3, 4, 3, 3, 3)
    ref = F.conv3d(x, w, stride=stride, padding=1)
    y = np.zeros(tuple(ref.shape), np.float32)
    xa, wa = arr(x), arr(w)
    clib.cds_c_conv3d(fp(xa), fp(wa), 1, 4, 5, 6, 7, 3, stride, fp(y))
    np.testing.assert_allclose(y, ref.numpy(), atol=2e-5, rtol=1e-5)


def test_conv_transpose3d(clib):
    torch.manual_seed(2)
    x, w = torch.randn(1, 3, 2, 3, 4), torch.randn(3, 5, 3, 3, 3)
    ref = F.conv_transpose3d(x, w, stride=2, padding=1, output_padding=1)
    y = np.zeros(tuple(ref.shape), np.float32)
    xa, wa = arr(x), arr(w)
    clib.cds_c_conv_transpose3d(fp(xa), fp(wa), 1, 3, 2, 3, 4, 5, fp(y))
    np.testing.assert_allclose(y, ref.numpy(), atol=2e-5, rtol=1e-5)


def test_bilinear_gather_matches_oracle_and_grid_sample(clib):
    torch.manual_seed(3)
    fea = torch.randn(1, 4, 7, 9)
    u = torch.rand(1, 50) * 12 - 2
    v = torch.rand(1, 50) * 10 - 2
    mine = O.bilinear_gather_zeros(fea, u, v)
    # the oracle applies the reference's normalise/unnormalise round trip; feed the C version the same coordinates
    un, vn = u / ((9 - 1) / 2) - 1, v / ((7 - 1) / 2) - 1
    u2, v2 = (un + 1) / 2 * (9 - 1), (vn + 1) / 2 * (7 - 1)
    out = np.zeros((4, 50), np.float32)
    fa, ua, va = arr(fea[0]), arr(u2[0]), arr(v2[0])
    clib.cds_c_bilinear_zeros(fp(fa), 4, 7, 9, fp(ua), fp(va), 50, fp(out))
    np.testing.assert_allclose(out, mine[0].numpy(), atol=1e-6)
    grid = torch.stack((un, vn), -1).reshape(1, 1, 50, 2)
    gs = F.grid_sample(fea, grid, mode="bilinear", padding_mode="zeros", align_corners=True)[0, :, 0]
    np.testing.assert_allclose(out, gs.numpy(), atol=1e-5)


def test_softmax_regress(clib):
    torch.manual_seed(4)
    D, P = 8, 40
    logits, depth = 3 * torch.randn(D, P), 400 + 500 * torch.rand(D, P)
    p = torch.softmax(logits.reshape(1, D, P, 1), 1)
    ref_d = O.depth_regression(p, depth.reshape(1, D, P, 1))[0, :, 0]
    ref_c = O.conf_regression(p)[0, :, 0]
    d_out, c_out = np.zeros(P, np.float32), np.zeros(P, np.float32)
    la, da = arr(logits), arr(depth)
    clib.cds_c_softmax_regress(fp(la), fp(da), D, P, fp(d_out), fp(c_out))
    np.testing.assert_allclose(d_out, ref_d.numpy(), rtol=1e-5)
    assert (np.abs(c_out - ref_c.numpy()) > 1e-5).mean() < 0.05   # window index is a truncation


def test_bilinear_backward_and_stage_loss(clib):
    """Training slice: the plain-C scatter (adjoint of the gather) and stage loss against the torch-CPU restatements."""
    torch.manual_seed(6)
    C_, h, w, M = 3, 7, 9, 60
    u = torch.rand(1, M) * (w + 3) - 2            # some samples fall outside the image
    v = torch.rand(1, M) * (h + 3) - 2
    g = torch.randn(1, C_, M)
    want = O.bilinear_scatter_zeros(g, u, v, h, w)[0]
    # the python restatement goes through the reference's normalise / un-normalise round trip; feed C the same coordinates
    un, vn = u / ((w - 1) / 2) - 1, v / ((h - 1) / 2) - 1
    u2, v2 = (un + 1) / 2 * (w - 1), (vn + 1) / 2 * (h - 1)
    out = np.zeros((C_, h, w), np.float32)
    ga, ua, va = arr(g[0]), arr(u2[0]), arr(v2[0])
    clib.cds_c_bilinear_zeros_backward(fp(ga), C_, h, w, fp(ua), fp(va), M, fp(out))
    np.testing.assert_allclose(out, want.numpy(), atol=2e-6)
    # adjoint identity against the C gather
    fea = torch.randn(C_, h, w)
    fwd = np.zeros((C_, M), np.float32)
    fa = arr(fea)
    clib.cds_c_bilinear_zeros(fp(fa), C_, h, w, fp(ua), fp(va), M, fp(fwd))
    assert abs((fwd.astype(np.float64) * ga).sum() - (fa.astype(np.float64) * out).sum()) < 1e-4

    n = 500
    gt = 425 + 500 * torch.rand(n)
    est = gt + 4 * torch.randn(n)
    mask, iv, curv = (torch.rand(n) > 0.3).float(), torch.full((n,), 2.65), torch.rand(n)
    res = np.zeros(2, np.float32)
    clib.cds_c_stage_loss.argtypes = [FP, FP, FP, FP, FP, ctypes.c_long, FP]
    ea, gta, ma, iva, ca = arr(est), arr(gt), arr(mask), arr(iv), arr(curv)
    clib.cds_c_stage_loss(fp(ea), fp(gta), fp(ma), fp(iva), fp(ca), n, fp(res))
    dl, cm = O.stage_loss(est.reshape(1, 1, n), gt.reshape(1, 1, n), mask.reshape(1, 1, n), torch.tensor([2.65]), curv.reshape(1, 1, n))
    np.testing.assert_allclose(res, [dl.item(), cm.item()], rtol=2e-6)
