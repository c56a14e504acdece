#!/usr/bin/env python
"""One call of cds_conv2d_3x3s2_rows at a size with several tiles per persistent CTA, checked against the gather form."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cds_mvsnet_b200 import _lib, weights as W
from cds_mvsnet_b200._lib import call, ptr

for cin, cout, H, Wd, n in ((8, 16, 600, 800, 2), (16, 32, 300, 800, 3)):
    torch.manual_seed(0)
    x = torch.randn(2, n, H, Wd, cin, device="cuda").half()
    x[1] *= 1e-3
    xs = (x[0].float() + x[1].float())
    stats = torch.stack((xs.double().sum((1, 2)), (xs.double() ** 2).sum((1, 2))), -1).contiguous()
    wt = torch.randn(9, cin, cout) * 0.1
    packed = W.pack_conv2d_s2rows(wt).cuda()
    packed_g = W.pack_conv2d_gtc(wt).cuda()
    out = torch.zeros(2, n, H // 2, Wd // 2, cout, device="cuda", dtype=torch.float16)
    ref = torch.zeros_like(out)
    ostats = torch.zeros(n, cout, 2, device="cuda", dtype=torch.float64)
    call("cds_conv2d_3x3s2_rows", ptr(x[0]), ptr(x[1]), ptr(stats), 1, ptr(packed), n, cin, cout, H, Wd, ptr(out[0]), ptr(out[1]), ptr(ostats))
    torch.cuda.synchronize()
    call("cds_conv2d_3x3s2_tc", ptr(x[0]), ptr(x[1]), ptr(stats), 1, ptr(packed_g), n, cin, cout, H, Wd, ptr(ref[0]), ptr(ref[1]), None)
    torch.cuda.synchronize()
    a, b = out[0].float() + out[1].float(), ref[0].float() + ref[1].float()
    print(cin, cout, "max diff vs gather form", float((a - b).abs().max()), "scale", float(b.abs().max()))
