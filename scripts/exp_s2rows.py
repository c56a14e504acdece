#!/usr/bin/env python
"""Time cds_conv2d_3x3s2_rows at the cfg2 shapes (CDS_S2_DEBUG selects which stage is stubbed out; one process per setting)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cds_mvsnet_b200 import _lib, weights as W
from cds_mvsnet_b200._lib import call, ptr

for cin, cout, H, Wd in ((8, 16, 1184, 1600), (16, 32, 592, 800)):
    n = 8
    torch.manual_seed(0)
    x = torch.randn(2, n, H, Wd, cin, device="cuda").half()
    stats = torch.stack((x[0].double().sum((1, 2)), (x[0].double() ** 2).sum((1, 2))), -1).contiguous()
    wt = torch.randn(9, cin, cout) * 0.1
    packed = W.pack_conv2d_s2rows(wt).cuda()
    out = torch.empty(2, n, H // 2, Wd // 2, cout, device="cuda", dtype=torch.float16)
    ostats = torch.zeros(n, cout, 2, device="cuda", dtype=torch.float64)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    def run():
        call("cds_conv2d_3x3s2_rows", ptr(x[0]), ptr(x[1]), ptr(stats), 1, ptr(packed), n, cin, cout, H, Wd, ptr(out[0]), ptr(out[1]), ptr(ostats))
    for _ in range(3): run()
    ts = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    mb = (x.numel() + out.numel()) * 2 / 1e6
    t = sorted(ts)[len(ts) // 2]
    print(f"dbg={os.environ.get('CDS_S2_DEBUG', '0')} {cin}->{cout} {H}x{Wd}: {t:.3f} ms  {mb / t / 1e3:.2f} TB/s")
