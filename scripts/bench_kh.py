"""GPU micro-benchmark: trunk DynamicConv layers at cfg2 sizes on csrc/dynconv_kh.cu vs csrc/dynconv_tc.cu, split / single plane."""
import ctypes, sys, json
import numpy as np, torch
sys.path.insert(0, ".")
from cds_mvsnet_b200 import _lib, weights as W
from cds_mvsnet_b200._lib import call, ptr
torch.set_grad_enabled(False)
z = np.load("tests/golden/weights_both_dtu_blended.npz"); sd = {k: torch.from_numpy(z[k]) for k in z.files}
DEV = "cuda"
res = {}
def timeit(fn, iters=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters
for name, hw, n in (("conv01", (1184, 1600), 8), ("conv10", (592, 800), 8), ("conv20", (296, 400), 8)):
    cin, cout, ks, pre = W.DYN_LAYERS[name]
    w = W.pack_dynamic_conv(sd, pre, cin, cout, ks, DEV); w.tc = W.pack_dynamic_conv_tc(w); w.kh = W.pack_dynamic_conv_kh(w)
    planes = torch.randn(2, n, *hw, cin, device=DEV).half()
    planes[1] *= 1e-3
    stats = torch.stack((torch.zeros(n, cin, dtype=torch.float64), torch.full((n, cin), float(hw[0] * hw[1]), dtype=torch.float64)), -1).to(DEV).contiguous()
    epi = (torch.randn(n, 2) * 500).to(DEV)
    out = torch.empty(n, *hw, cout, device=DEV, dtype=torch.float16); out_lo = torch.empty_like(out)
    ostats = torch.zeros(n, cout, 2, device=DEV, dtype=torch.float64); ncsq = torch.zeros(n, *hw, device=DEV)
    kz = (ctypes.c_int * len(ks))(*ks)
    for split in (1, 0):
        res[f"{name} kh split={split}"] = timeit(lambda: call("cds_dynamic_conv_kh", ptr(planes), n, None, ptr(stats), 1, ptr(epi), 1.0, ptr(w.kh), None, ptr(w.gate),
            n, cin, cout, hw[0], hw[1], len(ks), kz, 0.01, split, ptr(out), ptr(out_lo), ptr(ostats), None, ptr(ncsq), 0, None, 0, 0))
        res[f"{name} tc split={split}"] = timeit(lambda: call("cds_dynamic_conv_tc", ptr(planes), n, None, ptr(stats), 1, ptr(epi), 1.0, ptr(w.tc), None, ptr(w.gate),
            n, cin, cout, hw[0], hw[1], len(ks), kz, 0.01, split, ptr(out), ptr(out_lo), ptr(ostats), None, ptr(ncsq), 0, None))
    del planes, out, out_lo
for k, v in res.items(): print(f"{k:28s} {v:.3f} ms")
json.dump(res, open("gpurun_out/bench_kh.json", "w"), indent=1)
