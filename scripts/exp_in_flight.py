#!/usr/bin/env python
"""Experiment: throughput of K independent cascades (one CUDA graph each, own buffers) replayed on K streams at once
against one cascade at a time -- do the tails / small grids of one map's kernels leave room for a second map?

    python scripts/exp_in_flight.py [--workload cfg2] [--k 2] [--steps 20]
"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import cds_mvsnet_b200 as C  # noqa: E402
from cds_mvsnet_b200 import synthetic, weights as W  # noqa: E402
from cds_mvsnet_b200.engine import CascadeEngine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="cfg2")
ap.add_argument("--k", type=int, default=2)
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--offset-ms", type=float, default=0.0, help="delay lane 1's first replay by this long (phase experiment)")
args = ap.parse_args()
torch.set_grad_enabled(False)
cfg = dict(synthetic.CONFIGS[args.workload])
z = np.load(os.path.join(ROOT, "tests", "golden", "weights_both_dtu_blended.npz"))
model = C.CDSMVSNet(refine=False, ndepths=cfg["ndepths"], depth_interals_ratio=cfg["ratios"])
model.load_state_dict({k: torch.from_numpy(z[k]) for k in z.files})
model = model.cuda().eval()
dev = torch.device("cuda", 0)
s = synthetic.make_sample(cfg, "noise", seed=0)
imgs, dv = s.imgs.cuda(), s.depth_values.cuda()
proj = {k: v.cuda() for k, v in s.proj_matrices.items()}
mw = W.pack_model(model.state_dict(), model.num_stage, dev)
engines = [CascadeEngine(mw, model.ndepths, model.depth_interals_ratio, model.storage, dev) for _ in range(args.k)]
streams = [torch.cuda.Stream() for _ in range(args.k)]
outs = []
for e, st in zip(engines, streams):
    with torch.cuda.stream(st):
        outs.append(e.forward_graph(imgs, proj, dv, 0.01))
torch.cuda.synchronize()
ref = outs[0]["stage3"]["depth"].clone()


def run(k, steps, offset_ms=0.0):
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    ev0.record()
    for st in streams[:k]:
        st.wait_event(ev0)
    if offset_ms > 0 and k > 1:
        with torch.cuda.stream(streams[1]):
            torch.cuda._sleep(int(offset_ms * 1.9e6))   # ~1.9 GHz SM clock
    for i in range(steps):
        j = i % k
        with torch.cuda.stream(streams[j]):
            engines[j]._graph.replay()
    for st in streams[:k]:
        torch.cuda.current_stream().wait_stream(st)
    ev1.record()
    torch.cuda.synchronize()
    return ev0.elapsed_time(ev1) / steps


for k in range(1, args.k + 1):
    run(k, 4)
    ms = run(k, args.steps)
    print(f"in flight {k}: {ms:.3f} ms/map  {1e3 / ms:.2f} maps/s")
    if k > 1 and args.offset_ms > 0:
        for off in (args.offset_ms, 2 * args.offset_ms, 3 * args.offset_ms):
            ms = run(k, args.steps, off)
            # lane 0 works alone during the delay, so the truth lies between the raw figure and the one with the delay taken out
            print(f"in flight {k}, lane 1 delayed {off:.1f} ms: {ms:.3f} ms/map raw, {(ms * args.steps - off) / args.steps:.3f} with the delay taken out")
for o in outs:
    d = o["stage3"]["depth"]
    print("vs first pass of engine 0: max abs", float((d - ref).abs().max()), "mean rel", float(((d - ref).abs() / ref.abs()).mean()))
# the same engine, one map at a time, replay against replay
with torch.cuda.stream(streams[0]):
    engines[0]._graph.replay()
torch.cuda.synchronize()
a = outs[0]["stage3"]["depth"].clone()
with torch.cuda.stream(streams[0]):
    engines[0]._graph.replay()
torch.cuda.synchronize()
b = outs[0]["stage3"]["depth"]
print("replay vs replay (alone): max abs", float((a - b).abs().max()), "mean rel", float(((a - b).abs() / a.abs()).mean()))
