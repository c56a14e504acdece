#!/usr/bin/env python
"""Benchmark of the depth-inference hot path: depth-maps/sec at BASELINE.json's configuration.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2]

A "step" is one full CDSMVSNet.forward (feature extractor for 2(N-1) images + the 3-stage cascade) over
one batch of synthetic input.  ``value`` = depth maps / s with inputs resident in HBM (device-timed with
CUDA events, max over ranks); ``e2e`` = the same through the public drop-in API with HOST (pinned)
inputs, host->device and device->host copies inside the timed region.  One JSON line on stdout.

Multi-GPU: one process per GPU (torchrun), every rank runs its own replica on its own work items (depth
maps are independent units -- SURVEY.md 8e); there is no data-path collective, ``scaling`` is "weak".

``--impl reference`` times the reference's own algorithm on the host CPU cores: the reference is pure
Python/PyTorch that cannot travel to the GPU box, so the arm runs the oracle port (oracle/oracle.py,
pinned against the live reference by tests/golden) with all host threads on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "depth_maps_per_sec"
UNIT = "maps/s"
TEMPERATURE = 0.01   # reference test.py:52


def load_weights():
    z = np.load(os.path.join(ROOT, "tests", "golden", "weights_both_dtu_blended.npz"))
    return {k: torch.from_numpy(z[k]) for k in z.files}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=float(d["hbm_gbs"]), tensor=float(d["bf16_tflops_sustained"]), src="measured (MEASURED_PEAKS.json; "
                    "tensor = sustained dense bf16/fp16 pipe)")
    return dict(hbm=6650.0, tensor=1400.0, src="fallback (B200_PROFILING.md)")


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "200"], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons, pw = [], [], set(), []
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
            except (ValueError, IndexError):
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].strip().lower() == "active":
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
def cpu_oracle_rate(cfg, n_threads, budget_s=25.0, steps=1, warmup=0):
    """Depth maps / s of the oracle port on the host, measured on a bounded sample of the workload.

    The sample keeps N and the depth-plane counts and shrinks the image (cost is linear in pixels); the
    rate is scaled back by the pixel ratio and reported for the FULL workload."""
    from cds_mvsnet_b200 import synthetic
    from oracle import oracle as O
    O.FAST_GATHER = True   # gather through ATen grid_sample, as the reference itself does
    torch.set_num_threads(n_threads)
    sd = load_weights()
    refine = bool(cfg.get("refine", False))
    if refine:
        zr = np.load(os.path.join(ROOT, "tests", "golden", "weights_refine_both_dtu_blended.npz"))
        sd.update({k: torch.from_numpy(zr[k]) for k in zr.files})
    q = 64 if refine else 32
    full_px = cfg["H"] * cfg["W"]
    ladder = [(cfg["H"], cfg["W"])]
    for f in (2, 4, 8):
        h, w = max(q, (cfg["H"] // f) // q * q), max(q, (cfg["W"] // f) // q * q)
        if (h, w) != ladder[-1]:
            ladder.append((h, w))

    def run(h, w):
        c = dict(cfg, H=h, W=w, B=1)
        s = synthetic.make_sample(c, "noise", seed=0)
        t0 = time.perf_counter()
        with torch.no_grad():
            O.cdsmvsnet_forward(sd, s.imgs, s.proj_matrices, s.depth_values, c["ndepths"], c["ratios"], TEMPERATURE, refine=refine)
        return time.perf_counter() - t0

    # calibrate on the smallest rung, then take the largest rung whose predicted total fits the budget
    h0, w0 = ladder[-1]
    t_small = run(h0, w0)
    per_px = t_small / (h0 * w0)
    n_runs = steps + warmup
    pick = ladder[-1]
    for h, w in ladder:
        if per_px * h * w * n_runs <= budget_s:
            pick = (h, w)
            break
    times = []
    for i in range(n_runs):
        t = run(*pick)
        if i >= warmup:
            times.append(t)
    t_step = float(np.mean(times))
    rate = (1.0 / t_step) * (pick[0] * pick[1] / full_px)
    sample = (f"oracle port, full cascade N={cfg['N']} D={list(cfg['ndepths'])} on a {pick[1]}x{pick[0]} image "
              f"({pick[0] * pick[1] / full_px:.3f} of the {cfg['W']}x{cfg['H']} workload's pixels), {len(times)} run(s) of "
              f"{t_step:.1f} s, rate scaled by the pixel ratio")
    return rate, t_step, sample


def run_reference_arm(args, cfg, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    rate, t_step, sample = cpu_oracle_rate(cfg, cores, budget_s=150.0, steps=args.steps, warmup=args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 / rate, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.workload, cfg), "note": "reference algorithm on host CPU cores (oracle port; "
                   "the pure-Python reference tree does not exist on the GPU box)"},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def workload_name(key, cfg):
    return (f"{key}: CDSMVSNet.forward {cfg['W']}x{cfg['H']} N={cfg['N']} D={'/'.join(str(d) for d in cfg['ndepths'])} "
            f"ratios={'/'.join(str(r) for r in cfg['ratios'])} B={cfg['B']} refine={bool(cfg.get('refine', False))} T={TEMPERATURE}")


# ------------------------------------------------------------------------------------------------
_JSON_FD = None


def _claim_stdout():
    """Everything that writes to stdout while the job runs (NCCL's version banner, library chatter) is sent to stderr; the one
    JSON line goes to the real stdout at the end."""
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--storage", default="float16", choices=["float16", "float32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="issue the forward launch by launch instead of replaying a CUDA graph")
    ap.add_argument("--kernel-table", default=None, help="write the per-kernel timing table (JSON) here")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3

    from cds_mvsnet_b200 import synthetic
    cfg = dict(synthetic.CONFIGS[args.workload])
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference_arm(args, cfg, rank, world)
        return

    import torch.distributed as dist

    import cds_mvsnet_b200 as C
    from cds_mvsnet_b200 import _lib
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    torch.set_grad_enabled(False)

    storage = getattr(torch, args.storage)
    refine = bool(cfg.get("refine", False))
    model = C.CDSMVSNet(refine=refine, ndepths=cfg["ndepths"], depth_interals_ratio=cfg["ratios"], storage=storage)
    sd = load_weights()
    if refine:
        zr = np.load(os.path.join(ROOT, "tests", "golden", "weights_refine_both_dtu_blended.npz"))
        sd.update({k: torch.from_numpy(zr[k]) for k in zr.files})
    model.load_state_dict({k: v for k, v in sd.items() if k in model.state_dict()})   # fewer stages (cfg1): fewer entries
    model = model.to(dev).eval()

    # every rank works on its own depth map (different seed => different work item)
    s = synthetic.make_sample(cfg, "plane", seed=rank)
    host = {"imgs": s.imgs.pin_memory(), "dv": s.depth_values.pin_memory(),
            "proj": {k: v.pin_memory() for k, v in s.proj_matrices.items()}}
    d_imgs, d_dv = host["imgs"].to(dev), host["dv"].to(dev)
    d_proj = {k: v.to(dev) for k, v in host["proj"].items()}
    B = cfg["B"]
    engine = model.engine(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    use_graph = not args.no_graph

    def step_device():
        # one CUDA-graph launch per forward (inputs copied device->device into the graph's static tensors every step)
        if use_graph:
            return engine.forward_graph(d_imgs, d_proj, d_dv, TEMPERATURE)
        return engine.forward(d_imgs, d_proj, d_dv, TEMPERATURE)

    for _ in range(args.warmup):
        step_device()
    barrier()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()

    # ---- timed region 1: EXACTLY K steps, inputs resident in HBM, one pair of CUDA events around the region -------------
    l0 = _lib.LAUNCHES
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        step_device()
    ev1.record()
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    launches = _lib.LAUNCHES - l0

    # ---- timed region 1b: the same K steps again with a CUDA-event pair around EVERY launch (the per-kernel durations of
    # the roofline and of the kernel table; the ~140 extra event records per step cost ~3 %, so `value` is not taken here)
    ev2, ev3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    overlap_was, engine.overlap = engine.overlap, False   # one stream: concurrent kernels would share the SMs and blur each other's durations
    with _lib.LaunchProfile() as prof:
        ev2.record()
        for _ in range(args.steps):
            engine.forward(d_imgs, d_proj, d_dv, TEMPERATURE)   # launch by launch: a graph node cannot carry an event pair
        ev3.record()
    barrier()
    ms_instrumented = ev2.elapsed_time(ev3)
    table = prof.summary()
    engine.overlap = overlap_was

    # ---- timed region 2: end to end through the public API with HOST buffers -----------------------------------
    # (a) one call at a time: model(...) on pinned host inputs, results read back, host blocks on every item
    out_host = None

    def step_e2e():
        nonlocal out_host
        imgs = host["imgs"].to(dev, non_blocking=True)
        dv = host["dv"].to(dev, non_blocking=True)
        proj = {k: v.to(dev, non_blocking=True) for k, v in host["proj"].items()}
        out = model(imgs, proj, dv, temperature=TEMPERATURE)
        flat = {f"{k}.{kk}": vv for k, v in out.items() if isinstance(v, dict) for kk, vv in v.items()}
        if getattr(model, "refine", False):   # refine=True: the full-resolution refined depth is a map of its own (the streamed leg returns it too)
            flat["refined_depth"] = out["refined_depth"]
        if out_host is None:
            out_host = {k: torch.empty(v.shape, dtype=v.dtype).pin_memory() for k, v in flat.items()}
        for k, v in flat.items():
            out_host[k].copy_(v, non_blocking=True)
        torch.cuda.current_stream().synchronize()   # the user reads the depth map on the host (test.py:206-207)

    step_e2e()
    h2d = host["imgs"].numel() * 4 + host["dv"].numel() * 4 + sum(v.numel() * 4 for v in host["proj"].values())
    d2h = sum(v.numel() * v.element_size() for v in out_host.values())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step_e2e()
    e1.record()
    barrier()
    ms_e2e_seq = e0.elapsed_time(e1)

    # (b) the job as it is run: the work list streamed through DepthMapStream (upload of item i+1 and download of item
    # i-1 overlap the kernels of item i; every item's host->device and device->host copies are inside the timed region)
    from cds_mvsnet_b200.streaming import DepthMapStream
    pipe = DepthMapStream(model, temperature=TEMPERATURE)
    for _ in range(len(pipe.slots)):                                   # every slot allocates its staging buffers once
        pipe.result(pipe.submit(host["imgs"], host["proj"], host["dv"]))
    barrier()
    e0.record()
    prev = None
    checksum = 0.0
    last_key = f"stage{len(cfg['ndepths'])}.depth"
    for _ in range(args.steps):
        t = pipe.submit(host["imgs"], host["proj"], host["dv"])
        if prev is not None:
            checksum += float(pipe.result(prev)[last_key][0, 0, 0])   # the host consumes every result
        prev = t
    checksum += float(pipe.result(prev)[last_key][0, 0, 0])
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    ph2d, pd2h = pipe.bytes_per_item()
    assert (ph2d, pd2h) == (h2d, d2h), "streamed path moves different bytes than the direct call"
    clocks = sampler.stop() if rank == 0 else None

    from cds_mvsnet_b200 import parallel
    ms_total = parallel.max_over_ranks(ms_total, dev)   # the slowest rank's interval
    ms_e2e = parallel.max_over_ranks(ms_e2e, dev)
    ms_e2e_seq = parallel.max_over_ranks(ms_e2e_seq, dev)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    maps = world * args.steps * B
    value = maps / (ms_total / 1e3)
    e2e_value = maps / (ms_e2e / 1e3)

    # ---- roofline of the dominant kernel ------------------------------------------------------------
    pk = peaks()
    rows = []
    for (name, tag), d in table.items():
        flops, nbytes = d["meta"] if d["meta"] else (0.0, 0.0)
        ms = d["ms_total"] / d["launches"]
        rows.append(dict(kernel=name, tag=tag, ms_per_launch=ms, launches_per_step=d["launches"] / args.steps,
                         share=d["ms_total"] / max(sum(x["ms_total"] for x in table.values()), 1e-9),
                         alg_gflop=flops / 1e9, alg_mb=nbytes / 1e6,
                         tflops=flops / (ms * 1e-3) / 1e12 if ms > 0 else 0.0, gbs=nbytes / (ms * 1e-3) / 1e9 if ms > 0 else 0.0))
    rows.sort(key=lambda r: -r["share"])
    top = rows[0]
    # a kernel is tensor-bound when its algorithmic intensity is above the ridge point of the measured peaks
    ridge = pk["tensor"] * 1e12 / (pk["hbm"] * 1e9)
    intensity = (top["alg_gflop"] * 1e9) / max(top["alg_mb"] * 1e6, 1.0)
    if intensity > ridge:
        roof = {"bound": "tensor", "achieved": top["tflops"], "peak": pk["tensor"], "unit": "TFLOP/s", "frac": top["tflops"] / pk["tensor"]}
    else:
        roof = {"bound": "hbm", "achieved": top["gbs"], "peak": pk["hbm"], "unit": "GB/s", "frac": top["gbs"] / pk["hbm"]}
    # DRAM traffic of the same kernel from the committed `ncu --set full` capture (dram read + write bytes per launch)
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(f"{top['kernel']}[{top['tag']}]@{args.workload}")
    roof.update({"traffic": traffic, "kernel": f"{top['kernel']}[{top['tag']}]", "share_of_step": top["share"],
                 "ms_per_launch": top["ms_per_launch"], "peak_source": pk["src"],
                 "algorithmic": {"gflop_per_launch": top["alg_gflop"], "mb_per_launch": top["alg_mb"]}})
    if top["kernel"].startswith("cds_dynamic_conv_tc"):
        # what saturates first in the tap-GEMM formulation is the tensor core's shared-memory operand path, not the math pipe
        roof["note"] = ("tap-GEMM DynamicConv: every K=16 MMA re-reads its 4 KB A slab pair and its B columns from shared memory; the "
                        "committed ncu capture (profiles/r01_v12_ncu_conv00_pairs.txt) shows l1tex throughput 84.6 % with the tensor "
                        "pipe 31 % active -- the kernel sits at ~5/6 of the shared-memory operand roofline (DESIGN.md section 5)")
    if args.kernel_table:
        os.makedirs(os.path.dirname(os.path.abspath(args.kernel_table)), exist_ok=True)
        json.dump({"workload": workload_name(args.workload, cfg), "storage": args.storage, "ms_per_step": ms_total / args.steps,
                   "peaks": pk, "kernels": rows}, open(args.kernel_table, "w"), indent=1)
    print("top kernels (share of summed kernel time):", file=sys.stderr)
    for r in rows[:12]:
        print(f"  {r['share'] * 100:5.1f}%  {r['ms_per_launch']:8.3f} ms  {r['tflops']:7.2f} TFLOP/s  {r['gbs']:7.1f} GB/s  {r['kernel']}[{r['tag']}]",
              file=sys.stderr)

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        cores = os.cpu_count() or 1
        rate, _, sample = cpu_oracle_rate(cfg, cores, budget_s=25.0)
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f16" if storage == torch.float16 else "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.workload, cfg), "storage": f"{args.storage} activations, fp32 accumulate",
                   "weights": "pretrained both_dtu_blended (tests/golden/weights_both_dtu_blended.npz)",
                   "parallelism": f"replicas x{world}, work-list sharding, no collective",
                   "launch": "one CUDA graph per forward" if use_graph else "launch by launch",
                   "l2": "working set per step (>1 GB of activations) exceeds the 126 MB L2; no explicit flush",
                   "roofline_region": f"the same {args.steps} steps repeated with a CUDA-event pair around every launch "
                                      f"({ms_instrumented / args.steps:.3f} ms/step instrumented)",
                   "buffers_mb": engine.buf.nbytes() / 1e6},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e / args.steps, "api": "cds_mvsnet_b200.streaming.DepthMapStream (double-buffered copies)",
                "one_call_at_a_time": {"value": maps / (ms_e2e_seq / 1e3), "ms_per_step": ms_e2e_seq / args.steps,
                                       "api": "CDSMVSNet.__call__ on pinned host tensors, blocking per item"}},
        "gpu_launches": launches,
        "roofline": roof,
        "cpu_baseline": cpu,
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
