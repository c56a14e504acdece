#!/bin/bash
# One GPU-box visit: the other BASELINE workloads' bench lines (parity + cpu baseline legs included where they fit).  bash scripts/gpu_workloads.sh <tag>
TAG=${1:-rXX}
O=gpurun_out
mkdir -p $O
for W in cfg1 cfg3 cfg4 cfg5 dtu_refine; do
  timeout 600 python bench.py --workload $W --steps 10 --warmup 3 --no-incumbent > $O/${TAG}_bench_$W.json 2> $O/${TAG}_bench_$W.err
  python - <<PY
import json
try:
    d=json.load(open("$O/${TAG}_bench_$W.json"))
    par=d.get("parity") or {}
    print("$W", round(d["value"],2), "maps/s", round(d["ms_per_step"],3), "ms  e2e", round(d["e2e"]["value"],2), " parity", {k: round(v["depth_rel_l1"],6) for k,v in (par.get("stages") or {}).items()}, par.get("refined_depth_rel_l1"))
except Exception as e:
    print("$W FAILED", e)
PY
done
