#!/bin/bash
# A/B of the stage-1 cost-volume sweeps on fp32 features: four channels per lane (CDS_AGG_QUAD = resident blocks 2/3/4,
# CDS_ENT_QUAD=1) against eight (0)
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_ops.py -q -x -k "costvol or aggregate or entropy" 2>&1 | tail -3
for q in ${QS:-00 31}; do
  CDS_AGG_QUAD=${q:0:1} CDS_ENT_QUAD=${q:1:1} timeout 600 python bench.py --steps 20 --warmup 5 --no-incumbent ${BARGS:---no-cpu-baseline} --kernel-table $O/agg_q${q}_table.json > $O/agg_q${q}.json 2> $O/agg_q${q}.err
  python - <<PY
import json
d=json.load(open("$O/agg_q${q}.json")); t=json.load(open("$O/agg_q${q}_table.json"))
k=[r for r in t["kernels"] if "s0.costvol" in r["tag"]]
print("agg/ent quad=$q value",d["value"],"single",d["one_map_at_a_time"]["value"],"ms",d["ms_per_step"],"parity",d["parity"]["stages"] if d.get("parity") else None)
for r in k: print("   ",r["tag"],r["kernel"][:40],r["ms_per_launch"])
PY
done
