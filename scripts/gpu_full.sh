#!/bin/bash
# bash scripts/gpu_full.sh <tag>: whole GPU suite, error budgets (small multi-seed, full size), bench
TAG=${1:-r02}
O=gpurun_out
mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q -rs --durations=5 > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/${TAG}_pytest.log
grep -E "passed|failed|FAILED|Error" $O/${TAG}_pytest.log | tail -20
timeout 600 python scripts/err_budget4.py --seeds=0-7 > $O/${TAG}_err_small.txt 2>&1; tail -2 $O/${TAG}_err_small.txt
timeout 900 python scripts/err_budget4.py --hw=1184x1600 --n=5 --seeds=0-2 > $O/${TAG}_err_full.txt 2>&1; tail -2 $O/${TAG}_err_full.txt
timeout 900 python bench.py --steps 20 --warmup 5 ${BENCH_ARGS} --kernel-table $O/${TAG}_kernel_table_cfg2.json > $O/${TAG}_bench_cfg2.json 2> $O/${TAG}_bench.err
python - <<PY
import json
d=json.load(open("$O/${TAG}_bench_cfg2.json"))
print("value",d["value"],"ms",d["ms_per_step"],"e2e",d["e2e"]["value"],"seq",d["e2e"]["one_call_at_a_time"]["value"])
print("parity",json.dumps(d["parity"]["stages"]) if d.get("parity") else None)
PY
tail -30 $O/${TAG}_bench.err | head -26
