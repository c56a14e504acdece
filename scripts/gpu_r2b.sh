#!/bin/bash
# bash scripts/gpu_r2b.sh <tag> : ops + parity tests (no -x), error budgets (small multi-seed, full-size), bench
TAG=${1:-r02}
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -rs --durations=5 ${PYTEST_ARGS} > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/${TAG}_pytest.log
grep -E "passed|failed|FAILED|Error" $O/${TAG}_pytest.log | tail -30
if [ -n "$ERR_SMALL" ]; then timeout 600 python scripts/err_budget4.py --seeds=0-7 $ERR_SMALL > $O/${TAG}_err_small.txt 2>&1; tail -4 $O/${TAG}_err_small.txt; fi
if [ -n "$ERR_FULL" ]; then timeout 900 python scripts/err_budget4.py --hw=1184x1600 --n=5 --seeds=0-2 $ERR_FULL > $O/${TAG}_err_full.txt 2>&1; tail -6 $O/${TAG}_err_full.txt; fi
timeout 900 python bench.py --steps 20 --warmup 5 --no-incumbent --kernel-table $O/${TAG}_kernel_table_cfg2.json > $O/${TAG}_bench_cfg2.json 2> $O/${TAG}_bench.err
python - <<PY
import json
d=json.load(open("$O/${TAG}_bench_cfg2.json"))
print("value",d["value"],"ms",d["ms_per_step"],"e2e",d["e2e"]["value"],"seq",d["e2e"]["one_call_at_a_time"]["value"])
print("parity",json.dumps(d["parity"]["stages"]) if d.get("parity") else None)
PY
tail -24 $O/${TAG}_bench.err | head -20
