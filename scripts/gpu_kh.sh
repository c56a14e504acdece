#!/bin/bash
# bash scripts/gpu_kh.sh <tag>: the kh kernel's own tests first; then the suite, error budgets and bench (with CDS_DYN_KH=0 if they failed)
TAG=${1:-r02}
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_dynconv_kh.py -q -x -s > $O/${TAG}_kh_pytest.log 2>&1; KH=$?
echo "kh pytest rc=$KH" | tee -a $O/${TAG}_kh_pytest.log
grep -E "^kh |passed|failed|Error|error" $O/${TAG}_kh_pytest.log | tail -40
if [ $KH -ne 0 ]; then export CDS_DYN_KH=0; echo "FALLING BACK TO CDS_DYN_KH=0"; fi
timeout 1500 python -m pytest tests -m gpu -q -rs --deselect tests/test_gpu_dynconv_kh.py > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/${TAG}_pytest.log
grep -E "passed|failed|FAILED|Error" $O/${TAG}_pytest.log | tail -20
timeout 600 python scripts/err_budget4.py --seeds=0-5 > $O/${TAG}_err_small.txt 2>&1; tail -2 $O/${TAG}_err_small.txt
timeout 900 python bench.py --steps 20 --warmup 5 --no-incumbent --kernel-table $O/${TAG}_kernel_table_cfg2.json > $O/${TAG}_bench_cfg2.json 2> $O/${TAG}_bench.err
python - <<PY
import json
d=json.load(open("$O/${TAG}_bench_cfg2.json"))
print("value",d["value"],"ms",d["ms_per_step"],"e2e",d["e2e"]["value"],"seq",d["e2e"]["one_call_at_a_time"]["value"])
print("parity",json.dumps(d["parity"]["stages"]) if d.get("parity") else None)
PY
tail -28 $O/${TAG}_bench.err | head -24
