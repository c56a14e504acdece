#!/bin/bash
TAG=${1:-r02}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_parity.py -q -rs -x > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/${TAG}_pytest.log
grep -E "passed|failed|FAILED|Error" $O/${TAG}_pytest.log | tail -20
timeout 900 python bench.py --steps 20 --warmup 5 --no-incumbent ${BENCH_ARGS} --kernel-table $O/${TAG}_kernel_table_cfg2.json > $O/${TAG}_bench_cfg2.json 2> $O/${TAG}_bench.err
python - <<PY
import json
d=json.load(open("$O/${TAG}_bench_cfg2.json"))
print("value",d["value"],"ms",d["ms_per_step"],"single",d["one_map_at_a_time"],"e2e",d["e2e"]["value"],"seq",d["e2e"]["one_call_at_a_time"]["value"])
print("parity",json.dumps(d["parity"]["stages"]) if d.get("parity") else None)
print("launches",d["gpu_launches"],"clocks",d["clocks"])
PY
tail -8 $O/${TAG}_bench.err | head -6
