#!/bin/bash
TAG=${1:-r02}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_ops.py -q -s -k "s2_rows or fused_tail or 3x3s2" > $O/${TAG}_s2_pytest.log 2>&1; echo "s2 pytest rc=$?" | tee -a $O/${TAG}_s2_pytest.log
grep -E "^3x3 s2|passed|failed|Error|error" $O/${TAG}_s2_pytest.log | tail -40
timeout 900 python -m pytest tests/test_gpu_dynconv_kh.py -q -s -k "heads or rejects" > $O/${TAG}_kh_pytest.log 2>&1; echo "kh pytest rc=$?" | tee -a $O/${TAG}_kh_pytest.log
grep -E "^kh |passed|failed|Error|error" $O/${TAG}_kh_pytest.log | tail -20
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_e2e.py -q -rs -x > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/${TAG}_pytest.log
grep -E "passed|failed|FAILED|Error" $O/${TAG}_pytest.log | tail -20
timeout 900 python bench.py --steps 20 --warmup 5 --no-incumbent --kernel-table $O/${TAG}_kernel_table_cfg2.json > $O/${TAG}_bench_cfg2.json 2> $O/${TAG}_bench.err
python - <<PY
import json
d=json.load(open("$O/${TAG}_bench_cfg2.json"))
print("value",d["value"],"ms",d["ms_per_step"],"e2e",d["e2e"]["value"],"seq",d["e2e"]["one_call_at_a_time"]["value"])
print("parity",json.dumps(d["parity"]["stages"]) if d.get("parity") else None)
for k in json.load(open("$O/${TAG}_kernel_table_cfg2.json"))["kernels"]:
    if k["tag"].startswith("feat."): print(f"  {k['ms_per_launch']:.3f} {k['kernel']}[{k['tag']}]")
PY
tail -8 $O/${TAG}_bench.err | head -6
