#!/bin/bash
TAG=${1:-r02}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_ops.py -q -k "visnet" > $O/${TAG}_vis_pytest.log 2>&1; echo "vis pytest rc=$?" | tee -a $O/${TAG}_vis_pytest.log
grep -E "passed|failed|Error|error" $O/${TAG}_vis_pytest.log | tail -5
bash scripts/gpu_e2e.sh $TAG
python - <<PY
import json
for k in json.load(open("$O/${TAG}_kernel_table_cfg2.json"))["kernels"]:
    if "visnet" in k["tag"]: print(f"  {k['ms_per_launch']:.3f} {k['kernel']}[{k['tag']}]")
PY
