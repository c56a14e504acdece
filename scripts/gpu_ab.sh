#!/bin/bash
# A/B visit: parity tests once, then the bench + kernel table once per environment setting.
# bash scripts/gpu_ab.sh <tag> <VAR=value> [<VAR=value> ...]     (PYTEST_K filters the tests; GREP picks the table rows shown)
TAG=$1; shift 1
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q -s ${PYTEST_K:+-k "$PYTEST_K"} > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; grep -h "float16: depth rel-L1" $O/${TAG}_pytest.log
tail -8 $O/${TAG}_pytest.log
for KV in "$@"; do
  echo "== $KV"
  N=${KV//=/_}
  env $KV timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --kernel-table $O/${TAG}_${N}_kernel_table_cfg2.json > $O/${TAG}_${N}_bench_cfg2.json 2> $O/${TAG}_${N}_bench.err
  cut -c1-200 $O/${TAG}_${N}_bench_cfg2.json; grep -E "${GREP:-costvol}" $O/${TAG}_${N}_bench.err
done
