#!/bin/bash
# Quick GPU visit: parity tests (optional -k filter in $PYTEST_K) + bench with kernel table.  bash scripts/gpu_quick.sh <tag>
TAG=${1:-quick}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q ${PYTEST_K:+-k "$PYTEST_K"} > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
tail -15 $O/${TAG}_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --kernel-table $O/${TAG}_kernel_table_cfg2.json > $O/${TAG}_bench_cfg2.json 2> $O/${TAG}_bench.err
cat $O/${TAG}_bench_cfg2.json | cut -c1-330; tail -25 $O/${TAG}_bench.err
