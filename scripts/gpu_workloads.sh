#!/bin/bash
# One GPU-box visit: the training-slice tests and the other BASELINE workloads' bench lines.  bash scripts/gpu_workloads.sh <tag>
TAG=${1:-rXX}
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_train.py -x -q > $O/${TAG}_pytest_train.log 2>&1; echo "pytest rc=$?"; tail -3 $O/${TAG}_pytest_train.log
for W in cfg1 cfg3 cfg4 cfg5 dtu_refine; do
  timeout 300 python bench.py --workload $W --steps 10 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench_$W.json 2> $O/${TAG}_bench_$W.err
  cut -c1-160 $O/${TAG}_bench_$W.json
done
timeout 200 python scripts/bench_train_ops.py > $O/${TAG}_train_ops_bench.json 2>/dev/null
