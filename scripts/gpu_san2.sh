#!/bin/bash
TAG=$1; O=gpurun_out; mkdir -p $O
timeout 900 compute-sanitizer --tool memcheck --print-limit 30 python scripts/run_forward.py --workload cfg2 --iters 1 > $O/${TAG}_san_cfg2.log 2>&1
echo "sanitizer rc=$?"; grep -E "=========" $O/${TAG}_san_cfg2.log | head -70
timeout 300 python scripts/run_forward.py --workload cfg4 --iters 1 > $O/${TAG}_cfg4.log 2>&1; echo "cfg4 plain rc=$?"; tail -3 $O/${TAG}_cfg4.log
