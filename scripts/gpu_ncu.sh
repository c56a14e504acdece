#!/bin/bash
# ncu --set full capture of ONE launch: bash scripts/gpu_ncu.sh <tag> <kernel regex> <skip>
TAG=$1; K=$2; S=${3:-0}
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s $S -c 1 -f -o gpurun_out/${TAG} \
    python scripts/run_forward.py --iters 1 > gpurun_out/${TAG}.log 2>&1
tail -3 gpurun_out/${TAG}.log
