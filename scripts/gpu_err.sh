#!/bin/bash
# error budget only: bash scripts/gpu_err.sh <tag> <err_budget4 args...>
TAG=$1; shift
mkdir -p gpurun_out
timeout 1500 python scripts/err_budget4.py "$@" > gpurun_out/${TAG}_err_budget.txt 2>&1; tail -20 gpurun_out/${TAG}_err_budget.txt
