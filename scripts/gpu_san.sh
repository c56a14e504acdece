#!/bin/bash
# bash scripts/gpu_san.sh <tag> <pytest -k expression>: one test selection under compute-sanitizer memcheck, then plainly
TAG=$1; K=$2
O=gpurun_out; mkdir -p $O
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_dynconv_kh.py -x -q -s -k "$K" > $O/${TAG}_san.log 2>&1
echo "sanitizer rc=$?"; grep -v "^$" $O/${TAG}_san.log | grep -E "=========|kh |passed|failed|Error" | head -60
timeout 600 python -m pytest tests/test_gpu_dynconv_kh.py -q -s -k "150x700 or 200x900" > $O/${TAG}_big.log 2>&1; echo "plain rc=$?"
grep -E "^kh |passed|failed|Error" $O/${TAG}_big.log | tail -20
