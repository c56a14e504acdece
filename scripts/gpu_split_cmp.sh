#!/bin/bash
# bench with and without split-precision conv10/conv11 inputs
for S in 0 1; do
  CDS_SPLIT=$S timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --kernel-table gpurun_out/split${S}_table.json > gpurun_out/split${S}_bench.json 2> gpurun_out/split${S}.err
  cut -c1-200 gpurun_out/split${S}_bench.json
done
