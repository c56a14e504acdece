#!/bin/bash
# One GPU-box visit of round 2: bash scripts/gpu_r2.sh <tag> [tests|notests] [errbudget args...]
TAG=${1:-r02}; TESTS=${2:-tests}; shift; shift
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > $O/${TAG}_gpu.txt 2>&1
if [ "$TESTS" = "tests" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q -rs --durations=8 > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/${TAG}_pytest.log
  tail -25 $O/${TAG}_pytest.log
fi
if [ -n "$ERRBUDGET" ]; then
  timeout 900 python scripts/err_budget4.py $ERRBUDGET > $O/${TAG}_err_budget.txt 2>&1; tail -12 $O/${TAG}_err_budget.txt
fi
timeout 900 python bench.py --steps 20 --warmup 5 --kernel-table $O/${TAG}_kernel_table_cfg2.json > $O/${TAG}_bench_cfg2.json 2> $O/${TAG}_bench.err
cat $O/${TAG}_bench_cfg2.json; tail -28 $O/${TAG}_bench.err
if [ -n "$NCU_LIST" ]; then
  OURS='regex:(dynconv|conv3d|deconv3d|entropy|aggregate|visnet|conv1x1|conv3x3|conv2d|instnorm|softmax_regress|regress|hypotheses|nc_mean|camera_setup|image_to|u8_to|prob_conv|homo_warp|warp_coeffs|costvol)'
  N=${NCU_LIST_COUNT:-69}
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$OURS" -s $N -c $N --csv --log-file $O/${TAG}_launches.csv \
      python scripts/run_forward.py --iters 2 > $O/${TAG}_ncu_list.log 2>&1
fi
for KS in $NCU_KERNELS; do
  K=${KS%%:*}; S=${KS##*:}
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s $S -c 1 -f -o $O/${TAG}_${K}_$S \
      python scripts/run_forward.py --iters 1 > $O/${TAG}_ncu_${K}_$S.log 2>&1
done
ls -la $O | tail -8
