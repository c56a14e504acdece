#!/bin/bash
# One GPU-box visit: parity tests, bench with kernel table, ncu launch list, ncu full captures of the top kernels.
# Usage (on the box, from the repo root): bash scripts/gpu_round.sh <tag> [skip_tests]
TAG=${1:-rXX}
O=gpurun_out
mkdir -p $O
if [ -z "$2" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/${TAG}_pytest.log
  tail -3 $O/${TAG}_pytest.log
fi
timeout 600 python bench.py --steps 20 --warmup 3 --kernel-table $O/${TAG}_kernel_table_cfg2.json > $O/${TAG}_bench_cfg2.json 2> $O/${TAG}_bench.err
cat $O/${TAG}_bench_cfg2.json; tail -14 $O/${TAG}_bench.err
# every launch of one warm forward (the second of two)
OURS='regex:(dynconv|conv3d|deconv3d|entropy|aggregate|visnet|conv1x1|conv3x3|conv2d|instnorm|softmax_regress|hypotheses|nc_mean|camera_setup|image_to|prob_conv|homo_warp|warp_coeffs)'
N=${NCU_LIST_COUNT:-69}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$OURS" -s $N -c $N --csv --log-file $O/${TAG}_launches.csv \
    python scripts/run_forward.py --iters 2 > $O/${TAG}_ncu_list.log 2>&1
# --set full of the kernels named in NCU_KERNELS as "<regex>:<skip>" pairs
for KS in ${NCU_KERNELS:-dynconv_tc_kernel:0 aggregate_f16_kernel:1}; do
  K=${KS%%:*}; S=${KS##*:}
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s $S -c 1 -f -o $O/${TAG}_${K}_$S \
      python scripts/run_forward.py --iters 1 > $O/${TAG}_ncu_${K}_$S.log 2>&1
done
ls -la $O | tail -12
