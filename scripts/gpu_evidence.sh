#!/bin/bash
# bash scripts/gpu_evidence.sh <tag>: the evidence set of a round -- GPU suite, smoke, bench (all legs), ncu launch list, ncu --set full
# captures of the kernels named in NCU_KERNELS ("regex:skip" pairs)
TAG=${1:-r02}
O=gpurun_out
mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q -rs --durations=5 > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/${TAG}_pytest.log
grep -E "passed|failed|FAILED|Error" $O/${TAG}_pytest.log | tail -8
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; tail -2 $O/${TAG}_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 --kernel-table $O/${TAG}_kernel_table_cfg2.json > $O/${TAG}_bench_cfg2.json 2> $O/${TAG}_bench.err
python - <<PY
import json
d=json.load(open("$O/${TAG}_bench_cfg2.json"))
print("value",d["value"],"ms",d["ms_per_step"],"single",d.get("one_map_at_a_time"),"e2e",d["e2e"]["value"],"seq",d["e2e"]["one_call_at_a_time"]["value"])
print("parity",json.dumps(d["parity"]["stages"]) if d.get("parity") else None)
print("incumbent",json.dumps(d.get("reference_eager_gpu")))
print("cpu",json.dumps(d.get("cpu_baseline")))
PY
tail -32 $O/${TAG}_bench.err | head -26
if [ -n "$EXTRA_BENCH" ]; then
  timeout 600 python bench.py --steps 20 --warmup 5 --no-incumbent --no-cpu-baseline $EXTRA_BENCH > $O/${TAG}_bench_extra.json 2> $O/${TAG}_bench_extra.err
  python -c "import json; d=json.load(open('$O/${TAG}_bench_extra.json')); print('extra [$EXTRA_BENCH]: value', d['value'], 'e2e', d['e2e']['value'])"
fi
OURS='regex:(dynconv|conv3d|deconv3d|entropy|aggregate|visnet|conv1x1|conv3x3|conv2d|instnorm|softmax_regress|regress|hypotheses|nc_mean|camera_setup|image_to|u8_to|prob_conv|homo_warp|warp_coeffs|costvol)'
N=${NCU_LIST_COUNT:-69}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$OURS" -s $N -c $N --csv --log-file $O/${TAG}_launches.csv \
    python scripts/run_forward.py --iters 2 > $O/${TAG}_ncu_list.log 2>&1
for KS in $NCU_KERNELS; do
  K=${KS%%:*}; S=${KS##*:}
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s $S -c 1 -f -o $O/${TAG}_${K}_$S \
      python scripts/run_forward.py --iters 1 > $O/${TAG}_ncu_${K}_$S.log 2>&1
done
ls -la $O | tail -12
