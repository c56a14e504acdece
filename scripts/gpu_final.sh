#!/bin/bash
# last visit of a change: whole GPU suite, smoke, default bench (+ kernel table), launch list, one ncu --set full of kernel $1
TAG=${TAG:-r02_v26}; K=${1:-aggregate_f32q}; S=${2:-0}
O=gpurun_out; mkdir -p $O/$TAG
timeout 1800 python -m pytest tests -m gpu -q -rs --durations=5 > $O/$TAG/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/$TAG/pytest.log
grep -E "passed|failed|FAILED|Error" $O/$TAG/pytest.log | tail -8
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/$TAG/smoke.log 2>&1; tail -2 $O/$TAG/smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 --kernel-table $O/$TAG/kernel_table_cfg2.json > $O/$TAG/bench_cfg2.json 2> $O/$TAG/bench.err
timeout 600 python bench.py > $O/$TAG/bench_default_flags.json 2> $O/$TAG/bench_default_flags.err
python - <<PY
import json
for f in ("bench_cfg2", "bench_default_flags"):
    d=json.load(open("$O/$TAG/%s.json" % f))
    print(f, "value",d["value"],"ms",d["ms_per_step"],"single",d["one_map_at_a_time"]["value"],"e2e",d["e2e"]["value"],"seq",d["e2e"]["one_call_at_a_time"]["value"])
    print(" parity",json.dumps(d["parity"]["stages"]) if d.get("parity") else None)
    print(" incumbent",json.dumps(d.get("reference_eager_gpu")), "launches", d["gpu_launches"], "clocks", d["clocks"])
    print(" roofline", json.dumps(d["roofline"])[:300])
PY
OURS='regex:(dynconv|conv3d|deconv3d|entropy|aggregate|visnet|conv1x1|conv3x3|conv2d|instnorm|softmax_regress|regress|hypotheses|nc_mean|camera_setup|image_to|u8_to|prob_conv|homo_warp|warp_coeffs|costvol|s2rows)'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$OURS" -s 69 -c 69 --csv --log-file $O/$TAG/launches_raw.csv \
    python scripts/run_forward.py --iters 2 > $O/$TAG/ncu_list.log 2>&1
python scripts/ncu_summary.py launches $O/$TAG/launches_raw.csv $O/$TAG/launches.csv; rm -f $O/$TAG/launches_raw.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s $S -c 1 -f -o /tmp/rep_$K \
    python scripts/run_forward.py --iters 1 > $O/$TAG/ncu_$K.log 2>&1
python scripts/ncu_summary.py full /tmp/rep_$K.ncu-rep $O/$TAG/ncu_$K.txt; cat $O/$TAG/ncu_$K.txt | head -30
