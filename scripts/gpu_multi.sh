#!/bin/bash
# bash scripts/gpu_multi.sh <N> <tag>: both bench arms under torchrun exactly as the driver launches them
N=$1; TAG=${2:-rXX}
O=gpurun_out; mkdir -p $O
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 20 --warmup 5 > $O/${TAG}_bench_cfg2_${N}gpu.json 2> $O/${TAG}_bench_cfg2_${N}gpu.err
cut -c1-260 $O/${TAG}_bench_cfg2_${N}gpu.json; tail -3 $O/${TAG}_bench_cfg2_${N}gpu.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > $O/${TAG}_bench_ref_${N}gpu.json 2> $O/${TAG}_bench_ref_${N}gpu.err
cut -c1-400 $O/${TAG}_bench_ref_${N}gpu.json; tail -2 $O/${TAG}_bench_ref_${N}gpu.err
