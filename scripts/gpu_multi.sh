N=$1
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r01_v11_bench_cfg2_${N}gpu.json 2> gpurun_out/r01_v11_bench_cfg2_${N}gpu.err
cut -c1-220 gpurun_out/r01_v11_bench_cfg2_${N}gpu.json; tail -3 gpurun_out/r01_v11_bench_cfg2_${N}gpu.err
