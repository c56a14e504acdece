mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "train or cascade or feature_net or full_size or stream or graph" > gpurun_out/v11a_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/v11a_pytest.log
for S in 8 6; do
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dynconv_tc_kernel -s $S -c 1 -f -o gpurun_out/v11a_dynconv_$S python scripts/run_forward.py --iters 1 > gpurun_out/v11a_ncu_$S.log 2>&1
done
ls -la gpurun_out | tail -5
