mkdir -p gpurun_out
if [ -n "$QUICK" ]; then
  timeout 300 python -m pytest tests/test_kernels_gpu.py tests/test_parity_gpu.py -q -k "resident" > gpurun_out/t_res.log 2>&1; echo "resident rc=$?"; grep -v "^$" gpurun_out/t_res.log | grep -v "^E   " | tail -15
else
  SECONDS=0
  timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/t_all_gpu.log 2>&1; echo "gpu tests rc=$? in ${SECONDS}s"; grep -v "^$" gpurun_out/t_all_gpu.log | grep -v "^E   " | tail -15
fi
timeout 600 python tools/bench_configs.py --only ${CONFIGS:-cfg1,cfg5} > gpurun_out/bench_configs.log 2>&1; echo "configs rc=$?"; cat gpurun_out/bench_configs.log | tail -5
