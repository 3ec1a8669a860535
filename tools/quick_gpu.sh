# quick validation on one B200: tcgen05 kernel tests, end-to-end tensor-path parity, short bench lines
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_tc_gpu.py tests/test_kernels_gpu.py -q -x > gpurun_out/q_tc.log 2>&1; echo "tc+kernels rc=$?"; tail -2 gpurun_out/q_tc.log
timeout 900 python -m pytest tests/test_parity_gpu.py -q -x > gpurun_out/q_par.log 2>&1; echo "parity rc=$?"; tail -2 gpurun_out/q_par.log
timeout 400 python bench.py --config cfg3 --steps 15 --warmup 3 --no-e2e > gpurun_out/q_bench_cfg3.log 2>&1; echo "bench cfg3 rc=$?"; tail -c 1500 gpurun_out/q_bench_cfg3.log
