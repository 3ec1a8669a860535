# quick validation on one B200: tcgen05 kernel tests, end-to-end tensor-path parity, short bench lines
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_tc_gpu.py tests/test_kernels_gpu.py -q -x > gpurun_out/q_tc.log 2>&1; echo "tc+kernels rc=$?"; tail -2 gpurun_out/q_tc.log
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_peer_gpu.py -q -x > gpurun_out/q_par.log 2>&1; echo "parity rc=$?"; tail -2 gpurun_out/q_par.log
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/q_bench.log 2>&1; echo "bench rc=$?"; tail -c 2600 gpurun_out/q_bench.log
