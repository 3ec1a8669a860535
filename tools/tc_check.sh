mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_tc_gpu.py -q > gpurun_out/t_tc.log 2>&1; echo "tc tests rc=$?"; grep -v "^$" gpurun_out/t_tc.log | grep -v "^E  " | tail -8
timeout 300 python -m pytest tests/test_parity_gpu.py -q -x -k "generic_and_tensor or single_rank" > gpurun_out/t_par.log 2>&1; echo "parity rc=$?"; tail -3 gpurun_out/t_par.log
timeout 200 python tools/prof_tc.py --m 65536 --n 65536 --k 10 --reps 2 --kl 2>&1 | tail -2
