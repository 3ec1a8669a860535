mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_parity_gpu.py -x -q -k "ensemble or oracle_live or numpy_like" > gpurun_out/t_ens.log 2>&1; echo "ens rc=$?"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'tc_pass_kernel|tc_kl_kernel' -c 8 -f -o gpurun_out/r01_full_65536 python tools/prof_tc.py --m 65536 --n 65536 --k 32 --reps 1 --kl > gpurun_out/ncu_full.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/r01_full_65536.ncu-rep --page raw --csv > gpurun_out/r01_full_65536_raw.csv 2>/dev/null
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r01_launches_bench_final.csv python bench.py --steps 2 --warmup 1 --no-e2e > gpurun_out/bench_under_ncu.log 2>&1; echo "launchlist rc=$?"
timeout 600 python bench.py > gpurun_out/bench_final.log 2>&1; echo "bench rc=$?"; tail -c 3000 gpurun_out/bench_final.log
tail -3 gpurun_out/t_ens.log
