# Round-end evidence on one B200: GPU test-suite, smoke, ncu --set full capture of the A-streaming kernels at the
# benchmark shape, launch list of a short bench run, secondary configurations, and the default bench run.
mkdir -p gpurun_out
SECONDS=0
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/t_all_gpu.log 2>&1; echo "gpu tests rc=$? in ${SECONDS}s"; tail -3 gpurun_out/t_all_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'tc_pass_kernel|tc_kl_kernel' -c 8 -f -o gpurun_out/r01_full_65536 python tools/prof_tc.py --m 65536 --n 65536 --k 32 --reps 1 --kl > gpurun_out/ncu_full.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/r01_full_65536.ncu-rep --page raw --csv > gpurun_out/r01_full_65536_raw.csv 2>/dev/null
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r01_launches_bench_final.csv python bench.py --steps 2 --warmup 1 --no-e2e > gpurun_out/bench_under_ncu.log 2>&1; echo "launchlist rc=$?"
timeout 600 python tools/bench_configs.py --only cfg1,cfg4,cfg5 > gpurun_out/bench_configs.log 2>&1; echo "configs rc=$?"; grep "^{" gpurun_out/bench_configs.log | cut -c1-400
timeout 600 python bench.py > gpurun_out/bench_final.log 2>&1; echo "bench rc=$?"; tail -c 3500 gpurun_out/bench_final.log
