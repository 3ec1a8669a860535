# Round-2 evidence on ONE B200: GPU test-suite, smoke, default bench line (+ the CPU reference arm), launch list of a short
# bench run, ncu --set full capture of the A-streaming kernels at the benchmark shape, cfg3 / cfg4 through bench.py, the
# tiny configurations.  Everything lands in gpurun_out/ (scratch); the judged summaries are copied to profiles/ afterwards.
mkdir -p gpurun_out
SECONDS=0
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r02_t_all_gpu.log 2>&1; echo "gpu tests rc=$? in ${SECONDS}s"; tail -3 gpurun_out/r02_t_all_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r02_smoke.log
timeout 900 python bench.py > gpurun_out/r02_bench_1gpu.log 2>&1; echo "bench rc=$? at ${SECONDS}s"; tail -c 4500 gpurun_out/r02_bench_1gpu.log
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'tc_pass_kernel|tc_kl_kernel' -c 8 -f -o gpurun_out/r02_full_65536 python tools/prof_tc.py --m 65536 --n 65536 --k 32 --reps 1 --kl > gpurun_out/r02_ncu_full.log 2>&1; echo "ncu rc=$? at ${SECONDS}s"
ncu -i gpurun_out/r02_full_65536.ncu-rep --page raw --csv > gpurun_out/r02_full_65536_raw.csv 2>/dev/null
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_bench_steps2.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r02_bench_under_ncu.log 2>&1; echo "launchlist rc=$? at ${SECONDS}s"
timeout 600 python bench.py --config cfg4 --steps 10 --warmup 3 --no-e2e > gpurun_out/r02_bench_cfg4_1gpu.log 2>&1; echo "cfg4 rc=$? at ${SECONDS}s"; tail -c 1500 gpurun_out/r02_bench_cfg4_1gpu.log
timeout 600 python bench.py --config cfg3 --steps 10 --warmup 3 --no-e2e > gpurun_out/r02_bench_cfg3_1gpu.log 2>&1; echo "cfg3 rc=$? at ${SECONDS}s"; tail -c 1200 gpurun_out/r02_bench_cfg3_1gpu.log
timeout 600 python tools/bench_configs.py --only cfg1,cfg5 > gpurun_out/r02_bench_configs_1gpu.log 2>&1; echo "configs rc=$? at ${SECONDS}s"; grep "^{" gpurun_out/r02_bench_configs_1gpu.log | cut -c1-500
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_arm.log 2>&1; echo "reference arm rc=$? at ${SECONDS}s"; tail -c 1500 gpurun_out/r02_bench_reference_arm.log
