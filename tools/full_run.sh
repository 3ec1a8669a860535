mkdir -p gpurun_out
SECONDS=0
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/t_all_gpu.log 2>&1; echo "gpu tests rc=$? in ${SECONDS}s"; grep -v "^$" gpurun_out/t_all_gpu.log | grep -v "^E   " | tail -25
timeout 600 python tools/bench_configs.py --only ${CONFIGS:-cfg1,cfg5} > gpurun_out/bench_configs.log 2>&1; echo "configs rc=$?"; cat gpurun_out/bench_configs.log | tail -5
for cfg in "--m 1024 --n 256 --k 4 --norm kl" "--m 1024 --n 256 --k 4 --norm fro"; do
  timeout 120 python tools/tiny_profile.py $cfg 2>&1 | tail -2
done
