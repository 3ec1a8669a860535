mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_tc_gpu.py -q -x -k "kl" > gpurun_out/t_tc.log 2>&1; echo "tc tests rc=$?"; tail -3 gpurun_out/t_tc.log
grep '"tag": "kl' gpurun_out/tc_diag.jsonl | tail -8 | cut -c1-120
timeout 200 python tools/prof_tc.py --m 65536 --n 65536 --k 32 --reps 3 --kl 2>&1 | tail -2
if [ -n "$BENCH" ]; then
timeout 400 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/bench_quick.log 2>&1; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_quick.log') if l.startswith('{')][-1])
print('value', d['value'], 'by_norm', {k:round(v['ms_per_step'],3) for k,v in d['by_norm'].items()})
print({k:(round(v['mean_ms'],3), round(v['frac'],3)) for k,v in d['roofline']['per_kernel'].items()}, d['clocks'])
PY
fi
