# Round-2 multi-GPU evidence on ONE box with N GPUs:  bash tools/multi_gpu_r02.sh N [quick]
#   NCCL parity vs the reference goldens, cfg2 / cfg3 / cfg4 through bench.py, cfg5 (NMFk as specified) through tools/bench_cfg5.py.
N=${1:-2}; QUICK=${2:-}
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
mkdir -p gpurun_out
SECONDS=0
run() { # name, timeout, command...
  local name=$1 to=$2; shift 2
  timeout $to "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$? at ${SECONDS}s"; grep "^{" gpurun_out/$name.log | tail -1 | cut -c1-900; grep -i "nccl parity:\|FAIL\|Error\|error:" gpurun_out/$name.log | head -5
}
STEPS=${STEPS:-20}
run r02_nccl_parity_${N}gpu 500 $TR --nproc-per-node $N --master-port 29511 tools/nccl_parity.py
run r02_bench_cfg2_${N}gpu 700 $TR --nproc-per-node $N --master-port 29512 bench.py --gpus $N --steps 30 --warmup 5
run r02_bench_cfg3_${N}gpu 500 $TR --nproc-per-node $N --master-port 29513 bench.py --gpus $N --config cfg3 --steps $STEPS --warmup 3 --no-e2e
run r02_bench_cfg4_${N}gpu 500 $TR --nproc-per-node $N --master-port 29514 bench.py --gpus $N --config cfg4 --steps $STEPS --warmup 3 --no-e2e
if [ -n "$QUICK" ]; then
  run r02_cfg5_quick_${N}gpu 300 $TR --nproc-per-node $N --master-port 29515 tools/bench_cfg5.py --quick
fi
run r02_cfg5_${N}gpu 600 $TR --nproc-per-node $N --master-port 29516 tools/bench_cfg5.py
if [ "$N" = "8" ]; then
  run r02_bench_cfg4_4gpu 400 $TR --nproc-per-node 4 --master-port 29518 bench.py --gpus 4 --config cfg4 --steps $STEPS --warmup 3 --no-e2e
  run r02_bench_cfg3_4gpu 500 $TR --nproc-per-node 4 --master-port 29520 bench.py --gpus 4 --config cfg3 --steps $STEPS --warmup 3 --no-e2e
fi
echo "total ${SECONDS}s"
