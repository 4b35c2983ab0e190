#!/bin/bash
# 2-GPU (or N-GPU) data-parallel bench.  usage: gpurun --gpus N -- 'bash scripts/gpu_ddp.sh <tag> <N>'
TAG=$1; N=${2:-2}
mkdir -p gpurun_out
run() {  # name, env...
  local name=$1; shift
  env HG_BENCH_LIGHT=1 "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
      --master-port 29511 bench.py --gpus $N --steps 16 --warmup 3 > gpurun_out/bench_train_${N}gpu_${TAG}_$name.json 2> gpurun_out/bench_train_${N}gpu_${TAG}_$name.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_train_${N}gpu_${TAG}_$name.json'))
    print('$name', {k: d[k] for k in ('value','ms_per_step','n_gpus')}, d['config'].get('step_ms'))
except Exception as e:
    print('$name FAILED', e); print(open('gpurun_out/bench_train_${N}gpu_${TAG}_$name.err').read()[-1500:])
PY
}
MODES=${3:-"pipelined4 pipelined8 unpipelined"}
for m in $MODES; do
  case $m in
    pipelined4) run pipelined4 HG_EXCHANGE_CHUNKS=4 ;;
    pipelined8) run pipelined8 HG_EXCHANGE_CHUNKS=8 ;;
    unpipelined) run unpipelined HG_EXCHANGE_CHUNKS=1 ;;
    nosplit) run nosplit HG_SPLIT_G=0 ;;
    legacy) run legacy HG_GRAD_ARENA=0 HG_SPLIT_G=0 HG_EXCHANGE_CHUNKS=1 ;;
  esac
done
# the same window on ONE GPU of the same box, for the efficiency denominator (skipped with a 4th argument "no1":
# under gpurun --gpus 8 a single-GPU minute is charged eight times)
[ "$4" = "no1" ] && exit 0
HG_BENCH_LIGHT=1 timeout 600 python bench.py --gpus 1 --steps 16 --warmup 3 > gpurun_out/bench_train_1gpu_${TAG}_samewindow.json 2> gpurun_out/bench_train_1gpu_${TAG}_samewindow.err
python -c "
import json; d=json.load(open('gpurun_out/bench_train_1gpu_${TAG}_samewindow.json')); print('1gpu', d['value'], d['ms_per_step'], d['config'].get('step_ms'))"
