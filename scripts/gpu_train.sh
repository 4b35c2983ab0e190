#!/bin/bash
# usage: bash scripts/gpu_train.sh <tag> [tests] [profile]
TAG=${1:-r1}
mkdir -p gpurun_out
if [[ " $* " == *" tests "* ]]; then
  timeout 900 python -m pytest tests/test_gan_gpu.py tests/test_trainer_gpu.py -m gpu -q -s > gpurun_out/pytest_gan_$TAG.log 2>&1
  grep -E "passed|failed|^FAILED|Error|graphed" gpurun_out/pytest_gan_$TAG.log | grep -v print | cut -c1-400
fi
timeout 600 python bench.py --steps 8 --warmup 3 > gpurun_out/bench_train_$TAG.json 2> gpurun_out/bench_train_$TAG.err
tail -3 gpurun_out/bench_train_$TAG.err; python - <<PY
import json
d=json.load(open('gpurun_out/bench_train_$TAG.json'))
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'convTF', d['roofline']['achieved'], 'launches/step', d['gpu_launches']/8, 'graphs', d['config'].get('cuda_graphs'))
PY
HG_CUDA_GRAPHS=0 timeout 600 python bench.py --steps 8 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('eager: value', d['value'], 'ms', d['ms_per_step'])"
if [[ " $* " == *" profile "* ]]; then
  timeout 600 python scripts/profile_train.py $TAG > gpurun_out/profile_train_$TAG.log 2>&1
  tail -4 gpurun_out/profile_train_$TAG.log | cut -c1-200
fi
