#!/bin/bash
# usage: bash scripts/gpu_train.sh <tag> [tests]
TAG=${1:-r1}
mkdir -p gpurun_out
if [ "$2" == "tests" ]; then
  timeout 900 python -m pytest tests/test_gan_gpu.py tests/test_trainer_gpu.py -m gpu -q -s > gpurun_out/pytest_gan_$TAG.log 2>&1
  grep -E "passed|failed|^FAILED|Error|vs golden" gpurun_out/pytest_gan_$TAG.log | grep -v print | cut -c1-600
fi
timeout 600 python bench.py --steps 8 --warmup 3 > gpurun_out/bench_train_$TAG.json 2> gpurun_out/bench_train_$TAG.err
tail -3 gpurun_out/bench_train_$TAG.err; cut -c1-300 gpurun_out/bench_train_$TAG.json
timeout 600 python scripts/profile_train.py $TAG > gpurun_out/profile_train_$TAG.log 2>&1
tail -4 gpurun_out/profile_train_$TAG.log | cut -c1-200
