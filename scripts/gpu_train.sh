#!/bin/bash
TAG=${1:-r1}
mkdir -p gpurun_out
timeout 600 python bench.py --steps 8 --warmup 3 > gpurun_out/bench_train_$TAG.json 2> gpurun_out/bench_train_$TAG.err
tail -3 gpurun_out/bench_train_$TAG.err; cut -c1-1500 gpurun_out/bench_train_$TAG.json
timeout 600 python scripts/profile_train.py $TAG > gpurun_out/profile_train_$TAG.log 2>&1
tail -60 gpurun_out/profile_train_$TAG.log | cut -c1-260
