#!/bin/bash
# One GPU-box session: gpu tests, bench, ncu launch list, ncu full capture of the
# histogram kernels.  Usage: gpurun -- 'bash scripts/gpu_round.sh <tag>'
TAG=${1:-r1}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/pytest_gpu_$TAG.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv \
    --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_bench_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:hist_.*_fast_kernel -s 4 -c 2 \
    -o gpurun_out/prof_hist_$TAG python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_full_$TAG.log 2>&1
tail -5 gpurun_out/pytest_gpu_$TAG.log; cut -c1-400 gpurun_out/bench_$TAG.json
