#!/bin/bash
# usage: bash scripts/gpu_r2i.sh <tag>: kernel tests for the rewritten kernels, micro-bench (fast vs generic), bench, launch list
TAG=$1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_fused_gpu.py -m gpu -q -x 2>&1 | tail -6 | cut -c1-300
timeout 300 python scripts/bench_small.py 2>&1 | grep -v Warn | tee gpurun_out/bench_small_$TAG.txt
echo "--- generic kernels"; HG_SMALL_GENERIC=1 timeout 300 python scripts/bench_small.py 2>&1 | grep "small" | tee gpurun_out/bench_small_generic_$TAG.txt
bash scripts/gpu_r2.sh $TAG tests bench
bash scripts/gpu_ncu_graph.sh $TAG 2529
