#!/bin/bash
TAG=$1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fused_gpu.py -m gpu -q -k "to_rgb" 2>&1 | tail -3 | cut -c1-300
timeout 300 python scripts/bench_small.py 2>&1 | grep -E "to_rgb|small" | tee gpurun_out/bench_small_$TAG.txt
bash scripts/gpu_r2.sh $TAG tests bench
bash scripts/gpu_ncu_graph.sh $TAG 2529
