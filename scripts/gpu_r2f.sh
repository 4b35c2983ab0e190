#!/bin/bash
TAG=$1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_gpu.py -m gpu -q 2>&1 | tail -8 | cut -c1-250
timeout 600 python scripts/debug_pl.py 2>&1 | grep -v Warning | tail -20 | cut -c1-600
bash scripts/gpu_r2.sh $TAG tests bench
bash scripts/gpu_ncu_graph.sh $TAG 2529
