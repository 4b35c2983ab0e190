#!/bin/bash
TAG=$1
mkdir -p gpurun_out
timeout 300 python scripts/debug_small2.py 2>&1 | grep -v Warn | cut -c1-400
HG_SMALL_GENERIC=1 timeout 300 python scripts/debug_small2.py 2>&1 | grep -v Warn | cut -c1-400
timeout 600 python -m pytest tests/test_conv_gpu.py -m gpu -q -k "image_input" 2>&1 | tail -3 | cut -c1-300
timeout 300 python scripts/bench_small.py 2>&1 | grep -E "wgrad|to_rgb|linear"
