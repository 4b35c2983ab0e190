#!/bin/bash
TAG=$1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_fused_gpu.py -m gpu -q 2>&1 | tail -6 | cut -c1-300
timeout 200 python scripts/debug_small.py 2>&1 | grep -v Warn
HG_SMALL_GENERIC=1 timeout 200 python scripts/debug_small.py 2>&1 | grep -v Warn
timeout 600 python -m pytest tests/test_gan_gpu.py -m gpu -q -k "discriminator_256" 2>&1 | tail -4 | cut -c1-300
HG_SMALL_GENERIC=1 timeout 600 python -m pytest tests/test_gan_gpu.py -m gpu -q -k "discriminator_256" 2>&1 | tail -4 | cut -c1-300
timeout 300 ncu --set full --clock-control none --import-source on -k regex:torgb_fwd -s 2 -c 1 \
    -o gpurun_out/prof_torgb_$TAG python scripts/one_torgb.py > gpurun_out/ncu_torgb_$TAG.log 2>&1
tail -1 gpurun_out/ncu_torgb_$TAG.log
