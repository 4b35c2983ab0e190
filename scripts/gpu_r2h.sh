#!/bin/bash
# usage: bash scripts/gpu_r2h.sh <tag>: conv tests, all gpu tests, the three benches, ncu --set full captures
TAG=$1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_gpu.py -m gpu -q 2>&1 | tail -6 | cut -c1-250
bash scripts/gpu_r2.sh $TAG tests bench hist rehisto
cap() {  # name, kernel regex, skip, one_conv args...
  local name=$1 rx=$2 skip=$3; shift 3
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 \
      -o gpurun_out/prof_${name}_$TAG python scripts/one_conv.py "$@" > gpurun_out/ncu_${name}_$TAG.log 2>&1
  tail -1 gpurun_out/ncu_${name}_$TAG.log | cut -c1-200
}
cap conv32 conv_tf32_kernel 2 fwd 32 32 256
cap conv256 conv_tf32_kernel 2 fwd 256 256 32
cap wgrad32 conv_wgrad 2 wgrad 32 32 256
cap smallfwd conv_small_fwd 2 small 3 16 256
cap smalldgrad conv_small_dgrad 2 small 3 16 256
cap smallwgrad conv_small_wgrad_kernel 2 small 3 16 256
ls -la gpurun_out | grep $TAG | tail -14
