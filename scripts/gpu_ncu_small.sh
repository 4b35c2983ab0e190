#!/bin/bash
# ncu --set full captures of the constant-memory image-input conv kernels and the bulk-copy torgb_fwd
TAG=$1
mkdir -p gpurun_out
cap() {  # name, kernel regex, skip, script, args...
  local name=$1 rx=$2 skip=$3 script=$4; shift 4
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 \
      -o gpurun_out/prof_${name}_$TAG python $script "$@" > gpurun_out/ncu_${name}_$TAG.log 2>&1
  tail -1 gpurun_out/ncu_${name}_$TAG.log | cut -c1-200
}
cap small16fwd conv_small_fwd16 2 scripts/one_conv.py small 3 16 256
cap small16dgrad conv_small_dgrad16 2 scripts/one_conv.py small 3 16 256
cap small16wgrad conv_small_wgrad16_k3 2 scripts/one_conv.py small 3 16 256
cap torgbbulk torgb_fwd 2 scripts/one_torgb.py
