#!/bin/bash
# which descriptor convention makes the single-box halo conv (HALO == 2) correct?
mkdir -p gpurun_out
for m in 1 2; do
  HG_CONV_HALO2=$m timeout 300 python -m pytest tests/test_conv_gpu.py -q -x -k "matches_torch" > gpurun_out/halo2_mode$m.log 2>&1
  echo "mode $m: $(tail -1 gpurun_out/halo2_mode$m.log)"
done
