#!/bin/bash
# halo2 descriptor-mode detection, then the regular session with the working mode
bash scripts/gpu_halo2.sh
M=0
grep -q "passed" gpurun_out/halo2_mode2.log && ! grep -q "failed" gpurun_out/halo2_mode2.log && M=2
grep -q "passed" gpurun_out/halo2_mode1.log && ! grep -q "failed" gpurun_out/halo2_mode1.log && M=1
echo "HALO2 mode selected: $M"
export HG_CONV_HALO2=$M
bash scripts/gpu_r2.sh $1 tests bench
bash scripts/gpu_ncu_graph.sh $1 2529
