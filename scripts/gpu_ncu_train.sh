#!/bin/bash
# ncu evidence for the train step: launch list (all kernels, one step) + full capture of the
# tcgen05 conv kernels.  usage: bash scripts/gpu_ncu_train.sh <tag>
TAG=${1:-r1}
mkdir -p gpurun_out
cat > /tmp/one_step.py <<PY
import os, sys
sys.path.insert(0, os.getcwd())
import torch, bench
from histogan_b200.trainer import Trainer
dev = torch.device("cuda", 0)
out = os.path.join(bench.ROOT, "gpurun_out", "ncu_train")
tr = Trainer("p", out + "/results", out + "/models", image_size=256, network_capacity=16, batch_size=32,
             hist_insz=150, hist_resizing="interpolation", save_every=10 ** 9, fast_rng=True)   # eager (no graphs): ncu sees every kernel
tr.loader = bench.DeviceLoader(0, dev); tr.loader_evaluate = bench.DeviceLoader(0, dev, eval_only=True)
tr.steps = 2501
for _ in range(int(sys.argv[1])):
    tr.train()
torch.cuda.synchronize()
PY
# step 2501 = warm-up (skipped by -s), step 2502 profiled
ncu --metrics gpu__time_duration.sum --clock-control none -s 2400 -c 2600 --csv \
    --log-file gpurun_out/launches_train_$TAG.csv python /tmp/one_step.py 3 > gpurun_out/ncu_train_list_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_tf32_kernel -s 120 -c 6 \
    -o gpurun_out/prof_conv_$TAG python /tmp/one_step.py 2 > gpurun_out/ncu_conv_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_wgrad_tf32_kernel -s 60 -c 4 \
    -o gpurun_out/prof_wgrad_$TAG python /tmp/one_step.py 2 > gpurun_out/ncu_wgrad_$TAG.log 2>&1
ls -la gpurun_out | tail -8; tail -2 gpurun_out/ncu_conv_$TAG.log
