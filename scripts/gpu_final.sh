#!/bin/bash
# Round-end evidence in one GPU session: full gpu tests, the three bench workloads, the ncu launch
# list of one (eager) train step and --set full captures of the dominant conv kernels.
# usage: bash scripts/gpu_final.sh <tag>
TAG=${1:-final}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -3 gpurun_out/pytest_gpu_$TAG.log
timeout 400 python bench.py --steps 16 --warmup 3 > gpurun_out/bench_train_$TAG.json 2> gpurun_out/bench_train_$TAG.err
timeout 300 python bench.py --workload hist --steps 20 --warmup 3 > gpurun_out/bench_hist_$TAG.json 2> gpurun_out/bench_hist_$TAG.err
timeout 300 python bench.py --workload rehisto --steps 8 --warmup 3 > gpurun_out/bench_rehisto_$TAG.json 2> gpurun_out/bench_rehisto_$TAG.err
for f in train hist rehisto; do cut -c1-180 gpurun_out/bench_${f}_$TAG.json; done
cat > /tmp/one_step.py <<PY
import os, sys
sys.path.insert(0, os.getcwd())
import torch, bench
from histogan_b200.trainer import Trainer
dev = torch.device("cuda", 0)
out = os.path.join(bench.ROOT, "gpurun_out", "ncu_train")
tr = Trainer("p", out + "/results", out + "/models", image_size=256, network_capacity=16, batch_size=32,
             hist_insz=150, hist_resizing="interpolation", save_every=10 ** 9, fast_rng=True)   # eager: ncu sees every kernel
tr.loader = bench.DeviceLoader(0, dev); tr.loader_evaluate = bench.DeviceLoader(0, dev, eval_only=True)
tr.steps = 2501
for _ in range(int(sys.argv[1])):
    tr.train()
torch.cuda.synchronize()
PY
# steps 2501 (gradient penalty), 2502: the second one is listed
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2300 -c 2600 --csv \
    --log-file gpurun_out/launches_train_$TAG.csv python /tmp/one_step.py 2 > gpurun_out/ncu_train_list_$TAG.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_tf32_kernel -s 2 -c 1 \
    -o gpurun_out/prof_conv32_$TAG python scripts/one_conv.py fwd 32 32 256 > gpurun_out/ncu_conv32_$TAG.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_tf32_kernel -s 2 -c 1 \
    -o gpurun_out/prof_conv256_$TAG python scripts/one_conv.py fwd 256 256 32 > gpurun_out/ncu_conv256_$TAG.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_wgrad_col -s 2 -c 1 \
    -o gpurun_out/prof_wgcol_$TAG python scripts/one_conv.py wgrad 32 32 256 > gpurun_out/ncu_wgcol_$TAG.log 2>&1
ls -la gpurun_out | grep $TAG | tail -12
