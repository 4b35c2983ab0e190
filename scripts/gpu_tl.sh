#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 scripts/ddp_timeline.py $N 2>&1 | grep -v -i "warn" | tail -40
echo "=== 1 GPU, split G forced"; HG_SPLIT_G=1 timeout 300 python scripts/ddp_timeline.py 1 2>&1 | grep -v -i warn | tail -30
echo "=== 1 GPU default"; timeout 300 python scripts/ddp_timeline.py 1 2>&1 | grep -v -i warn | tail -30
