#!/bin/bash
# usage: bash scripts/gpu_r2e.sh <tag>   -- tests, bench (default), bench with HG_CH_PAD=16, graph launch list
TAG=$1
bash scripts/gpu_r2.sh $TAG tests bench
HG_CH_PAD=16 timeout 600 python bench.py > gpurun_out/bench_train_${TAG}_ch16.json 2> gpurun_out/bench_train_${TAG}_ch16.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_train_${TAG}_ch16.json'))
print('CH16:', {k: d[k] for k in ('value','ms_per_step')}, d['config'].get('step_ms'), 'convTF', d['roofline']['achieved'], d['roofline']['pass_us'])
PY
HG_CH_PAD=16 timeout 600 python -m pytest tests/test_gan_gpu.py tests/test_conv_gpu.py -m gpu -q 2>&1 | tail -3
bash scripts/gpu_ncu_graph.sh $TAG 2529
