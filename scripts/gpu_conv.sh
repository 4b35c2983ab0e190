#!/bin/bash
TAG=${1:-r1}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_gpu.py tests/test_gan_gpu.py -m gpu -q -x 2>&1 | tail -5
timeout 600 python bench.py --steps 8 --warmup 3 > gpurun_out/bench_train_$TAG.json 2> gpurun_out/bench_train_$TAG.err
tail -3 gpurun_out/bench_train_$TAG.err; python - <<PY
import json
d=json.load(open('gpurun_out/bench_train_$TAG.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['achieved'], d['gpu_launches']/8)
print(d['roofline']['per_layer_tflops'])
PY
