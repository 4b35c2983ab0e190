#!/bin/bash
# A/B of the 1-GPU step flow (split G phase, gradient arena), then the hist / rehisto bench lines
TAG=$1
mkdir -p gpurun_out
ab() {
  local name=$1; shift
  env HG_BENCH_LIGHT=1 "$@" timeout 400 python bench.py --steps 32 --warmup 3 > gpurun_out/bench_ab_${TAG}_$name.json 2> gpurun_out/bench_ab_${TAG}_$name.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_ab_${TAG}_$name.json')); print('$name', d['value'], d['ms_per_step'], d['config'].get('step_ms'))"
}
ab default HG_X=1
ab split HG_SPLIT_G=1
ab split_arena HG_SPLIT_G=1 HG_GRAD_ARENA=1
ab arena HG_GRAD_ARENA=1
ab default2 HG_X=1
bash scripts/gpu_r2.sh $TAG hist rehisto
