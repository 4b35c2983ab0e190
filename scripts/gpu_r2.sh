#!/bin/bash
# Round-2 GPU session.  usage: bash scripts/gpu_r2.sh <tag> [tests] [smoke] [bench] [hist] [rehisto] [profile]
TAG=${1:-r2a}
mkdir -p gpurun_out
rm -f gpurun_out/parity_records.jsonl
if [[ " $* " == *" tests "* ]]; then
  timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu_$TAG.log 2>&1
  tail -25 gpurun_out/pytest_gpu_$TAG.log | cut -c1-300
fi
if [[ " $* " == *" smoke "* ]]; then
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
fi
if [[ " $* " == *" bench "* ]]; then
  timeout 900 python bench.py > gpurun_out/bench_train_$TAG.json 2> gpurun_out/bench_train_$TAG.err
  tail -3 gpurun_out/bench_train_$TAG.err | cut -c1-300; python - <<PY
import json
d=json.load(open('gpurun_out/bench_train_$TAG.json'))
print({k: d[k] for k in ('value','ms_per_step','steps','gpu_launches')}, 'e2e', d['e2e']['value'], 'convTF', d['roofline']['achieved'])
print({k: v for k, v in d['config'].items() if k in ('step_ms','final_losses','peak_mem_gib')})
PY
fi
if [[ " $* " == *" hist "* ]]; then
  timeout 300 python bench.py --workload hist --steps 20 --warmup 3 > gpurun_out/bench_hist_$TAG.json 2> gpurun_out/bench_hist_$TAG.err
  cut -c1-300 gpurun_out/bench_hist_$TAG.json
fi
if [[ " $* " == *" rehisto "* ]]; then
  timeout 300 python bench.py --workload rehisto --steps 8 --warmup 3 > gpurun_out/bench_rehisto_$TAG.json 2> gpurun_out/bench_rehisto_$TAG.err
  cut -c1-300 gpurun_out/bench_rehisto_$TAG.json
fi
if [[ " $* " == *" refgpu "* ]]; then
  timeout 900 python bench.py --impl reference-gpu --steps 4 > gpurun_out/bench_refgpu_$TAG.json 2> gpurun_out/bench_refgpu_$TAG.err
  tail -2 gpurun_out/bench_refgpu_$TAG.err | cut -c1-300; cut -c1-900 gpurun_out/bench_refgpu_$TAG.json
fi
if [[ " $* " == *" profile "* ]]; then
  timeout 600 python scripts/profile_train.py $TAG > gpurun_out/profile_train_$TAG.log 2>&1
  tail -60 gpurun_out/profile_train_$TAG.log | cut -c1-200
fi
