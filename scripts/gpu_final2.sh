#!/bin/bash
# round-end evidence: all gpu tests, smoke, the three bench workloads, the graph launch list, the reference arm
TAG=${1:-final}
mkdir -p gpurun_out
bash scripts/gpu_r2.sh $TAG tests smoke bench hist rehisto
bash scripts/gpu_ncu_graph.sh $TAG 2529 | head -30
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_$TAG.json 2> gpurun_out/bench_reference_$TAG.err
cut -c1-400 gpurun_out/bench_reference_$TAG.json
