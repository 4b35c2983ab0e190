#!/bin/bash
# ncu launch list of ONE graph-replayed train step (kernel nodes).  usage: bash scripts/gpu_ncu_graph.sh <tag> [first_step]
TAG=${1:-r2}
FIRST=${2:-2529}
mkdir -p gpurun_out
timeout 900 ncu --graph-profiling node --nvtx --nvtx-include "hgstep/" --metrics gpu__time_duration.sum \
    --clock-control none --csv --log-file gpurun_out/launches_graph_${TAG}.csv \
    python scripts/one_step_graph.py $FIRST 1 > gpurun_out/ncu_graph_${TAG}.log 2>&1
tail -3 gpurun_out/ncu_graph_${TAG}.log
python scripts/summarize_launches.py gpurun_out/launches_graph_${TAG}.csv | head -70
