"""one Trainer.train step of the bench configuration in CUDA-graph mode inside an NVTX range,
for `ncu --graph-profiling node --nvtx --nvtx-include "hgstep/"` (lists the kernel nodes of the
replayed graphs).  usage: python scripts/one_step_graph.py [first_step] [n_steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from histogan_b200.trainer import Trainer

first = int(sys.argv[1]) if len(sys.argv) > 1 else 2529
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1
dev = torch.device("cuda", 0)
out = os.path.join(bench.ROOT, "gpurun_out", "ncu_graph")
tr = Trainer("p", out + "/results", out + "/models", image_size=256, network_capacity=16, batch_size=32,
             hist_insz=150, hist_resizing="interpolation", save_every=10 ** 9, fast_rng=True, cuda_graphs=True)
tr.loader = bench.DeviceLoader(0, dev)
tr.loader_evaluate = bench.DeviceLoader(0, dev, eval_only=True)
for st in (2528, 2529, 2530, 2532):          # capture every variant
    tr.steps = st
    tr.train()
torch.cuda.synchronize()
tr.steps = first
torch.cuda.nvtx.range_push("hgstep")
for _ in range(n):
    tr.train()
torch.cuda.synchronize()
torch.cuda.nvtx.range_pop()
