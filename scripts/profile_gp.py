"""GPU: which torch (at::) ops a gradient-penalty step still launches, with input shapes and the Python
line that issued them (eager path, torch.profiler).  Output: gpurun_out/gp_ops_<tag>.txt"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import bench
from histogan_b200.trainer import Trainer

tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
dev = torch.device("cuda", 0)
out_dir = os.path.join(bench.ROOT, "gpurun_out", "prof_gp")
tr = Trainer("p", out_dir + "/results", out_dir + "/models", image_size=256, network_capacity=16,
             batch_size=32, hist_insz=150, hist_resizing="interpolation", save_every=10 ** 9, fast_rng=True)
tr.loader = bench.DeviceLoader(0, dev)
tr.loader_evaluate = bench.DeviceLoader(0, dev, eval_only=True)
for s in (2524, 2525, 2532):
    tr.steps = s
    tr.train()
torch.cuda.synchronize()
tr.steps = 2536                                  # gradient penalty, no path length
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True, with_stack=True) as prof:
    tr.train()
    torch.cuda.synchronize()
rows = []
for e in prof.key_averages(group_by_input_shape=True, group_by_stack_n=6):
    t = getattr(e, "device_time_total", None) or getattr(e, "cuda_time_total", 0)
    if e.key.startswith("aten::") and t > 20:
        stack = [s for s in e.stack if "histogan_b200" in s or "torch/autograd" in s][:3]
        rows.append((t, e.count, e.key, str(e.input_shapes)[:90], " <- ".join(s.split("/")[-1][:60] for s in stack)))
rows.sort(reverse=True)
txt = "\n".join(f"{t:9.0f} us  x{c:<4d} {k:28s} {sh:90s} {st}" for t, c, k, sh, st in rows[:70])
open(os.path.join(bench.ROOT, "gpurun_out", f"gp_ops_{tag}.txt"), "w").write(txt)
print(txt)
