"""GPU: torch.profiler kernel-time table of a few Trainer.train steps at the bench config
(256x256, capacity 16, batch 32).  Output: gpurun_out/train_kernels_<tag>.txt"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import bench
from histogan_b200.trainer import Trainer

tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
dev = torch.device("cuda", 0)
out_dir = os.path.join(bench.ROOT, "gpurun_out", "prof_train")
tr = Trainer("p", out_dir + "/results", out_dir + "/models", image_size=256, network_capacity=16,
             batch_size=B, hist_insz=150, hist_resizing="interpolation", save_every=10 ** 9, fast_rng=True)
bench.B_PER_GPU = B
tr.loader = bench.DeviceLoader(0, dev)
tr.loader_evaluate = bench.DeviceLoader(0, dev, eval_only=True)
tr.steps = 2501
for _ in range(3):
    tr.train()
torch.cuda.synchronize()
tr.steps = 2505        # steps 2505, 2506, 2507: no GP / PL
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        tr.train()
    torch.cuda.synchronize()
txt = prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=90)
open(os.path.join(bench.ROOT, "gpurun_out", f"train_kernels_{tag}.txt"), "w").write(txt)
print(txt[-6000:])
