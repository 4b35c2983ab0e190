"""GPU: torch.profiler kernel-time table of a few recoloringTrainer.train steps at the bench
config (256x256, capacity 16, batch 16).  Output: gpurun_out/rehisto_kernels_<tag>.txt"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import bench
from histogan_b200.rehistogan import recoloringTrainer

tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
dev = torch.device("cuda", 0)
out_dir = os.path.join(bench.ROOT, "gpurun_out", "prof_rehisto")
tr = recoloringTrainer("p", out_dir + "/results", out_dir + "/models", image_size=256, network_capacity=16,
                       batch_size=bench.RH_BATCH, skip_conn_to_GAN=True, initialize_gan=True,
                       save_every=10 ** 9, fast_rng=True)
tr.loader = bench.RecolorLoader(0, dev)
tr.steps = 2501
for _ in range(3):
    tr.train()
torch.cuda.synchronize()
tr.steps = 2505
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        tr.train()
    torch.cuda.synchronize()
txt = prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=90)
open(os.path.join(bench.ROOT, "gpurun_out", f"rehisto_kernels_{tag}.txt"), "w").write(txt)
print(txt[-6000:])
