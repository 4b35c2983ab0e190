import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from histogan_b200.trainer import SyntheticLoader, Trainer, NanException
tmp = tempfile.mkdtemp()
torch.manual_seed(0)
t = Trainer("t", tmp + "/r", tmp + "/m", image_size=32, network_capacity=16, batch_size=4, hist_insz=150,
            hist_resizing="interpolation", save_every=1000, cuda_graphs=True, fast_rng=True, nan_check="immediate")
t.loader = SyntheticLoader(4, 32, seed=0); t.loader_evaluate = SyntheticLoader(4, 32, seed=1, eval_batch=4)
t.init_GAN(); t.steps = 2501
t.train(alpha=2); print("clean", t.d_loss, t.g_loss)
t.save(2)
names = [n for n, _ in t.GAN.D.named_parameters()]
print("first D parameter:", names[0], next(iter(t.GAN.D.parameters())).shape)
with torch.no_grad():
    next(iter(t.GAN.D.parameters())).fill_(float("nan"))
x = torch.rand(4, 3, 32, 32, device="cuda")
with torch.no_grad():
    out, _ = t.GAN.D(x)
print("eager D(x) with the poisoned weight:", out.flatten().tolist())
try:
    t.train(alpha=2); print("after poison: no exception; d_loss", t.d_loss, "g_loss", t.g_loss)
except NanException:
    print("NanException raised")
