import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from histogan_b200 import rehistogan as rh
from histogan_b200.trainer import SyntheticLoader
torch.manual_seed(0)
t = rh.recoloringTrainer("g", "gpurun_out/dbg_res", "gpurun_out/dbg_mod", 64, 16, batch_size=4,
                         skip_conn_to_GAN=True, initialize_gan=True, save_every=10 ** 9,
                         fast_rng=True, cuda_graphs=True)
t.loader = SyntheticLoader(4, 64, seed=0)
t.init_GAN(); t.GAN.train()
batch = next(t.loader)
t._static = {'images': batch['images'].clone(), 'hists': batch['histograms'].clone()}
fixed = torch.rand(4, 64, 64, 1, device='cuda')
real_rand = torch.rand
torch.rand = lambda *a, **k: fixed
g_named = [(k, p) for k, p in t.GAN.named_parameters() if not k.startswith("D.")]
params = [p for _, p in g_named]
fn = lambda: t._phase_g(32.0, 1.5, 4.0)
def grads():
    return {k: p.grad.detach().clone() for k, p in g_named if p.grad is not None}
o1 = [o.item() for o in fn()]; g1 = grads()
o2 = [o.item() for o in fn()]; g2 = grads()
print("eager vs eager losses", o1, o2)
d = sorted(((float((g1[k] - g2[k]).norm() / g2[k].norm().clamp_min(1e-20)), k) for k in g1), reverse=True)[:5]
print("eager vs eager worst grads", d)
o3 = [o.item() for o in t._graphed(('G',), fn, params)]; g3 = grads()
print("graph losses", o3)
d = sorted(((float((g1[k] - g3[k]).norm() / g1[k].norm().clamp_min(1e-20)), k) for k in g1), reverse=True)
print("eager vs graph worst grads", d[:12])
print("n bad", sum(1 for v, _ in d if v > 1e-2), "of", len(d))
# ---- D phase with gradient penalty: eager twice, then graph
d_named = [(k, p) for k, p in t.GAN.named_parameters() if k.startswith("D.")]
dparams = [p for _, p in d_named]
fd = lambda: t._phase_d(True)
def dgrads():
    return {k: p.grad.detach().clone() for k, p in d_named if p.grad is not None}
a1 = [o.item() for o in fd()]; h1 = dgrads()
a2 = [o.item() for o in fd()]; h2 = dgrads()
print("D eager vs eager (divergence, gp)", a1, a2)
a3 = [o.item() for o in t._graphed(('D', True), fd, dparams)]; h3 = dgrads()
print("D graph", a3)
d = sorted(((float((h1[k] - h3[k]).norm() / h1[k].norm().clamp_min(1e-20)), k) for k in h1), reverse=True)
print("D eager vs graph worst grads", d[:4])
d = sorted(((float((h1[k] - h2[k]).norm() / h1[k].norm().clamp_min(1e-20)), k) for k in h1), reverse=True)
print("D eager vs eager worst grads", d[:4])
