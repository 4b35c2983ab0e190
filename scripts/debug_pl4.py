"""GPU debug: full path-length step (golden case 32) with the grouped-linear forward switched per use"""
import os, sys, tempfile, pathlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
import histogan_b200.fused as fz
from oracle import make_golden_step as mgs
from tests import step_checks as sc
from tests.test_trainer_gpu import _golden_trainer

g = sc.load_golden()
real = fz.grouped_linear


def torch_fwd(xs, ws, bs, flags=0, slope=0.2, eps=1e-8):
    out = []
    for x, w, b in zip(xs, ws, bs):
        xe = x * x if flags & fz.LIN_SQUARE_INPUT else x
        v = xe @ w.t()
        if b is not None: v = v + b
        if flags & fz.LIN_RSQRT_EPS: v = torch.rsqrt(v + eps)
        if flags & fz.LIN_LRELU: v = F.leaky_relu(v, slope)
        if flags & fz.LIN_ADD_ONE: v = v + 1
        out.append(v.contiguous())
    return out


def make(demod_impl, style_impl, mlp_impl):
    def f(xs, ws, bs, flags=0, slope=0.2, eps=1e-8):
        impl = demod_impl if flags & fz.LIN_RSQRT_EPS else (mlp_impl if flags & fz.LIN_LRELU else style_impl)
        return impl(xs, ws, bs, flags, slope, eps)
    return f


def run(tag):
    case = 32
    with tempfile.TemporaryDirectory() as tmp:
        t = _golden_trainer(pathlib.Path(tmp))
        images, hists = mgs.step_inputs(case)
        t.loader = iter([{"images": images, "histograms": hists[0]}, {"images": images, "histograms": hists[1]}])
        t.steps, t.pl_mean = case, 0
        mgs.seed_step(case)
        t.train(alpha=mgs.ALPHA)
        tab = sc.compare_grads(t.GAN.G_opt.recorded, g["names_g"], g[f"c{case}_g_norms"], g[f"c{case}_g_samples"])
        bad = {k: (round(v[0], 3), round(v[1], 4)) for k, v in tab.items() if v[1] < 0.99}
        print(f"[{tag}] {len(bad)} tensors below 0.99; worst cos {sc.worst(tab)['cosine']}", flush=True)


for tag, impls in (("all mine", (real, real, real)), ("demod torch", (torch_fwd, real, real)),
                   ("style torch", (real, torch_fwd, real)), ("mlp torch", (real, real, torch_fwd)),
                   ("demod mine only", (real, torch_fwd, torch_fwd))):
    fz.grouped_linear = make(*impls)
    run(tag)
fz.grouped_linear = real
# sync after every grouped-linear forward (timing / ordering hypothesis)
def synced(xs, ws, bs, flags=0, slope=0.2, eps=1e-8):
    out = real(xs, ws, bs, flags, slope, eps)
    torch.cuda.synchronize()
    return out
fz.grouped_linear = synced
run("all mine + device sync after each forward launch")
fz.grouped_linear = real
os.environ["CUDA_LAUNCH_BLOCKING"] = "1"
