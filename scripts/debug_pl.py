"""GPU debug: G-phase gradients of the path-length step (golden case 32) under several switches"""
import os, sys, tempfile, pathlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from histogan_b200 import gan
from oracle import make_golden_step as mgs
from tests import step_checks as sc
from tests.test_trainer_gpu import _golden_trainer

g = sc.load_golden()


def run(case, tag):
    with tempfile.TemporaryDirectory() as tmp:
        t = _golden_trainer(pathlib.Path(tmp))
        images, hists = mgs.step_inputs(case)
        t.loader = iter([{"images": images, "histograms": hists[0]}, {"images": images, "histograms": hists[1]}])
        t.steps, t.pl_mean = case, 0
        mgs.seed_step(case)
        t.train(alpha=mgs.ALPHA)
        tab = sc.compare_grads(t.GAN.G_opt.recorded, g["names_g"], g[f"c{case}_g_norms"], g[f"c{case}_g_samples"])
        bad = {k: (round(v[0], 3), round(v[1], 4)) for k, v in tab.items() if v[1] < 0.99}
        print(f"[{tag}] case {case}: worst {sc.worst(tab)}\n   {len(bad)} tensors below 0.99: {dict(list(bad.items())[:8])}", flush=True)


for case in (32,):
    gan.USE_FUSED, gan.STYLE_PATH, gan.DEMOD_PATH = True, True, True
    run(case, "fused+style+demod")
    gan.DEMOD_PATH = False
    run(case, "fused+style, torch demod")
    gan.DEMOD_PATH = True
    import histogan_b200.fused as fz
    real_bwd = fz.grouped_linear_bwd
    def torch_bwd(xs, ws, gys, gws, gbs, gxs, flags=0):
        for x, w, gy, gw, gb, gx in zip(xs, ws, gys, gws, gbs, gxs):
            xe = x * x if flags & fz.LIN_SQUARE_INPUT else x
            if gw is not None: gw.copy_(gy.t() @ xe)
            if gb is not None: gb.copy_(gy.sum(0))
            if gx is not None:
                v = gy @ w
                if flags & fz.LIN_POST_2X: v = v * 2 * x
                gx.copy_(gx + v if flags & fz.LIN_ACCUMULATE else v)
    fz.grouped_linear_bwd = torch_bwd
    run(case, "fused+style+demod, style-linear backward by torch")
    fz.grouped_linear_bwd = real_bwd
    real_fwd = fz.grouped_linear
    def torch_fwd(xs, ws, bs, flags=0, slope=0.2, eps=1e-8):
        out = []
        for x, w, b in zip(xs, ws, bs):
            xe = x * x if flags & fz.LIN_SQUARE_INPUT else x
            v = xe @ w.t()
            if b is not None: v = v + b
            if flags & fz.LIN_RSQRT_EPS: v = torch.rsqrt(v + eps)
            if flags & fz.LIN_LRELU: v = torch.nn.functional.leaky_relu(v, slope)
            if flags & fz.LIN_ADD_ONE: v = v + 1
            out.append(v.contiguous())
        return out
    fz.grouped_linear = torch_fwd
    run(case, "fused+style+demod, grouped forward by torch")
    fz.grouped_linear = real_fwd
