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


for case in (1, 32):
    gan.USE_FUSED, gan.STYLE_PATH = True, True
    run(case, "fused+style")
    gan.STYLE_PATH = False
    run(case, "fused, per-block styles")
    gan.USE_FUSED = False
    run(case, "composed torch ops")
