"""Golden vectors for ONE optimisation step, produced by the UNMODIFIED reference
``Trainer.train`` (histoGAN/histoGAN.py:853-1020) running on the CPU.  TEST INFRASTRUCTURE.

    python -m oracle.make_golden_step

How the reference is made to run without a GPU (nothing in /root/reference is edited):
``Tensor.cuda`` / ``Module.cuda`` are patched to identities for the duration of the run, the
un-vendored ``torch_optimizer.DiffGrad`` is replaced by a RECORDING optimiser whose ``step()``
stores clones of the gradients and leaves the weights alone.  The golden therefore pins the
step's loss composition and every parameter gradient of both phases at one set of weights,
independently of the (parity-unpinned) optimiser.

Cases: trainer.steps = 1 (plain step), 4 (gradient penalty), 32 (gradient penalty +
path-length regulariser), each with alpha = 2 (the reference's histogram-loss weight) and with
alpha = 0 (keys `c<steps>n_*`).  Why the second set: the histogram term's gradient carries a factor
1/(I + 1e-6) per pixel (d log(I + eps)/dI, RGBuvHistBlock.py:112-115), so the few generated pixels
that happen to lie at +1e-6..1e-4 dominate it, and whether a pixel sits at +1e-5 or -1e-5 (cut by the
relu) is decided by rounding noise far below any parity tolerance.  The alpha = 0 gradients (adversarial
+ path-length terms) are well conditioned and pin the step; the alpha = 2 set is compared whenever
the realisation at hand is not dominated by such a pixel (tests/test_trainer_gpu.py).  Image 64x64, network_capacity 16, batch 2, hist_insz 150
'interpolation' (no resize at 64x64), alpha = 2.  Weights: gan_oracle.seeded_state_dict;
inputs: seeded generators; the latents come from the global CPU RNGs seeded per case, which
``train_oracle.draw_step_inputs`` reproduces.

Gradients are stored as fingerprints (float64 norm + up to 256 evenly strided entries per
tensor): a 1.5 MB fixture instead of 100 MB.
"""
from __future__ import annotations

import json
import os
import random
import sys
import tempfile
import types

import numpy as np
import torch

from . import gan_oracle as go
from . import ref_shim

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                          "tests", "golden")
IMAGE_SIZE, CAPACITY, BATCH, ALPHA = 64, 16, 2, 2.0
CASES = (1, 4, 32)
ALPHAS = (("", 2.0), ("n", 0.0))          # key suffix, alpha
N_SAMPLES = 256
SEEDS = {"G": 21, "D": 22, "S": 23, "H": 24}


def fingerprint_indices(numel):
    """evenly strided flat indices (all of them for small tensors)"""
    if numel <= N_SAMPLES:
        return np.arange(numel)
    return (np.arange(N_SAMPLES, dtype=np.int64) * numel) // N_SAMPLES


def fingerprint(t):
    flat = t.detach().cpu().contiguous().reshape(-1)
    return float(flat.double().norm()), flat[torch.from_numpy(fingerprint_indices(flat.numel()))].float().numpy()


def step_inputs(case):
    g = torch.Generator().manual_seed(7000 + case)
    images = torch.rand(BATCH, 3, IMAGE_SIZE, IMAGE_SIZE, generator=g)
    hists = []
    for _ in range(2):                      # D-phase batch, G-phase batch (two loader reads)
        t = torch.rand(BATCH, 3, 64, 64, generator=g)
        hists.append(t / t.sum(dim=(1, 2, 3), keepdim=True))
    return images, hists


def seed_step(case):
    torch.manual_seed(900 + case)
    random.seed(900 + case)


def seeded_gan_state(gan):
    """seeded weights for the S/H/G/D members of a HistoGAN container (ours or the reference's)"""
    sd = {}
    for name, seed in SEEDS.items():
        m = getattr(gan, name)
        shapes = {k: list(v.shape) for k, v in m.state_dict().items()}
        sd.update({f"{name}.{k}": v for k, v in go.seeded_state_dict(shapes, seed).items()})
    return sd


class RecordingOptimizer(torch.optim.Optimizer):
    """stands in for torch_optimizer.DiffGrad: records gradients, never moves the weights"""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), **kw):
        super().__init__(params, dict(lr=lr, betas=betas))
        self.recorded = None

    def step(self, closure=None):
        self.recorded = [None if p.grad is None else p.grad.detach().clone()
                         for g in self.param_groups for p in g["params"]]


def main():
    assert ref_shim.available()
    # our recording optimiser must be what `from torch_optimizer import DiffGrad` finds
    stub = types.ModuleType("torch_optimizer")
    stub.DiffGrad = RecordingOptimizer
    sys.modules["torch_optimizer"] = stub
    sys.modules.pop("histoGAN.histoGAN", None)
    gm = ref_shim.ref_gan_module()
    assert gm.DiffGrad is RecordingOptimizer
    torch.set_num_threads(os.cpu_count() or 1)

    saved = (torch.Tensor.cuda, torch.nn.Module.cuda)
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    out = {}
    try:
        with tempfile.TemporaryDirectory() as tmp:
            tr = gm.Trainer("g", tmp + "/results", tmp + "/models", image_size=IMAGE_SIZE,
                            network_capacity=CAPACITY, batch_size=BATCH, hist_insz=150,
                            hist_resizing="interpolation", save_every=10 ** 9)
            tr.histBlock.device = "cpu"          # attribute of the reference object, not its code
            tr.init_GAN()
            tr.GAN.load_state_dict(seeded_gan_state(tr.GAN), strict=False)
            tr.GAN.reset_parameter_averaging()
            names_d = [k for k, _ in tr.GAN.D.named_parameters()]
            names_g = (["G." + k for k, _ in tr.GAN.G.named_parameters()] +
                       ["S." + k for k, _ in tr.GAN.S.named_parameters()] +
                       ["H." + k for k, _ in tr.GAN.H.named_parameters()])
            out["names_d"], out["names_g"] = json.dumps(names_d), json.dumps(names_g)
            for case in CASES:
              for suffix, alpha in ALPHAS:
                images, hists = step_inputs(case)
                tr.loader = iter([{"images": images, "histograms": hists[0]},
                                  {"images": images, "histograms": hists[1]}])
                tr.steps, tr.pl_mean = case, 0
                seed_step(case)
                tr.train(alpha=alpha)
                assert tr.steps == case + 1
                rec = {"d_loss": tr.d_loss, "g_loss": tr.g_loss, "h_loss": tr.h_loss,
                       "gp": tr.last_gp_loss if case % 4 == 0 else float("nan"),
                       "pl_mean": float(tr.pl_mean)}
                key = f"c{case}{suffix}"
                out[f"{key}_scalars"] = json.dumps(rec)
                for tag, opt, names in (("d", tr.GAN.D_opt, names_d), ("g", tr.GAN.G_opt, names_g)):
                    assert len(opt.recorded) == len(names)
                    norms, samples = [], []
                    for nm, g in zip(names, opt.recorded):
                        assert g is not None, nm
                        n, s = fingerprint(g)
                        norms.append(n)
                        samples.append(s)
                    out[f"{key}_{tag}_norms"] = np.array(norms, dtype=np.float64)
                    out[f"{key}_{tag}_samples"] = np.concatenate(samples)
                print(f"case steps={case} alpha={alpha}:", rec)
    finally:
        torch.Tensor.cuda, torch.nn.Module.cuda = saved
    np.savez_compressed(os.path.join(GOLDEN_DIR, "train_step_64.npz"), **out)
    print("wrote train_step_64.npz", {k: getattr(v, "shape", None) for k, v in out.items()})


if __name__ == "__main__":
    main()
