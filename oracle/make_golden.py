"""Generate tests/golden/*.npz by running the UNMODIFIED reference here.

    python -m oracle.make_golden

TEST INFRASTRUCTURE.  Needs /root/reference (this container only); the fixtures
travel to the GPU box, the reference does not.  Every fixture stores the exact
input, the constructor kwargs and the reference's outputs (and autograd
gradients) so that neither the oracle nor the CUDA path can drift unnoticed.
"""
from __future__ import annotations

import json
import os

import numpy as np
import torch
import torch.nn.functional as F

from . import hist_oracle as ho
from . import ref_shim

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                          "tests", "golden")

HIST_CASES = [
    # name, input maker, B, S, C, kwargs, alpha
    ("hist_c1_uniform_4x64", ho.synth_uniform, 4, 64, 3, dict(), 2.0),
    ("hist_interp_genlike_2x160", ho.synth_generator_like, 2, 160, 3, dict(insz=150), 2.0),
    ("hist_sampling_signed_2x160", ho.synth_signed, 2, 160, 3,
     dict(insz=150, resizing="sampling"), 32.0),
    ("hist_rbf_signed_2x48", ho.synth_signed, 2, 48, 3, dict(method="RBF"), 2.0),
    ("hist_threshold_uniform_2x48", ho.synth_uniform, 2, 48, 3, dict(method="thresholding"), 2.0),
    ("hist_generic_h32_rgba_2x40", ho.synth_signed, 2, 40, 4,
     dict(h=32, green_only=True, intensity_scale=False, hist_boundary=[-2, 4], sigma=0.05), 2.0),
    ("hist_h64_asym_boundary_2x40", ho.synth_generator_like, 2, 40, 3,
     dict(hist_boundary=[-2.5, 3], intensity_scale=True), 2.0),
]


def run_reference_hist(x, target, alpha, kwargs):
    """histBlock(F.relu(x)) + Hellinger loss + autograd grad, reference code only
    (histoGAN/histoGAN.py:955-960)."""
    mod = ref_shim.ref_hist_module()
    kw = {k: (list(v) if isinstance(v, (list, tuple)) else v) for k, v in kwargs.items()}
    block = mod.RGBuvHistBlock(device="cpu", **kw)
    xr = x.clone().requires_grad_(True)
    hist = block(F.relu(xr))
    scale = 1 / np.sqrt(2.0)
    loss = alpha * scale * (torch.sqrt(torch.sum(torch.pow(
        torch.sqrt(target) - torch.sqrt(hist), 2)))) / target.shape[0]
    # thresholding: the masks are constants for autograd but Iy (intensity_scale) still
    # carries a gradient -- ask the reference instead of assuming
    if hist.requires_grad:
        # (a) the training loss; (b) a linear functional <hist, target> that stays
        # finite where the Hellinger gradient is NaN (RBF underflows to exact zeros)
        (grad,) = torch.autograd.grad(loss, xr, retain_graph=True)
        (grad_lin,) = torch.autograd.grad((hist * target).sum(), xr)
    else:
        grad = torch.zeros_like(x)
        grad_lin = torch.zeros_like(x)
    return hist.detach(), loss.detach(), grad, grad_lin


def make_hist_goldens(only=None):
    for name, maker, B, S, Cc, kwargs, alpha in HIST_CASES:
        if only and name not in only:
            continue
        x = maker(B, S, seed=0, C=Cc)
        h = kwargs.get("h", 64)
        nc = 1 if kwargs.get("green_only", False) else 3
        target = ho.synth_random_target(B, h=h, seed=1, nc=nc)
        hist, loss, grad, grad_lin = run_reference_hist(x, target, alpha, kwargs)
        np.savez_compressed(
            os.path.join(GOLDEN_DIR, name + ".npz"),
            x=x.numpy(), target=target.numpy(), hist=hist.numpy(),
            loss=np.float32(loss.item()), grad_x=grad.numpy(), grad_x_lin=grad_lin.numpy(), alpha=np.float32(alpha),
            kwargs=json.dumps(kwargs))
        print(f"{name}: hist {tuple(hist.shape)} loss {loss.item():.7f}")


CHROMA_CASES = [
    ("chroma_default_2x48", ho.synth_generator_like, 2, 48, 3, dict()),
    ("chroma_interp_intensity_2x160", ho.synth_signed, 2, 160, 4, dict(insz=100, intensity_scale=True, h=32)),
    ("chroma_rbf_sampling_2x160", ho.synth_uniform, 2, 160, 3, dict(insz=100, resizing="sampling", method="RBF", sigma=0.05)),
]


LAB_CASES = [
    ("lab_default_2x48", ho.synth_uniform, 2, 48, 3, dict()),
    ("lab_interp_intensity_2x160", ho.synth_signed, 2, 160, 4, dict(insz=100, intensity_scale=True, h=32)),
    ("lab_rbf_sampling_2x160", ho.synth_generator_like, 2, 160, 3, dict(insz=100, resizing="sampling", method="RBF", sigma=0.05)),
]


def make_chroma_goldens():
    """histogram_classes/{rgChromaHistBlock,LabHistBlock}.py run unmodified on CPU."""
    import importlib
    ref_shim._ensure_path()
    _one_channel_goldens(importlib.import_module("histogram_classes.rgChromaHistBlock").rgChromaHistBlock,
                         CHROMA_CASES)
    _one_channel_goldens(importlib.import_module("histogram_classes.LabHistBlock").LabHistBlock, LAB_CASES)


def _one_channel_goldens(cls, cases):
    for name, maker, B, S, Cc, kwargs in cases:
        x = maker(B, S, seed=0, C=Cc)
        h = kwargs.get("h", 64)
        w = ho.synth_random_target(B, h=h, seed=1, nc=1)
        xr = x.clone().requires_grad_(True)
        hist = cls(device="cpu", **kwargs)(xr)
        (grad,) = torch.autograd.grad((hist * w).sum(), xr)
        np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), x=x.numpy(), target=w.numpy(),
                            hist=hist.detach().numpy(), grad_x_lin=grad.numpy(),
                            loss=np.float32(0), alpha=np.float32(0), kwargs=json.dumps(kwargs))
        print(f"{name}: hist {tuple(hist.shape)}")


def main():
    assert ref_shim.available(), "reference not mounted; goldens can only be made in the build container"
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    make_hist_goldens()
    make_chroma_goldens()
    try:
        from . import make_golden_gan
        make_golden_gan.main()
    except ImportError:
        pass


if __name__ == "__main__":
    main()
