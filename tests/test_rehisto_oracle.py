"""CPU: the ReHistoGAN oracle restatement against the reference-made golden vectors."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import gan_oracle as go
from oracle import make_golden_rehisto as mr
from oracle import rehisto_oracle as ro
from tests import parity


def _golden():
    z = np.load(os.path.join(parity.GOLDEN_DIR, "rehisto_64.npz"))
    return {k: z[k] for k in z.files}


def _rel(a, b):
    a, b = torch.as_tensor(np.asarray(a)).double(), torch.as_tensor(np.asarray(b)).double()
    return ((a - b).norm() / b.norm().clamp_min(1e-300)).item()


def is_prenorm_bias(key):
    """biases feeding an InstanceNorm (EncoderBlock.net.0 / net.3) cannot influence the output"""
    return key.startswith("encoder_blocks.") and key.endswith((".net.0.bias", ".net.3.bias"))


def test_rehisto_oracle_matches_reference_golden():
    g = _golden()
    shapes = json.loads(str(g["shapes"]))
    sds = {n: {k: v.clone().requires_grad_(v.is_floating_point())
               for k, v in go.seeded_state_dict(shapes[n], seed=mr.SEEDS[n]).items()} for n in shapes}
    inp = mr.rehisto_inputs()
    out = ro.g_phase(sds["ED"], sds["H"], sds["G"], sds["D"], inp["images"], inp["hists"], inp["noise"],
                     mr.IMAGE_SIZE, mr.ALPHA, mr.BETA, mr.GAMMA, hist_kw=mr.HIST_KW)
    for k, gk in (("latent", "latent"), ("p1", "p1"), ("p2", "p2"), ("rgb", "ed_rgb"),
                  ("generated", "generated")):
        assert _rel(out[k].detach(), g[gk]) < 2e-5, k
    for k in ("d_loss", "hist_loss", "rec_loss", "var_loss", "gen_loss"):
        assert abs(out[k].item() - float(g[k])) <= 2e-5 * abs(float(g[k])) + 1e-6, k
    sobel = mr.BETA * ro.reconstruction_loss(inp["images"], out["generated"].detach(), "sobel")
    assert abs(sobel.item() - float(g["rec_loss_sobel"])) <= 2e-5 * float(g["rec_loss_sobel"])
    for k in ("d", "hist", "rec", "var"):
        (dg,) = torch.autograd.grad(out[{"d": "d_loss", "hist": "hist_loss", "rec": "rec_loss",
                                         "var": "var_loss"}[k]], out["generated"], retain_graph=True)
        assert _rel(dg, g["dgen_" + k]) < 1e-4, k
    out["gen_loss"].backward()
    norms = json.loads(str(g["param_grad_norms"]))
    worst = 0.0
    for name in ("ED", "H", "G"):
        for k, v in sds[name].items():
            ref = norms.get(f"{name}.{k}")
            if ref is None:
                continue
            if is_prenorm_bias(k):      # exactly zero in exact arithmetic: rounding noise only
                wn = norms[f"{name}.{k[:-4]}weight"]
                assert ref < 1e-4 * wn and v.grad.norm().item() < 1e-4 * wn, (name, k)
                continue
            r = _rel(v.grad.flatten()[:mr.GRAD_SLICE], g[f"grad:{name}.{k}"])
            worst = max(worst, r)
            assert r < 2e-3, (name, k, r)
    print("worst parameter-gradient difference", worst)
