"""Shared ReHistoGAN parity checks (CPU emulation and GPU): run the histogan_b200 recolouring
modules on the golden inputs and return relative errors + the raw outputs."""
import json
import os

import numpy as np
import torch
import torch.nn.functional as F

from oracle import gan_oracle as go
from oracle import hist_oracle as ho
from oracle import make_golden_rehisto as mr
from oracle import rehisto_oracle as ro
from tests import parity
from tests.gan_checks import rel


def golden():
    z = np.load(os.path.join(parity.GOLDEN_DIR, "rehisto_64.npz"))
    return {k: z[k] for k in z.files}


def is_prenorm_bias(key):
    return "encoder_blocks." in key and key.endswith((".net.0.bias", ".net.3.bias"))


def build_modules(device):
    from histogan_b200 import rehistogan as rh
    from histogan_b200.gan import Discriminator, HistVectorizer
    g = golden()
    shapes = json.loads(str(g["shapes"]))
    mods = dict(ED=rh.RecoloringEncoderDecoder(mr.IMAGE_SIZE, network_capacity=mr.CAPACITY,
                                               skip_conn_to_GAN=True),
                H=HistVectorizer(64, mr.LATENT, 8),
                G=rh.RecoloringGAN(mr.IMAGE_SIZE, mr.LATENT, mr.CAPACITY),
                D=Discriminator(mr.IMAGE_SIZE, network_capacity=mr.CAPACITY))
    for n, m in mods.items():
        assert {k: list(v.shape) for k, v in m.state_dict().items()} == shapes[n], n   # drop-in state_dict
        m.load_state_dict(go.seeded_state_dict(shapes[n], seed=mr.SEEDS[n]))
        m.to(device)
    return g, mods


def g_phase_errors(device, losses_fn):
    """losses_fn(images, hists, generated, D) -> dict(d_loss, hist_loss, rec_loss, var_loss)"""
    g, mods = build_modules(device)
    inp = {k: v.to(device) for k, v in mr.rehisto_inputs().items()}
    h_w = mods["H"](inp["hists"])
    latent, rgb, p1, p2 = mods["ED"](inp["images"], inp["hists"])
    gen = mods["G"](latent, rgb, h_w, inp["noise"], p1, p2)
    e = {"latent": rel(latent, g["latent"]), "p1": rel(p1, g["p1"]), "p2": rel(p2, g["p2"]),
         "ed_rgb": rel(rgb, g["ed_rgb"]), "generated": rel(gen, g["generated"])}
    L = losses_fn(inp["images"], inp["hists"], gen, mods["D"])
    for k in ("d_loss", "hist_loss", "rec_loss", "var_loss"):
        e[k] = abs(L[k].item() - float(g[k])) / abs(float(g[k]))
    outs = {"latent": latent, "p1": p1, "p2": p2, "generated": gen}
    # d loss_i / d generated.  The Hellinger term is ill-conditioned in `generated` (1/sqrt of
    # near-empty bins: a 1e-6 change of the image moves it by 1e-2), so every term is
    # differentiated AT THE REFERENCE's generated image ...
    gen_ref = torch.as_tensor(g["generated"]).to(device).requires_grad_(True)
    Lr = losses_fn(inp["images"], inp["hists"], gen_ref, mods["D"])
    upstream = 0
    for k, name in (("d", "d_loss"), ("hist", "hist_loss"), ("rec", "rec_loss"), ("var", "var_loss")):
        (dg,) = torch.autograd.grad(Lr[name], gen_ref, retain_graph=True)
        e["dgen_" + k] = rel(dg, g["dgen_" + k])
        outs["dgen_" + k] = dg
        upstream = upstream + torch.as_tensor(g["dgen_" + k]).to(device)
    # ... and the networks are back-propagated from the reference's d loss / d generated
    for m in mods.values():
        m.zero_grad()
    gen.backward(upstream)
    norms = json.loads(str(g["param_grad_norms"]))
    worst, worst_key = 0.0, None
    for name in ("ED", "H", "G"):
        for k, p in mods[name].named_parameters():
            ref = norms[f"{name}.{k}"]
            if ref is None:
                assert p.grad is None or float(p.grad.abs().max()) == 0.0, (name, k)
                continue
            if is_prenorm_bias(k):
                continue
            outs[f"grad:{name}.{k}"] = p.grad.flatten()[:mr.GRAD_SLICE]
            r = rel(p.grad.flatten()[:mr.GRAD_SLICE], g[f"grad:{name}.{k}"])
            if r > worst:
                worst, worst_key = r, f"{name}.{k}"
    e["param_grads_max"] = worst
    print("worst parameter gradient:", worst_key, worst)
    return e, {k: v.detach().float().cpu() for k, v in outs.items()}


def oracle_losses(images, hists, gen, D):
    """the four loss terms from CPU torch ops + the oracle histogram (for the emulation run)"""
    fake, _ = D(gen)
    return dict(d_loss=mr.GAMMA * fake.mean(),
                hist_loss=ho.hellinger_loss(hists, ho.rgb_uv_hist(F.relu(gen), **mr.HIST_KW), mr.ALPHA),
                rec_loss=mr.BETA * ro.reconstruction_loss(images, gen, "laplacian"),
                var_loss=ro.variance_loss(images, gen, hists, mr.BETA, mr.HIST_KW))
