"""Shared G/D parity checks: run the histogan_b200 modules against the reference's
golden vectors and return every relative error (so that GPU and CPU-emulation runs
report the same table)."""
import json
import os

import numpy as np
import torch

from oracle import gan_oracle as go
from oracle import make_golden_gan as mg
from tests import parity


def load(name):
    z = np.load(os.path.join(parity.GOLDEN_DIR, name))
    return {k: z[k] for k in z.files}


def rel(a, b):
    a = torch.as_tensor(np.asarray(a.detach().cpu() if torch.is_tensor(a) else a)).double()
    b = torch.as_tensor(np.asarray(b.detach().cpu() if torch.is_tensor(b) else b)).double()
    return ((a - b).norm() / b.norm().clamp_min(1e-300)).item()


def generator_errors(device):
    from histogan_b200.gan import Generator
    g = load("gan_generator_32.npz")
    G = Generator(mg.IMAGE_SIZE, mg.LATENT, network_capacity=mg.CAPACITY)
    shapes = json.loads(str(g["shapes"]))
    assert {k: list(v.shape) for k, v in G.state_dict().items()} == shapes   # drop-in state_dict
    G.load_state_dict(go.seeded_state_dict(shapes, seed=1))
    G.to(device)
    inp = {k: v.to(device) for k, v in mg.gan_inputs().items()}
    styles = inp["styles"].clone().requires_grad_(True)
    hists = inp["hists"].clone().requires_grad_(True)
    acts = []
    hooks = [b.register_forward_hook(lambda m, i, o: acts.append(o[0].detach())) for b in G.blocks]
    rgb = G(styles, hists, inp["noise"])
    for h in hooks:
        h.remove()
    assert rgb.shape == (mg.B, 3, mg.IMAGE_SIZE, mg.IMAGE_SIZE)
    e = {"rgb": rel(rgb, g["rgb"]), "act_last": rel(acts[-1], g["act_last"]),
         "act_norms": float(np.max(np.abs(np.array([a.norm().item() for a in acts]) / g["act_norms"] - 1)))}
    loss = (rgb * inp["w_rgb"]).sum()
    e["loss"] = abs(loss.item() - float(g["loss"])) / abs(float(g["loss"]))
    loss.backward()
    e["g_styles"] = rel(styles.grad, g["g_styles"])
    e["g_hists"] = rel(hists.grad, g["g_hists"])
    e["g_initial_block"] = rel(G.initial_block.grad, g["g_initial_block"])
    e["g_conv1_w_b1"] = rel(G.blocks[1].conv1.weight.grad[:8, :8], g["g_conv1_w_b1"])
    e["g_rgb_w_b3"] = rel(G.blocks[3].to_rgb.conv.weight.grad, g["g_rgb_w_b3"])
    norms = json.loads(str(g["param_grad_norms"]))
    e["param_grad_norms_max"] = max(abs(p.grad.norm().item() - norms[k]) / max(norms[k], 1e-6)
                                    for k, p in G.named_parameters())
    outs = {"rgb": rgb, "act_last": acts[-1], "g_styles": styles.grad, "g_hists": hists.grad}
    outs.update({"grad:" + k: p.grad for k, p in G.named_parameters()})
    return e, {k: v.detach().float().cpu() for k, v in outs.items()}


def discriminator_errors(device):
    from histogan_b200.gan import Discriminator
    from histogan_b200.trainer import gradient_penalty
    g = load("gan_discriminator_32.npz")
    D = Discriminator(mg.IMAGE_SIZE, network_capacity=mg.CAPACITY)
    shapes = json.loads(str(g["shapes"]))
    assert {k: list(v.shape) for k, v in D.state_dict().items()} == shapes
    D.load_state_dict(go.seeded_state_dict(shapes, seed=2))
    D.to(device)
    images = mg.gan_inputs()["images"].to(device).requires_grad_(True)
    logits, q = D(images)
    assert logits.shape == (mg.B,) and q.shape == (1,)
    e = {"logits": rel(logits, g["logits"])}
    gp = gradient_penalty(images, logits)          # histoGAN/histoGAN.py:156-163 (double backward)
    e["gp"] = abs(gp.item() - float(g["gp"])) / float(g["gp"])
    (logits.sum() + gp * mg.GP_WEIGHT_IN_TEST).backward()
    e["g_images"] = rel(images.grad, g["g_images"])
    e["g_to_logit_w"] = rel(D.to_logit.weight.grad, g["g_to_logit_w"])
    e["g_b0_net0_w"] = rel(D.blocks[0].net[0].weight.grad, g["g_b0_net0_w"])
    e["g_b2_down_w"] = rel(D.blocks[2].downsample.weight.grad[:8, :8], g["g_b2_down_w"])
    norms = json.loads(str(g["param_grad_norms"]))
    e["param_grad_norms_max"] = max(abs(p.grad.norm().item() - norms[k]) / max(norms[k], 1e-6)
                                    for k, p in D.named_parameters())
    outs = {"logits": logits, "gp": gp.reshape(1), "g_images": images.grad}
    outs.update({"grad:" + k: p.grad for k, p in D.named_parameters()})
    return e, {k: v.detach().float().cpu() for k, v in outs.items()}


def max_rel_between(a: dict, b: dict):
    """largest Frobenius-relative difference over the common entries of two output dicts"""
    worst, key = 0.0, None
    table = {}
    for k in a:
        r = rel(a[k], b[k])
        table[k] = r
        if r > worst:
            worst, key = r, k
    print("per-tensor relative difference:", {k: f"{v:.1e}" for k, v in sorted(table.items(), key=lambda kv: -kv[1])})
    return worst, key


# ------------------------------------------------------- the benchmarked shape (256^2) ----

def _fingerprint_table(named_params, g):
    from tests import step_checks as sc
    names = json.loads(str(g["grad_names"]))
    params = dict(named_params)
    return sc.compare_grads({k: params[k].grad for k in names}, names, g["grad_norms"], g["grad_samples"])


def generator_errors_256(device):
    """Generator at image 256 / capacity 16 (7 blocks, 64 -> 2048 -> ... -> 32 channels: the
    resident-filter, column-halo wgrad and split-K paths) against the reference golden."""
    from histogan_b200.gan import Generator
    from tests import step_checks as sc
    g = load("gan_generator_256.npz")
    G = Generator(mg.IMAGE_SIZE_L, mg.LATENT, network_capacity=mg.CAPACITY)
    shapes = json.loads(str(g["shapes"]))
    assert {k: list(v.shape) for k, v in G.state_dict().items()} == shapes
    G.load_state_dict(go.seeded_state_dict(shapes, seed=1))
    G.to(device)
    inp = {k: v.to(device) for k, v in mg.gan_inputs(mg.IMAGE_SIZE_L, mg.B_L, seed=5).items()}
    styles = inp["styles"].clone().requires_grad_(True)
    hists = inp["hists"].clone().requires_grad_(True)
    acts = []
    hooks = [b.register_forward_hook(lambda m, i, o: acts.append(o[0].detach())) for b in G.blocks]
    rgb = G(styles, hists, inp["noise"])
    for h in hooks:
        h.remove()
    e = {"rgb": rel(rgb, g["rgb"]),
         "rgb_max_abs": float((rgb.detach().cpu() - torch.from_numpy(g["rgb"])).abs().max() /
                              np.abs(g["rgb"]).max()),
         "act_norms": float(np.max(np.abs(np.array([a.double().norm().item() for a in acts]) / g["act_norms"] - 1))),
         "act_samples": max(rel(mg.strided(a.cpu(), mg.ACT_SAMPLES), s) for a, s in zip(acts, g["act_samples"]))}
    loss = (rgb * inp["w_rgb"]).sum()
    e["loss"] = abs(loss.item() - float(g["loss"])) / abs(float(g["loss"]))
    loss.backward()
    e["g_styles"] = rel(styles.grad, g["g_styles"])
    e["g_hists"] = rel(hists.grad, g["g_hists"])
    return e, _fingerprint_table(G.named_parameters(), g)


def discriminator_errors_256(device):
    """returns (scalar errors, first-order gradient table, adversarial + gradient-penalty table)"""
    from histogan_b200.gan import Discriminator
    from histogan_b200.trainer import gradient_penalty
    from tests import step_checks as sc
    g = load("gan_discriminator_256.npz")
    D = Discriminator(mg.IMAGE_SIZE_L, network_capacity=mg.CAPACITY)
    shapes = json.loads(str(g["shapes"]))
    assert {k: list(v.shape) for k, v in D.state_dict().items()} == shapes
    D.load_state_dict(go.seeded_state_dict(shapes, seed=2))
    D.to(device)
    images = mg.gan_inputs(mg.IMAGE_SIZE_L, mg.B_L, seed=5)["images"].to(device).requires_grad_(True)
    logits, _ = D(images)
    e = {"logits": rel(logits, g["logits"])}
    names = json.loads(str(g["grad_names"]))
    params = dict(D.named_parameters())
    # (i) first order
    logits.sum().backward(retain_graph=True)
    e["g1_images_norm"] = abs(images.grad.double().norm().item() - float(g["g1_images_norm"])) / float(g["g1_images_norm"])
    e["g1_images_samples"] = rel(mg.strided(images.grad.cpu(), 65536), g["g1_images_samples"])
    t1 = sc.compare_grads({k: params[k].grad for k in names}, names, g["grad1_norms"], g["grad1_samples"])
    D.zero_grad()
    images.grad = None
    # (ii) + gradient penalty (histoGAN.py:156-163: second order through every conv)
    gp = gradient_penalty(images, logits)
    e["gp"] = abs(gp.item() - float(g["gp"])) / float(g["gp"])
    (logits.sum() + gp * mg.GP_WEIGHT_IN_TEST).backward()
    e["g_images_norm"] = abs(images.grad.double().norm().item() - float(g["g_images_norm"])) / float(g["g_images_norm"])
    e["g_images_samples"] = rel(mg.strided(images.grad.cpu(), 65536), g["g_images_samples"])
    t2 = sc.compare_grads({k: params[k].grad for k in names}, names, g["grad_norms"], g["grad_samples"])
    return e, t1, t2
