"""CPU oracle for ONE optimisation step of HistoGAN -- TEST INFRASTRUCTURE.

Functional restatement of ``Trainer.train`` (histoGAN/histoGAN.py:853-989) on top of
the G/D/histogram oracles: which loss terms exist, how they are combined, which
parameters receive which gradient.

  d_phase   :887-932   hinge divergence (+ gradient penalty :919-922) -> grads of D
  g_phase   :934-989   fake_output.mean() + Hellinger histogram loss (:955-963)
                       (+ path-length regulariser :965-975)  -> grads of G, S, H
  draw_step_inputs     the random draws of one step in the reference's order
                       (:891-893, :936-937, :967; helpers :166-189)
  diffgrad_step        scalar restatement of the DiffGrad update (torch-optimizer's
                       ``DiffGrad``; Dubey et al., "diffGrad: An Optimization Method for
                       Convolutional Neural Networks", IEEE TNNLS 2019, Algorithm 1).
                       PARITY UNPINNED: the package is neither vendored by the reference
                       nor installed here; this follows the published
                       update rule and torch-optimizer 0.3.0's argument conventions.

Pinned: ``tests/golden/train_step_64.npz`` is produced by the UNMODIFIED reference
``Trainer.train`` (``oracle/make_golden_step.py``); ``tests/test_train_oracle.py``
checks these functions against it on the CPU.
"""
from __future__ import annotations

import math
from random import random

import numpy as np
import torch
import torch.nn.functional as F

from . import gan_oracle as go
from . import hist_oracle as ho

EPS = 1e-8            # histoGAN/histoGAN.py:53


def styles_tensor(sd_s, style_def):
    """latent_to_w + styles_def_to_tensor (:183-185, :215-217)"""
    return torch.cat([go.mlp(sd_s, "net", 8, z)[:, None, :].expand(-1, n, -1) for z, n in style_def],
                     dim=1)


def hist_latent(sd_h, hists):
    """H(hist) duplicated for the last two blocks (:900-902)"""
    hw = go.mlp(sd_h, "fcs", 8, hists.reshape(hists.shape[0], -1)).unsqueeze(1)
    return torch.cat((hw, hw), dim=1)


def gradient_penalty(images, output, weight=10):
    """:156-163"""
    (g,) = torch.autograd.grad(outputs=output, inputs=images, grad_outputs=torch.ones(output.size()),
                               create_graph=True, retain_graph=True, only_inputs=True)
    g = g.reshape(images.shape[0], -1)
    return weight * ((g.norm(2, dim=1) - 1) ** 2).mean()


def d_phase(sd_g, sd_d, sd_s, sd_h, images, hists, style_def, inoise, image_size, apply_gp):
    """returns dict(divergence, gp, grads={name: d loss/d D-param})"""
    images = images.detach().clone().requires_grad_(True)
    with torch.no_grad():              # the reference detaches the fake batch (:910)
        fake = go.generator(sd_g, styles_tensor(sd_s, style_def), hist_latent(sd_h, hists), inoise,
                            image_size)
    fake_out = go.discriminator(sd_d, fake, image_size)
    real_out = go.discriminator(sd_d, images, image_size)
    divergence = (F.relu(1 + real_out) + F.relu(1 - fake_out)).mean()            # :913
    loss, gp = divergence, None
    if apply_gp:
        gp = gradient_penalty(images, real_out)                                  # :919-922
        loss = loss + gp
    names = list(sd_d)
    grads = torch.autograd.grad(loss, [sd_d[k] for k in names], allow_unused=True)
    return {"divergence": divergence.detach(), "gp": gp.detach() if gp is not None else None,
            "grads": dict(zip(names, grads))}


def g_phase(sd_g, sd_d, sd_s, sd_h, hists, style_def, inoise, image_size, alpha=2.0,
            hist_kw=None, pl_noise=None, pl_mean=0):
    """returns dict(loss, hist_loss, avg_pl, grads={'G.'|'S.'|'H.' + name: ...})"""
    hist_kw = hist_kw or {}
    hists = hists.detach().clone().requires_grad_(True)                          # :940
    hw = hist_latent(sd_h, hists)
    w_styles = styles_tensor(sd_s, style_def)
    fake = go.generator(sd_g, w_styles, hw, inoise, image_size)
    fake_out = go.discriminator(sd_d, fake, image_size)
    gen_hist = ho.rgb_uv_hist(F.relu(fake), **hist_kw)                           # :955
    hist_loss = ho.hellinger_loss(hists, gen_hist, alpha)                        # :957-960
    loss = fake_out.mean()                                                       # :962
    gen_loss = loss + hist_loss
    avg_pl = None
    if pl_noise is not None:                                                     # :965-975
        std = 0.1 / (w_styles.std(dim=0, keepdim=True) + EPS)
        w2 = w_styles + pl_noise / (std + EPS)
        pl_images = go.generator(sd_g, w2, hw, inoise, image_size)
        pl_lengths = ((pl_images - fake) ** 2).mean(dim=(1, 2, 3))
        avg_pl = float(np.mean(pl_lengths.detach().numpy()))
        if pl_mean is not None:
            pl_loss = ((pl_lengths - pl_mean) ** 2).mean()
            if not torch.isnan(pl_loss):
                gen_loss = gen_loss + pl_loss
    named = [("G." + k, v) for k, v in sd_g.items()] + [("S." + k, v) for k, v in sd_s.items()] + \
            [("H." + k, v) for k, v in sd_h.items()]
    grads = torch.autograd.grad(gen_loss, [v for _, v in named], allow_unused=True)
    return {"loss": loss.detach(), "hist_loss": hist_loss.detach(), "avg_pl": avg_pl,
            "grads": {k: g for (k, _), g in zip(named, grads)}}


def draw_step_inputs(batch, layers, latent_dim, image_size, mixed_prob=0.9, path_penalty=False):
    """the random draws of one ``Trainer.train`` call, in the reference's order, from the
    global CPU generators (``random`` and torch): D-phase style + noise, G-phase style +
    noise, then the path-length perturbation.  `layers` = num_layers - 2."""
    def noise_list(n):
        return [(torch.randn(batch, latent_dim), n)]

    def mixed_list(n):
        tt = int(torch.rand(()).numpy() * n)
        return noise_list(tt) + noise_list(n - tt)

    def image_noise():
        return torch.FloatTensor(batch, image_size, image_size, 1).uniform_(0.0, 1.0)

    fn = mixed_list if random() < mixed_prob else noise_list        # drawn once per step (:891)
    out = {"d_style": fn(layers), "d_noise": image_noise()}
    out["g_style"] = fn(layers)
    out["g_noise"] = image_noise()
    out["pl_noise"] = torch.randn(batch, layers, latent_dim) if path_penalty else None
    return out


# ------------------------------------------------------------------ DiffGrad --

def diffgrad_step(p, g, state, lr=2e-4, betas=(0.5, 0.9), eps=1e-8, weight_decay=0.0):
    """one DiffGrad update of python-float lists, element by element (no vector code, so
    that it is an independent check of the fused kernel):

        m_t = b1 m + (1-b1) g ;  v_t = b2 v + (1-b2) g^2
        xi  = sigmoid(|g_{t-1} - g_t|)                       (the "friction" coefficient)
        p  -= lr * sqrt(1-b2^t)/(1-b1^t) * (m_t * xi) / (sqrt(v_t) + eps)

    `state` = dict(step, m, v, g_prev); returns the new parameter list."""
    b1, b2 = betas
    state["step"] = t = state.get("step", 0) + 1
    m = state.setdefault("m", [0.0] * len(p))
    v = state.setdefault("v", [0.0] * len(p))
    gp = state.setdefault("g_prev", [0.0] * len(p))
    step_size = lr * math.sqrt(1 - b2 ** t) / (1 - b1 ** t)
    out = []
    for i in range(len(p)):
        gi = g[i] + weight_decay * p[i]
        m[i] = b1 * m[i] + (1 - b1) * gi
        v[i] = b2 * v[i] + (1 - b2) * gi * gi
        xi = 1.0 / (1.0 + math.exp(-abs(gp[i] - gi)))
        gp[i] = gi
        out.append(p[i] - step_size * m[i] * xi / (math.sqrt(v[i]) + eps))
    return out


@torch.no_grad()
def diffgrad_step_tensors(params, grads, state, lr=2e-4, betas=(0.5, 0.9), eps=1e-8):
    """the same update, vectorised over torch tensors (CPU-baseline leg of bench.py: the scalar
    form above would take minutes for 190 M parameters)"""
    b1, b2 = betas
    state["step"] = t = state.get("step", 0) + 1
    step_size = lr * math.sqrt(1 - b2 ** t) / (1 - b1 ** t)
    for i, (p, g) in enumerate(zip(params, grads)):
        if g is None:
            continue
        m, v, gp = state.setdefault(("m", i), torch.zeros_like(p)), \
            state.setdefault(("v", i), torch.zeros_like(p)), state.setdefault(("gp", i), torch.zeros_like(p))
        m.mul_(b1).add_(g, alpha=1 - b1)
        v.mul_(b2).addcmul_(g, g, value=1 - b2)
        xi = torch.sigmoid((gp - g).abs())
        gp.copy_(g)
        p.addcdiv_(m * xi, v.sqrt().add_(eps), value=-step_size)
