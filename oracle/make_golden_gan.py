"""Golden vectors for the generator / discriminator stacks, produced by the
UNMODIFIED reference classes (stub-import shim) on CPU.  TEST INFRASTRUCTURE.

    python -m oracle.make_golden_gan

Weights are not stored: gan_oracle.seeded_state_dict regenerates them
deterministically from the parameter-shape table saved in the fixture."""
from __future__ import annotations

import json
import os

import numpy as np
import torch

from . import gan_oracle as go
from . import ref_shim

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                          "tests", "golden")
IMAGE_SIZE, CAPACITY, LATENT, B = 32, 16, 512, 2
GP_WEIGHT_IN_TEST = 1e-3


def gan_inputs(image_size=IMAGE_SIZE, b=B, seed=0):
    g = torch.Generator().manual_seed(1234 + seed)
    num_layers = int(np.log2(image_size) - 1)
    return dict(
        styles=torch.randn(b, num_layers - 2, LATENT, generator=g),
        hists=torch.randn(b, 2, LATENT, generator=g),
        noise=torch.rand(b, image_size, image_size, 1, generator=g),
        images=torch.rand(b, 3, image_size, image_size, generator=g),
        w_rgb=torch.randn(b, 3, image_size, image_size, generator=g),
    )


def gradient_penalty(images, output, weight=10):
    """histoGAN/histoGAN.py:156-163 without the .cuda() calls."""
    (gradients,) = torch.autograd.grad(outputs=output, inputs=images,
                                       grad_outputs=torch.ones(output.size()),
                                       create_graph=True, retain_graph=True, only_inputs=True)
    gradients = gradients.reshape(images.shape[0], -1)
    return weight * ((gradients.norm(2, dim=1) - 1) ** 2).mean()


def main():
    assert ref_shim.available()
    gm = ref_shim.ref_gan_module()
    torch.set_num_threads(os.cpu_count() or 1)
    inp = gan_inputs()

    # ---------------- generator
    G = gm.Generator(IMAGE_SIZE, LATENT, network_capacity=CAPACITY)
    shapes = {k: list(v.shape) for k, v in G.state_dict().items()}
    G.load_state_dict(go.seeded_state_dict(shapes, seed=1))
    styles = inp["styles"].clone().requires_grad_(True)
    hists = inp["hists"].clone().requires_grad_(True)
    acts = []
    hooks = [blk.register_forward_hook(lambda m, i, o: acts.append(o[0].detach())) for blk in G.blocks]
    rgb = G(styles, hists, inp["noise"])
    for h in hooks:
        h.remove()
    loss = (rgb * inp["w_rgb"]).sum()
    loss.backward()
    out = dict(
        shapes=json.dumps(shapes), rgb=rgb.detach().numpy(),
        act_norms=np.array([a.norm().item() for a in acts], dtype=np.float64),
        act_last=acts[-1].numpy(),
        loss=np.float64(loss.item()), g_styles=styles.grad.numpy(), g_hists=hists.grad.numpy(),
        g_initial_block=G.initial_block.grad.numpy(),
        g_conv1_w_b1=G.blocks[1].conv1.weight.grad.numpy()[:8, :8],
        g_rgb_w_b3=G.blocks[3].to_rgb.conv.weight.grad.numpy(),
        param_grad_norms=json.dumps({k: p.grad.norm().item() for k, p in G.named_parameters()}),
    )
    np.savez_compressed(os.path.join(GOLDEN_DIR, "gan_generator_32.npz"), **out)
    print("generator: rgb", tuple(rgb.shape), "loss", loss.item())

    # ---------------- discriminator (+ gradient penalty = double backward)
    D = gm.Discriminator(IMAGE_SIZE, network_capacity=CAPACITY)
    dshapes = {k: list(v.shape) for k, v in D.state_dict().items()}
    D.load_state_dict(go.seeded_state_dict(dshapes, seed=2))
    images = inp["images"].clone().requires_grad_(True)
    logits, q = D(images)
    gp = gradient_penalty(images, logits)
    # both terms active: the adversarial term (first order) and the penalty (second order)
    (logits.sum() + gp * GP_WEIGHT_IN_TEST).backward()
    dout = dict(
        shapes=json.dumps(dshapes), logits=logits.detach().numpy(), gp=np.float64(gp.item()),
        g_images=images.grad.numpy(),
        g_to_logit_w=D.to_logit.weight.grad.numpy(),
        g_b0_net0_w=D.blocks[0].net[0].weight.grad.numpy(),
        g_b2_down_w=D.blocks[2].downsample.weight.grad.numpy()[:8, :8],
        param_grad_norms=json.dumps({k: p.grad.norm().item() for k, p in D.named_parameters()}),
    )
    np.savez_compressed(os.path.join(GOLDEN_DIR, "gan_discriminator_32.npz"), **dout)
    print("discriminator: logits", logits.detach().numpy(), "gp", gp.item())


# ---------------------------------------------------------------------------------------
# The BENCHMARKED shape: image 256, capacity 16 (7 generator blocks 64->2048->...->32, 8
# discriminator blocks 3->16->...->2048), batch 2.  Large tensors are stored as fingerprints
# (float64 norm + evenly strided entries, make_golden_step.fingerprint) so the fixture stays
# ~2 MB: full `rgb`, per-block activation norms / samples, every parameter gradient.

IMAGE_SIZE_L, B_L = 256, 2
ACT_SAMPLES = 4096


def strided(t, n):
    flat = t.detach().contiguous().reshape(-1)
    idx = (torch.arange(min(n, flat.numel()), dtype=torch.int64) * flat.numel()) // min(n, flat.numel())
    return flat[idx].float().numpy()


def grad_fingerprints(named_params):
    from .make_golden_step import fingerprint
    names, norms, samples = [], [], []
    for k, p in named_params:
        n, s = fingerprint(p.grad)
        names.append(k); norms.append(n); samples.append(s)
    return json.dumps(names), np.array(norms, dtype=np.float64), np.concatenate(samples)


def main_256():
    assert ref_shim.available()
    gm = ref_shim.ref_gan_module()
    torch.set_num_threads(os.cpu_count() or 1)
    inp = gan_inputs(IMAGE_SIZE_L, B_L, seed=5)

    G = gm.Generator(IMAGE_SIZE_L, LATENT, network_capacity=CAPACITY)
    shapes = {k: list(v.shape) for k, v in G.state_dict().items()}
    G.load_state_dict(go.seeded_state_dict(shapes, seed=1))
    styles = inp["styles"].clone().requires_grad_(True)
    hists = inp["hists"].clone().requires_grad_(True)
    acts = []
    hooks = [blk.register_forward_hook(lambda m, i, o: acts.append(o[0].detach())) for blk in G.blocks]
    rgb = G(styles, hists, inp["noise"])
    for h in hooks:
        h.remove()
    loss = (rgb * inp["w_rgb"]).sum()
    loss.backward()
    names, norms, samples = grad_fingerprints(G.named_parameters())
    out = dict(shapes=json.dumps(shapes), rgb=rgb.detach().numpy(),
               act_norms=np.array([a.double().norm().item() for a in acts]),
               act_samples=np.stack([strided(a, ACT_SAMPLES) for a in acts]),
               loss=np.float64(loss.item()), g_styles=styles.grad.numpy(), g_hists=hists.grad.numpy(),
               grad_names=names, grad_norms=norms, grad_samples=samples)
    np.savez_compressed(os.path.join(GOLDEN_DIR, "gan_generator_256.npz"), **out)
    print("generator 256: rgb", tuple(rgb.shape), "loss", loss.item(), "act norms", out["act_norms"])
    del G, acts, rgb

    D = gm.Discriminator(IMAGE_SIZE_L, network_capacity=CAPACITY)
    dshapes = {k: list(v.shape) for k, v in D.state_dict().items()}
    D.load_state_dict(go.seeded_state_dict(dshapes, seed=2))
    images = inp["images"].clone().requires_grad_(True)
    logits, q = D(images)
    # (i) first order only: d logits.sum() / d(parameters, images)
    logits.sum().backward(retain_graph=True)
    names1, norms1, samples1 = grad_fingerprints(D.named_parameters())
    g1_images = images.grad.clone()
    D.zero_grad()
    images.grad = None
    # (ii) both terms: the adversarial term + the gradient penalty (second order through every conv)
    gp = gradient_penalty(images, logits)
    (logits.sum() + gp * GP_WEIGHT_IN_TEST).backward()
    names, norms, samples = grad_fingerprints(D.named_parameters())
    dout = dict(shapes=json.dumps(dshapes), logits=logits.detach().numpy(), gp=np.float64(gp.item()),
                g_images_norm=np.float64(images.grad.double().norm().item()),
                g_images_samples=strided(images.grad, 65536),
                grad_names=names, grad_norms=norms, grad_samples=samples,
                g1_images_norm=np.float64(g1_images.double().norm().item()),
                g1_images_samples=strided(g1_images, 65536),
                grad1_norms=norms1, grad1_samples=samples1)
    np.savez_compressed(os.path.join(GOLDEN_DIR, "gan_discriminator_256.npz"), **dout)
    print("discriminator 256: logits", logits.detach().numpy(), "gp", gp.item())


if __name__ == "__main__":
    import sys
    if len(sys.argv) > 1 and sys.argv[1] == "256":
        main_256()
    else:
        main()
