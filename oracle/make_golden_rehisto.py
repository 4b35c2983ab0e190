"""Golden vectors for the ReHistoGAN generator phase, produced by the UNMODIFIED reference
classes (RecoloringEncoderDecoder, RecoloringGAN, HistVectorizer, Discriminator,
reconstruction_loss, get_gaussian_kernel, RGBuvHistBlock) on CPU.  TEST INFRASTRUCTURE.

    python -m oracle.make_golden_rehisto

The loss expression is the one of recoloringTrainer.train (ReHistoGAN/rehistoGAN.py:1003-1026)
typed out around the reference objects (train() itself needs CUDA + torch_optimizer).
Weights are regenerated from the saved shape tables by gan_oracle.seeded_state_dict."""
from __future__ import annotations

import json
import os

import numpy as np
import torch
import torch.nn.functional as F

from . import gan_oracle as go
from . import hist_oracle as ho
from . import ref_shim

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                          "tests", "golden")
IMAGE_SIZE, CAPACITY, LATENT, B = 64, 16, 512, 2
ALPHA, BETA, GAMMA = 32.0, 1.5, 4.0e-3     # gamma scaled so that the four loss terms are comparable
                                             # with the seeded (untrained) discriminator
HIST_KW = dict(h=64, insz=150, resizing="sampling", method="inverse-quadratic", sigma=0.02)
SEEDS = dict(ED=11, H=12, G=13, D=14)
GRAD_SLICE = 4096           # leading elements of every parameter gradient kept in the fixture


def rehisto_inputs(image_size=IMAGE_SIZE, b=B, seed=0):
    g = torch.Generator().manual_seed(4321 + seed)
    images = torch.rand(b, 3, image_size, image_size, generator=g)
    # smooth-ish images so that the laplacian / variance terms are not pure noise
    images = F.avg_pool2d(F.pad(images, (1, 1, 1, 1), mode="replicate"), 3, stride=1)
    return dict(images=images,
                hists=ho.synth_random_target(b, h=64, seed=5 + seed, nc=3),
                noise=torch.rand(b, image_size, image_size, 1, generator=g))


def build_state_dicts(mods: dict):
    """{name: module} -> ({name: shapes}, {name: seeded state dict}); loads them in place"""
    shapes, sds = {}, {}
    for name, m in mods.items():
        shapes[name] = {k: list(v.shape) for k, v in m.state_dict().items()}
        sds[name] = go.seeded_state_dict(shapes[name], seed=SEEDS[name])
        m.load_state_dict(sds[name])
    return shapes, sds


def main():
    assert ref_shim.available()
    gm = ref_shim.ref_gan_module()
    rm = ref_shim.ref_rehisto_module()
    hm = ref_shim.ref_hist_module()
    torch.set_num_threads(os.cpu_count() or 1)
    mods = dict(ED=rm.RecoloringEncoderDecoder(IMAGE_SIZE, network_capacity=CAPACITY, skip_conn_to_GAN=True),
                H=gm.HistVectorizer(64, LATENT, 8),
                G=rm.RecoloringGAN(IMAGE_SIZE, LATENT, CAPACITY),
                D=gm.Discriminator(IMAGE_SIZE, network_capacity=CAPACITY))
    shapes, _ = build_state_dicts(mods)
    inp = rehisto_inputs()
    images, hist_batch, noise = inp["images"], inp["hists"], inp["noise"]
    hist_block = hm.RGBuvHistBlock(device="cpu", **HIST_KW)
    hist_block_in = hm.RGBuvHistBlock(device="cpu", **HIST_KW)
    saved = torch.cuda.current_device
    torch.cuda.current_device = lambda: "cpu"           # reconstruction_loss builds its stencils there
    try:
        rec = rm.reconstruction_loss("2nd gradient")
        rec_sobel = rm.reconstruction_loss("1st gradient")
    finally:
        torch.cuda.current_device = saved
    gauss = rm.get_gaussian_kernel(kernel_size=15, sigma=5, channels=3)

    # ---- generator phase, rehistoGAN.py:981-1030 with skip_conn_to_GAN and not internal_hist
    h_w_space = mods["H"](hist_batch)
    image_latent, rgb, processed_latent_2, processed_latent_1 = mods["ED"](images, hist_batch)
    generated = mods["G"](image_latent, rgb, h_w_space, noise, processed_latent_2, processed_latent_1)
    fake_output, _ = mods["D"](generated)
    d_loss = GAMMA * fake_output.mean()
    generated_histograms = hist_block(F.relu(generated))
    histogram_loss = ALPHA * ho.SCALE * (torch.sqrt(torch.sum(torch.pow(
        torch.sqrt(hist_batch) - torch.sqrt(generated_histograms), 2)))) / hist_batch.shape[0]
    reconstruction = BETA * rec.compute_loss(images, generated)
    input_histograms = hist_block_in(F.relu(hist_batch))
    input_gauss = rm.gaussian_op(images, kernel=gauss)
    generated_gauss = rm.gaussian_op(generated, kernel=gauss)
    var_loss = -1 * (BETA / 10) * torch.sum(torch.abs(hist_batch - input_histograms)) * torch.mean(
        torch.abs(torch.std(torch.std(input_gauss, dim=2), dim=2) -
                  torch.std(torch.std(generated_gauss, dim=2), dim=2)))
    gen_loss = d_loss + histogram_loss + reconstruction + var_loss
    per_term = {k: torch.autograd.grad(v, generated, retain_graph=True)[0].numpy()
                for k, v in dict(d=d_loss, hist=histogram_loss, rec=reconstruction, var=var_loss).items()}
    gen_loss.backward()

    out = dict(shapes=json.dumps(shapes),
               latent=image_latent.detach().numpy(), p1=processed_latent_2.detach().numpy(),
               p2=processed_latent_1.detach().numpy(), ed_rgb=rgb.detach().numpy(),
               generated=generated.detach().numpy(),
               d_loss=np.float64(d_loss.item()), hist_loss=np.float64(histogram_loss.item()),
               rec_loss=np.float64(reconstruction.item()), var_loss=np.float64(var_loss.item()),
               rec_loss_sobel=np.float64((BETA * rec_sobel.compute_loss(images, generated)).item()),
               gen_loss=np.float64(gen_loss.item()))
    out.update({f"dgen_{k}": v for k, v in per_term.items()})
    norms = {}
    for name in ("ED", "H", "G"):
        for k, p in mods[name].named_parameters():
            if p.grad is None:
                norms[f"{name}.{k}"] = None
                continue
            norms[f"{name}.{k}"] = p.grad.norm().item()
            out[f"grad:{name}.{k}"] = p.grad.flatten()[:GRAD_SLICE].numpy()
    out["param_grad_norms"] = json.dumps(norms)
    np.savez_compressed(os.path.join(GOLDEN_DIR, "rehisto_64.npz"), **out)
    print("rehisto: generated", tuple(generated.shape), "losses d/h/r/v",
          d_loss.item(), histogram_loss.item(), reconstruction.item(), var_loss.item())
    print("params without gradient:", [k for k, v in norms.items() if v is None])


if __name__ == "__main__":
    main()
