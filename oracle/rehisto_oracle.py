"""CPU oracle for the ReHistoGAN recolouring step -- TEST INFRASTRUCTURE.

Functional (state_dict in, tensors out) torch-CPU restatement of the reference's
``ReHistoGAN/rehistoGAN.py`` for the generator phase of ``recoloringTrainer.train``:

  encoder_block         EncoderBlock.forward               :499-504
  decoder_block         DecoderBlock.forward               :534-546  (internal_hist=False)
  encoder_decoder       RecoloringEncoderDecoder.forward   :603-634  (skip_conn_to_GAN=True)
  recoloring_head       RecoloringGAN.forward              :478-482
  gaussian_kernel       get_gaussian_kernel                :207-225
  reconstruction_loss   reconstruction_loss.compute_loss   :303-326
  variance_loss         the var_loss expression            :1016-1024
  g_phase               the generator half of train()      :981-1030

Only the product's tests / smoke / the bench's CPU arm may import this file.
Pinned against the unmodified reference classes by oracle/make_golden_rehisto.py
(tests/golden/rehisto_64.npz) and tests/test_rehisto_oracle.py.
"""
from __future__ import annotations

from math import log2, pi

import torch
import torch.nn.functional as F

from . import gan_oracle as go
from . import hist_oracle as ho


def _sub(sd, prefix):
    return {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}


def _conv(sd, p, x, stride=1, padding=0):
    return F.conv2d(x, sd[p + ".weight"], sd[p + ".bias"], stride=stride, padding=padding)


def _in_lrelu(x):
    return F.leaky_relu(F.instance_norm(x, eps=1e-5), 0.2)      # nn.InstanceNorm2d defaults


def encoder_block(sd, p, x):
    res = _conv(sd, p + ".conv_res", x)
    y = _in_lrelu(_conv(sd, p + ".net.0", x, padding=1))
    y = _in_lrelu(_conv(sd, p + ".net.3", y, padding=1))
    y = y + res
    return _conv(sd, p + ".downsample", y, stride=2, padding=1), y


def decoder_block(sd, p, x, prev_rgb, prev_latent):
    cur = F.leaky_relu(_conv(sd, p + ".block1.0", x, padding=1), 0.2)
    proc = F.leaky_relu(_conv(sd, p + ".block2.0", torch.cat((cur, prev_latent), dim=1), padding=1), 0.2)
    x = F.leaky_relu(_conv(sd, p + ".conv_out_latent.0", _conv(sd, p + ".conv_res", x) + proc,
                           padding=1), 0.2)
    rgb = _conv(sd, p + ".conv_out_rgb", x)
    if prev_rgb is not None:
        rgb = rgb + prev_rgb
    return go._up2(x), go._up2(rgb)


def encoder_decoder(sd, x, hists, image_size, style_depth=8):
    """skip_conn_to_GAN=True, internal_hist=False (the CLI defaults, rehistoGAN.py:528-533)
    -> (latent, rgb, processed_latent_1, processed_latent_2)."""
    n_enc, n_dec = int(log2(image_size) - 2), int(log2(image_size) - 4)
    hw = go.mlp(sd, "hist_projection.fcs", style_depth, hists.reshape(hists.shape[0], -1))
    h1 = go._linear(sd, "to_latent_1", hw)
    h2 = go._linear(sd, "to_latent_2", hw)
    x = _conv(sd, "mapping", x, padding=1)
    downs, ups = [], []
    for i in range(n_enc):
        x, xu = encoder_block(sd, f"encoder_blocks.{i}", x)
        downs.append(x)
        ups.append(xu)
    p1 = go.mod_conv(ups[1], h1, sd["conv_latent_1.weight"])
    p2 = go.mod_conv(ups[0], h2, sd["conv_latent_2.weight"])
    skips = downs[::-1][:-2]
    rgb = None
    for i in range(n_dec):
        x, rgb = decoder_block(sd, f"decoder_blocks.{i}", x, rgb, skips[i])
    return _conv(sd, "decoder_mapping", x), rgb, p1, p2


def _head_block(sd, p, x, prev_rgb, istyle, inoise, latent, upsample_rgb):
    """GeneratorBlock.forward with the `latent` skip (histoGAN/histoGAN.py:461-479)."""
    x = go._up2(x)
    nz = inoise[:, :x.shape[2], :x.shape[3], :]
    n1 = go._linear(sd, p + ".to_noise1", nz).permute(0, 3, 2, 1)
    n2 = go._linear(sd, p + ".to_noise2", nz).permute(0, 3, 2, 1)
    x = F.leaky_relu(go.mod_conv(x, go._linear(sd, p + ".to_style1", istyle), sd[p + ".conv1.weight"]) + n1, 0.2)
    if latent is not None:
        x = x + latent
    x = F.leaky_relu(go.mod_conv(x, go._linear(sd, p + ".to_style2", istyle), sd[p + ".conv2.weight"]) + n2, 0.2)
    return x, go.to_rgb(sd, p + ".to_rgb", x, prev_rgb, istyle, upsample_rgb)


def recoloring_head(sd, x, hists_w, noise, latent1, latent2):
    """RecoloringGAN.forward: the incoming rgb is discarded (:479)."""
    x, rgb = _head_block(sd, "blocks.0", x, None, hists_w, noise, latent1, True)
    x, rgb = _head_block(sd, "blocks.1", x, rgb, hists_w, noise, latent2, False)
    return rgb


def gaussian_kernel(kernel_size=15, sigma=3.0, channels=3):
    ax = torch.arange(kernel_size)
    xg = ax.repeat(kernel_size).view(kernel_size, kernel_size)
    xy = torch.stack([xg, xg.t()], dim=-1).float()
    mean, var = (kernel_size - 1) / 2., sigma ** 2.
    k = (1. / (2. * pi * var)) * torch.exp(-torch.sum((xy - mean) ** 2., dim=-1) / (2 * var))
    k = k / torch.sum(k)
    return k.view(1, 1, kernel_size, kernel_size).repeat(channels, 1, 1, 1)


_LAPLACIAN = [[0, 1, 0], [1, -4, 1], [0, 1, 0]]
_SOBEL_X = [[1, 0, -1], [2, 0, -2], [1, 0, -1]]
_SOBEL_Y = [[1, 2, 1], [0, 0, 0], [-1, -2, -1]]


def _stencil(x, k):
    w = torch.tensor(k, dtype=torch.float32).unsqueeze(0).expand(1, 3, 3, 3)
    return F.conv2d(x, w, stride=1, padding=1)


def reconstruction_loss(inp, target, kind="laplacian"):
    if kind is None:
        return torch.mean(torch.abs(inp - target))
    if kind == "sobel":
        gi = torch.sqrt(_stencil(inp, _SOBEL_X) ** 2 + _stencil(inp, _SOBEL_Y) ** 2)
        gt = torch.sqrt(_stencil(target, _SOBEL_X) ** 2 + _stencil(target, _SOBEL_Y) ** 2)
        return torch.mean(torch.abs(gi - gt))
    if kind == "laplacian":
        return torch.mean(torch.abs(_stencil(inp, _LAPLACIAN) - _stencil(target, _LAPLACIAN)))
    raise Exception("Unknown reconstruction losst!")


def variance_loss(images, generated, hist_batch, beta, hist_kw):
    input_h = ho.rgb_uv_hist(F.relu(hist_batch), **hist_kw)       # (sic) histogram OF the histogram
    k = gaussian_kernel(15, 5.0, 3)
    gi = F.conv2d(images, k, groups=3)
    gg = F.conv2d(generated, k, groups=3)
    return -1 * (beta / 10) * torch.sum(torch.abs(hist_batch - input_h)) * torch.mean(torch.abs(
        torch.std(torch.std(gi, dim=2), dim=2) - torch.std(torch.std(gg, dim=2), dim=2)))


def g_phase(sd_ed, sd_h, sd_g, sd_d, images, hist_batch, noise, image_size, alpha=32., beta=1.5,
            gamma=4., rec_loss="laplacian", variance=True, hist_kw=None, style_depth=8):
    """generator half of recoloringTrainer.train (:981-1030) -> dict of tensors."""
    hist_kw = dict(h=64, insz=150, resizing="sampling", method="inverse-quadratic", sigma=0.02) \
        if hist_kw is None else hist_kw
    hw = go.mlp(sd_h, "fcs", style_depth, hist_batch.reshape(hist_batch.shape[0], -1))
    latent, rgb, p1, p2 = encoder_decoder(sd_ed, images, hist_batch, image_size, style_depth)
    gen = recoloring_head(sd_g, latent, hw, noise, p1, p2)
    d_loss = gamma * go.discriminator(sd_d, gen, image_size).mean()
    gen_h = ho.rgb_uv_hist(F.relu(gen), **hist_kw)
    h_loss = ho.hellinger_loss(hist_batch, gen_h, alpha)
    r_loss = beta * reconstruction_loss(images, gen, rec_loss)
    out = dict(latent=latent, rgb=rgb, p1=p1, p2=p2, generated=gen, d_loss=d_loss, hist_loss=h_loss,
               rec_loss=r_loss)
    total = d_loss + h_loss + r_loss
    if variance:
        out["var_loss"] = variance_loss(images, gen, hist_batch, beta, hist_kw)
        total = total + out["var_loss"]
    out["gen_loss"] = total
    return out
