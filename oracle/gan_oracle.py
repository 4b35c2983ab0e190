"""CPU oracle for the generator / discriminator stacks -- TEST INFRASTRUCTURE.

Functional (state_dict in, tensors out) torch-CPU restatement of the reference
algorithms, evaluated the way the REFERENCE evaluates them (per-sample weights
+ grouped convolution), i.e. deliberately not the way the CUDA path does:

  mod_conv            Conv2DMod.forward           histoGAN/histoGAN.py:420-440
  generator_block     GeneratorBlock.forward      :461-479
  to_rgb              RGBBlock.forward            :380-390
  generator           Generator.forward           :558-568
  discriminator       Discriminator.forward       :613-631 (blocks :520-526)
  style_mlp/hist_mlp  StyleVectorizer/HistVectorizer :335-365

Pinned against the reference classes through oracle/ref_shim.py in
tests/test_gan_oracle_vs_reference.py and through tests/golden/gan_*.npz.
"""
from __future__ import annotations

from math import log2

import torch
import torch.nn.functional as F

EPS = 1e-8            # histoGAN/histoGAN.py:53


def mod_conv(x, y, weight, demod=True):
    """per-sample weights w*(y+1), optional demodulation, grouped conv, same padding."""
    b, c, h, w = x.shape
    cout, cin, k, _ = weight.shape
    wts = weight.unsqueeze(0) * (y.reshape(b, 1, cin, 1, 1) + 1)
    if demod:
        wts = wts * torch.rsqrt(wts.pow(2).sum(dim=(2, 3, 4), keepdim=True) + EPS)
    out = F.conv2d(x.reshape(1, b * c, h, w), wts.reshape(b * cout, cin, k, k),
                   padding=(k - 1) // 2, groups=b)
    return out.reshape(b, cout, h, w)


def _linear(sd, prefix, x):
    return F.linear(x, sd[prefix + ".weight"], sd[prefix + ".bias"])


def _up2(x):
    return F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False)


def to_rgb(sd, p, x, prev_rgb, istyle, upsample):
    rgb = mod_conv(x, _linear(sd, p + ".to_style", istyle), sd[p + ".conv.weight"], demod=False)
    if prev_rgb is not None:
        rgb = rgb + prev_rgb
    return _up2(rgb) if upsample else rgb


def generator_block(sd, p, x, prev_rgb, istyle, inoise, upsample, upsample_rgb):
    if upsample:
        x = _up2(x)
    nz = inoise[:, :x.shape[2], :x.shape[3], :]
    n1 = _linear(sd, p + ".to_noise1", nz).permute(0, 3, 2, 1)     # spatial transpose (:466)
    n2 = _linear(sd, p + ".to_noise2", nz).permute(0, 3, 2, 1)
    x = F.leaky_relu(mod_conv(x, _linear(sd, p + ".to_style1", istyle), sd[p + ".conv1.weight"]) + n1, 0.2)
    x = F.leaky_relu(mod_conv(x, _linear(sd, p + ".to_style2", istyle), sd[p + ".conv2.weight"]) + n2, 0.2)
    return x, to_rgb(sd, p + ".to_rgb", x, prev_rgb, istyle, upsample_rgb)


def generator(sd, styles, hists, noise, image_size, prefix="", return_all=False):
    """Generator.forward; sd holds 'initial_block', 'blocks.{i}....' (optionally prefixed)."""
    sd = {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}
    num_layers = int(log2(image_size) - 1)
    b = styles.shape[0]
    x = sd["initial_block"].expand(b, -1, -1, -1)
    per_block = torch.cat((styles.transpose(0, 1), hists.transpose(0, 1)), dim=0)
    rgb, acts = None, []
    for i in range(num_layers):
        x, rgb = generator_block(sd, f"blocks.{i}", x, rgb, per_block[i], noise,
                                 upsample=i != 0, upsample_rgb=i != num_layers - 1)
        acts.append(x)
    return (rgb, acts) if return_all else rgb


def discriminator(sd, x, image_size, prefix=""):
    sd = {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}
    n_blocks = int(log2(image_size) - 1) + 1
    for i in range(n_blocks):
        p = f"blocks.{i}"
        res = F.conv2d(x, sd[p + ".conv_res.weight"], sd[p + ".conv_res.bias"])
        x = F.leaky_relu(F.conv2d(x, sd[p + ".net.0.weight"], sd[p + ".net.0.bias"], padding=1), 0.2)
        x = F.leaky_relu(F.conv2d(x, sd[p + ".net.2.weight"], sd[p + ".net.2.bias"], padding=1), 0.2)
        x = x + res
        if i != n_blocks - 1:
            x = F.conv2d(x, sd[p + ".downsample.weight"], sd[p + ".downsample.bias"], stride=2,
                         padding=1)
    logits = F.linear(x.reshape(x.shape[0], -1), sd["to_logit.weight"], sd["to_logit.bias"])
    return logits.squeeze()


def mlp(sd, prefix, n_layers, x, first=0, step=2):
    """StyleVectorizer ('net') / HistVectorizer ('fcs'): Linear + LeakyReLU(0.2) stacks."""
    for i in range(n_layers):
        x = F.leaky_relu(_linear(sd, f"{prefix}.{first + step * i}", x), 0.2)
    return x


# ------------------------------------------------ deterministic test weights --

def seeded_state_dict(shapes: dict, seed: int = 0) -> dict:
    """Reproducible float32 parameters for a {name: shape} table: kaiming-like scale
    for matrices/filters, small values for biases and noise projections (so that every
    code path -- noise, bias, demodulation -- is numerically visible)."""
    out = {}
    for i, name in enumerate(sorted(shapes)):
        shape = tuple(shapes[name])
        g = torch.Generator().manual_seed(seed * 100003 + i)
        t = torch.randn(shape, generator=g)
        if name.endswith("initial_block"):
            pass
        elif len(shape) >= 2:
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            t = t * (2.0 / fan_in) ** 0.5
            if "to_noise" in name:
                t = t * 0.3
        else:
            t = t * 0.1
        out[name] = t
    return out
