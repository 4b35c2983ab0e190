"""Generator / discriminator stacks of HistoGAN on the tcgen05 convolution.

Same class names, constructor arguments, ``forward`` / ``forward_`` signatures,
output shapes and ``state_dict`` keys as ``histoGAN/histoGAN.py`` (reference
lines cited per class), so checkpoints and the calling scripts carry over.
What differs is HOW the modulated convolution is evaluated: the reference
materialises per-sample weights (B, Cout, Cin, k, k) and runs a grouped conv
(:423-437, 4.8 GB for one layer at batch 32); here the style scales the
activations, one shared-weight implicit GEMM runs on the tensor cores, and the
demodulation is a per-(sample, out-channel) factor on its output.

Activations are channels_last (NHWC in memory) float32; logical shapes stay
(B, C, H, W) as in the reference.
"""
from __future__ import annotations

from math import log2

import torch
import torch.nn.functional as F
from torch import nn

from . import fused, ops

# fused generator layers (fused.py); False = compose torch ops around ops.conv2d
USE_FUSED = True
# precomputed modulations / demodulation factors for the whole generator pass (fused.style_mods,
# demod_all); False = per-block nn.Linear + torch demodulation (debug switch)
import os as _os
STYLE_PATH = _os.environ.get("HG_STYLE_PATH", "1") != "0"
DEMOD_PATH = _os.environ.get("HG_DEMOD_PATH", "1") != "0"      # grouped demodulation + analytic backward

EPS = 1e-8          # histoGAN/histoGAN.py:53


def leaky_relu(p=0.2):
    return nn.LeakyReLU(p, inplace=True)


class Flatten(nn.Module):
    def forward(self, x):
        return x.reshape(x.shape[0], -1)


def _upsample2x():
    return nn.Upsample(scale_factor=2, mode='bilinear', align_corners=False)


def conv_weights_channels_last_(module):
    """store every nn.Conv2d weight of `module` channels_last (see Conv2DMod); shapes, values
    and state_dict keys are untouched, load_state_dict copies into this layout."""
    for m in module.modules():
        if isinstance(m, nn.Conv2d):
            m.weight.data = m.weight.data.contiguous(memory_format=torch.channels_last)
    return module


# --------------------------------------------------------------- operators ---

class Conv2DMod(nn.Module):
    """Modulated / demodulated convolution (histoGAN/histoGAN.py:404-440).

    forward(x, y): x (B, Cin, H, W), y (B, Cin) style -> (B, Cout, H, W).
    ``weight`` keeps the reference layout (Cout, Cin, k, k) and init (:413-415)."""

    def __init__(self, in_chan, out_chan, kernel, demod=True, stride=1, dilation=1, **kwargs):
        super().__init__()
        self.filters = out_chan
        self.demod = demod
        self.kernel = kernel
        self.stride = stride
        self.dilation = dilation
        w = torch.randn((out_chan, in_chan, kernel, kernel))
        nn.init.kaiming_normal_(w, a=0, mode='fan_in', nonlinearity='leaky_relu')
        # stored channels_last ([Cout][k][k][Cin] in memory; shape and state_dict unchanged): the
        # layout of the packed tensor-core operand and of the weight-gradient kernel's output
        self.weight = nn.Parameter(w.contiguous(memory_format=torch.channels_last))

    def _get_same_padding(self, size, kernel, dilation, stride):
        return ((size - 1) * (stride - 1) + dilation * (kernel - 1)) // 2

    def forward(self, x, y):
        if self.stride != 1 or self.dilation != 1:
            raise NotImplementedError("Conv2DMod on the sm_100a path supports stride=dilation=1 "
                                      "(all the reference ever instantiates)")
        h = x.shape[2]
        if (USE_FUSED and fused.fusable(x, self.weight) and x.shape[1] == self.weight.shape[1]
                and self.kernel in (1, 3)):
            return fused.mod_conv_layer(x, y, self.weight, self.demod, None, None, act=False)
        mod = y + 1                                                   # :423-425
        z = ops.conv2d(x * mod[:, :, None, None], self.weight, None, 1,
                       self._get_same_padding(h, self.kernel, self.dilation, self.stride))
        if self.demod:                                                # :427-429
            wsq = self.weight.pow(2).sum(dim=(2, 3))                  # (Cout, Cin)
            d = torch.rsqrt(mod.pow(2) @ wsq.t() + EPS)               # (B, Cout)
            z = z * d[:, :, None, None]
        return z


class RGBBlock(nn.Module):
    """toRGB: 1x1 modulated conv without demodulation + skip + 2x bilinear
    upsampling (histoGAN/histoGAN.py:368-401)."""

    def __init__(self, latent_dim, input_channel, upsample, rgba=False):
        super().__init__()
        self.input_channel = input_channel
        self.to_style = nn.Linear(latent_dim, input_channel)
        out_filters = 3 if not rgba else 4
        self.conv = Conv2DMod(input_channel, out_filters, 1, demod=False)
        self.upsample = _upsample2x() if upsample else None

    def forward(self, x, prev_rgb, istyle):
        return self.forward_(x, prev_rgb, self.to_style(istyle))

    def forward_(self, x, prev_rgb, style):
        if USE_FUSED and self.conv.filters == 3 and x.is_cuda and x.shape[1] % 4 == 0:
            x = fused.to_rgb(x, style, self.conv.weight, prev_rgb)
        else:
            x = self.conv(x, style)
            if prev_rgb is not None:
                x = x + prev_rgb
        if self.upsample is not None:
            if USE_FUSED and x.is_cuda and x.dtype == torch.float32:
                x = fused.upsample2x_planar(x)
            else:
                x = self.upsample(x)
        return x


class GeneratorBlock(nn.Module):
    """upsample -> [mod-conv + noise + LeakyReLU] x 2 -> toRGB
    (histoGAN/histoGAN.py:443-502)."""

    def __init__(self, latent_dim, input_channels, filters, upsample=True, upsample_rgb=True,
                 rgba=False):
        super().__init__()
        self.upsample = _upsample2x() if upsample else None
        self.to_style1 = nn.Linear(latent_dim, input_channels)
        self.to_noise1 = nn.Linear(1, filters)
        self.conv1 = Conv2DMod(input_channels, filters, 3)
        self.to_style2 = nn.Linear(latent_dim, filters)
        self.to_noise2 = nn.Linear(1, filters)
        self.conv2 = Conv2DMod(filters, filters, 3)
        self.activation = leaky_relu()
        self.to_rgb = RGBBlock(latent_dim, filters, upsample_rgb, rgba)

    def _noise_maps(self, inoise, x):
        inoise = inoise[:, :x.shape[2], :x.shape[3], :]
        # Linear(1 -> C) on the last axis, then (B,H,W,C) -> (B,C,W,H): the noise image is
        # used spatially TRANSPOSED, exactly as the reference does (:465-467)
        return (self.to_noise1(inoise).permute((0, 3, 2, 1)),
                self.to_noise2(inoise).permute((0, 3, 2, 1)))

    def forward(self, x, prev_rgb, istyle, inoise, latent=None, _mods=None):
        if _mods is not None:       # Generator.forward's fused path (precomputed modulations)
            return self.forward_mods(x, prev_rgb, *_mods, inoise)
        return self.forward_(x, prev_rgb, None, None, None, inoise=inoise, latent=latent,
                             _istyle=istyle)

    def forward_(self, x, prev_rgb, style1, style2, to_rgb_style, inoise=None, noise1=None,
                 noise2=None, latent=None, _istyle=None):
        if _istyle is not None:
            style1, style2 = self.to_style1(_istyle), self.to_style2(_istyle)
        up = self.upsample is not None
        use_fused = (USE_FUSED and noise1 is None and noise2 is None and inoise is not None
                     and fused.fusable(x, self.conv1.weight) and fused.fusable(x, self.conv2.weight)
                     and x.shape[1] == self.conv1.weight.shape[1]
                     and self.conv1.demod and self.conv2.demod
                     and x.shape[2] == x.shape[3]
                     and inoise.shape[1] >= x.shape[2] * (2 if up else 1)
                     and inoise.shape[2] >= x.shape[3] * (2 if up else 1))
        if use_fused:
            nz = inoise if inoise.is_contiguous() else inoise.contiguous()
            # the block's nn.Upsample is folded into conv1's producer kernel
            x = fused.mod_conv_layer(x, style1, self.conv1.weight, True, nz, self.to_noise1, upsample=up)
            if latent is not None:
                x = x + latent
            x = fused.mod_conv_layer(x, style2, self.conv2.weight, True, nz, self.to_noise2)
        else:
            if up:
                x = self.upsample(x)
            if noise1 is None or noise2 is None:
                if inoise is None:
                    raise Exception('No noise is given')
                noise1, noise2 = self._noise_maps(inoise, x)
            x = self.conv1(x, style1)
            x = self.activation(x + noise1)
            if latent is not None:
                x = x + latent
            x = self.conv2(x, style2)
            x = self.activation(x + noise2)
        if _istyle is not None:
            rgb = self.to_rgb(x, prev_rgb, _istyle)
        else:
            rgb = self.to_rgb.forward_(x, prev_rgb, to_rgb_style)
        return x, rgb

    def style_path_ok(self, x_shape, inoise):
        """can this block run on precomputed modulations (Generator.forward's fused path)?"""
        up = 2 if self.upsample is not None else 1
        w1, w2 = self.conv1.weight, self.conv2.weight
        return (w1.is_cuda and all(c % 4 == 0 for c in (w1.shape[0], w1.shape[1], w2.shape[0]))
                and w1.is_contiguous(memory_format=torch.channels_last)
                and w2.is_contiguous(memory_format=torch.channels_last)
                and self.to_rgb.conv.filters == 3 and x_shape[1] == w1.shape[1]
                and x_shape[2] == x_shape[3] and inoise.shape[1] >= x_shape[2] * up
                and inoise.shape[2] >= x_shape[3] * up)

    def forward_mods(self, x, prev_rgb, mod1, d1, mod2, d2, mod_rgb, inoise):
        """forward on precomputed modulations (style + 1) and demodulation factors"""
        nz = inoise if inoise.is_contiguous() else inoise.contiguous()
        wsq1, wsq2 = ops._packs.get(self.conv1.weight, 'wsq'), ops._packs.get(self.conv2.weight, 'wsq')
        if d1 is None:      # debug switch: demodulation by torch ops + autograd (style = mod - 1)
            x = fused.mod_conv_layer(x, mod1 - 1, self.conv1.weight, True, nz, self.to_noise1,
                                     upsample=self.upsample is not None)
            x = fused.mod_conv_layer(x, mod2 - 1, self.conv2.weight, True, nz, self.to_noise2)
        else:
            x = fused.mod_conv_layer_pre(x, mod1, self.conv1.weight, d1, wsq1, nz, self.to_noise1,
                                         upsample=self.upsample is not None)
            x = fused.mod_conv_layer_pre(x, mod2, self.conv2.weight, d2, wsq2, nz, self.to_noise2)
        rgb = fused.to_rgb_mod(x, mod_rgb, self.to_rgb.conv.weight, prev_rgb)
        if self.to_rgb.upsample is not None:
            rgb = fused.upsample2x_planar(rgb)
        return x, rgb


class DiscriminatorBlock(nn.Module):
    """1x1 residual + 2 x [3x3 conv + LeakyReLU], sum, stride-2 3x3
    (histoGAN/histoGAN.py:505-526).  The nn.Conv2d children only hold the parameters
    (same state_dict keys); the arithmetic runs on the tcgen05 kernels."""

    def __init__(self, input_channels, filters, downsample=True):
        super().__init__()
        self.conv_res = nn.Conv2d(input_channels, filters, 1)
        self.net = nn.Sequential(
            nn.Conv2d(input_channels, filters, 3, padding=1), leaky_relu(),
            nn.Conv2d(filters, filters, 3, padding=1), leaky_relu())
        self.downsample = nn.Conv2d(filters, filters, 3, padding=1, stride=2) if downsample else None
        conv_weights_channels_last_(self)

    @staticmethod
    def _conv(m: nn.Conv2d, x):
        return ops.conv2d(x, m.weight, m.bias, m.stride[0], m.padding[0])

    def forward(self, x):
        if USE_FUSED and x.is_cuda:
            small = ops.SMALL_CIN and ops._small(x.shape[1], self.net[0].weight, 1, 1) and \
                ops._small(x.shape[1], self.conv_res.weight, 1, 0)
            y = self.forward_padded(x if small else ops.round_pad(x), round_out=False)
            c = self.conv_res.out_channels
            return y if y.shape[1] == c else y[:, :c]
        res = self._conv(self.conv_res, x)
        x = F.leaky_relu(self._conv(self.net[0], x), 0.2)
        x = F.leaky_relu(self._conv(self.net[2], x), 0.2)
        x = x + res
        if self.downsample is not None:
            x = self._conv(self.downsample, x)
        return x

    def forward_padded(self, x, round_out=True):
        """fused path on TF32-rounded NHWC tensors whose channel count is padded to a multiple
        of 32 with zeros (only D's 3- and 16-channel ends are affected): 4 kernels per block, bias /
        LeakyReLU / residual sum live in the conv epilogues."""
        c1, c2, cr, dn = self.net[0], self.net[2], self.conv_res, self.downsample
        t = ops.conv_bias_act(x, c1.weight, c1.bias, None, 1, 1, act=True, x_rounded=True,
                              round_out=True)
        t = ops.conv_bias_act(t, c2.weight, c2.bias, None, 1, 1, act=True, x_rounded=True,
                              round_out=False)
        y = ops.conv_bias_act(x, cr.weight, cr.bias, t, 1, 0, act=False, x_rounded=True,
                              round_out=round_out or dn is not None)
        if dn is not None:
            y = ops.conv_bias_act(y, dn.weight, dn.bias, None, 2, 1, act=False, x_rounded=True,
                                  round_out=round_out)
        return y


# ---------------------------------------------------------------- networks ---

class HistVectorizer(nn.Module):
    """histogram -> latent MLP (histoGAN/histoGAN.py:335-351)."""

    def __init__(self, insize, emb, depth):
        super().__init__()
        self.flatten = Flatten()
        dims = [insize * insize * 3, emb * 2] + [emb] * (depth - 1)
        layers = []
        for i in range(depth):
            layers += [nn.Linear(dims[i], dims[i + 1]), leaky_relu()]
        self.fcs = nn.Sequential(*layers)

    def forward(self, x):
        x = self.flatten(x)
        linears = [m for m in self.fcs if isinstance(m, nn.Linear)]
        if USE_FUSED and fused.mlp_ok(x, linears):
            return fused.mlp_lrelu(x, linears, 0.2)
        return self.fcs(x)


class StyleVectorizer(nn.Module):
    """z -> w mapping network (histoGAN/histoGAN.py:354-365)."""

    def __init__(self, emb, depth):
        super().__init__()
        layers = []
        for _ in range(depth):
            layers += [nn.Linear(emb, emb), leaky_relu()]
        self.net = nn.Sequential(*layers)

    def forward(self, x):
        linears = [m for m in self.net if isinstance(m, nn.Linear)]
        if USE_FUSED and fused.mlp_ok(x, linears):
            return fused.mlp_lrelu(x, linears, 0.2)
        return self.net(x)


class Generator(nn.Module):
    """histoGAN/histoGAN.py:529-568."""

    def __init__(self, image_size, latent_dim, network_capacity=16, transparent=False):
        super().__init__()
        self.image_size = image_size
        self.latent_dim = latent_dim
        self.num_layers = int(log2(image_size) - 1)
        init_channels = 4 * network_capacity
        self.initial_block = nn.Parameter(torch.randn((init_channels, 4, 4)))
        filters = [init_channels] + [network_capacity * (2 ** (i + 1))
                                     for i in range(self.num_layers)][::-1]
        self.blocks = nn.ModuleList([
            GeneratorBlock(latent_dim, cin, cout, upsample=i != 0,
                           upsample_rgb=i != self.num_layers - 1, rgba=transparent)
            for i, (cin, cout) in enumerate(zip(filters[:-1], filters[1:]))])

    def forward(self, styles, hists, input_noise):
        batch_size = styles.shape[0]
        x = self.initial_block.expand(batch_size, -1, -1, -1)
        # blocks 0..L-3 take the mapped latent, the last two the histogram latent (:561-563)
        per_block = torch.cat((styles.transpose(0, 1), hists.transpose(0, 1)), dim=0)
        rgb = None
        if USE_FUSED and STYLE_PATH and x.is_cuda and self._style_path_ok(x.shape, input_noise):
            # the 21 to_style Linears (+1) of all blocks in one launch, the 14 demodulation factors
            # in a second one (fused.style_mods / demod_all); the blocks then run on those
            linears = []
            for i, b in enumerate(self.blocks):
                linears += [(i, b.to_style1), (i, b.to_style2), (i, b.to_rgb.to_style)]
            mods = fused.style_mods(per_block, linears)
            conv_mods = [m for i in range(len(self.blocks)) for m in mods[3 * i:3 * i + 2]]
            wsqs = [ops._packs.get(c.weight, 'wsq') for b in self.blocks for c in (b.conv1, b.conv2)]
            ds = fused.demod_all(conv_mods, wsqs, EPS) if DEMOD_PATH else [None] * len(conv_mods)
            for i, block in enumerate(self.blocks):
                # through __call__, so that forward hooks on the blocks keep firing
                x, rgb = block(x, rgb, per_block[i], input_noise,
                               _mods=(mods[3 * i], ds[2 * i], mods[3 * i + 1], ds[2 * i + 1], mods[3 * i + 2]))
            return rgb
        for style, block in zip(per_block, self.blocks):
            x, rgb = block(x, rgb, style, input_noise)
        return rgb

    def _style_path_ok(self, x_shape, inoise):
        shape = list(x_shape)
        for b in self.blocks:
            if not b.style_path_ok(shape, inoise):
                return False
            up = 2 if b.upsample is not None else 1
            shape = [shape[0], b.conv2.weight.shape[0], shape[2] * up, shape[3] * up]
        return True


class Discriminator(nn.Module):
    """histoGAN/histoGAN.py:572-631 with the default fq_layers = attn_layers = [] (the
    optional VectorQuantize / linear-attention blocks depend on packages that are not part
    of the reference tree and are out of scope)."""

    def __init__(self, image_size, network_capacity=16, fq_layers=[], fq_dict_size=256,
                 attn_layers=[], transparent=False):
        super().__init__()
        if len(fq_layers) or len(attn_layers):
            raise NotImplementedError("fq_layers / attn_layers need vector_quantize_pytorch / "
                                      "linear_attention_transformer (not vendored by the reference)")
        num_layers = int(log2(image_size) - 1)
        filters = [3 if not transparent else 4] + [network_capacity * (2 ** i)
                                                   for i in range(num_layers + 1)]
        pairs = list(zip(filters[:-1], filters[1:]))
        self.blocks = nn.ModuleList([DiscriminatorBlock(cin, cout, downsample=i != len(pairs) - 1)
                                     for i, (cin, cout) in enumerate(pairs)])
        self.attn_blocks = nn.ModuleList([None] * len(pairs))
        self.quantize_blocks = nn.ModuleList([None] * len(pairs))
        self.flatten = Flatten()
        self.to_logit = nn.Linear(2 * 2 * filters[-1], 1)

    def forward(self, x):
        quantize_loss = torch.zeros(1, device=x.device, dtype=x.dtype)
        if USE_FUSED and x.is_cuda:
            b0 = self.blocks[0]
            if not (ops.SMALL_CIN and ops._small(x.shape[1], b0.net[0].weight, 1, 1)
                    and ops._small(x.shape[1], b0.conv_res.weight, 1, 0)):
                x = ops.round_pad(x)                        # image: pad 3 -> 32 channels, round once
            # else: block 0's two image-input convs read the planar image directly (conv_small.cu)
            for i, block in enumerate(self.blocks):
                x = block.forward_padded(x, round_out=i != len(self.blocks) - 1)
            c = self.blocks[-1].conv_res.out_channels
            if x.shape[1] != c:
                x = x[:, :c]
        else:
            for block in self.blocks:
                x = block(x)
        x = self.to_logit(self.flatten(x))
        return x.squeeze(), quantize_loss
