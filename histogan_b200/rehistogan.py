"""ReHistoGAN recolouring networks and train step on the sm_100a kernels.

Same class names, constructor arguments, forward signatures and ``state_dict`` keys as
``ReHistoGAN/rehistoGAN.py`` (reference lines cited per class).  The convolutions run on
the tcgen05 implicit GEMM (``ops.conv_bias_act``: bias / LeakyReLU / residual sum live in the
conv epilogue), InstanceNorm + LeakyReLU, the laplacian reconstruction loss and the
gaussian filter of the variance loss are hand-written kernels (csrc/recolor.cu), the
recolouring head reuses the fused generator layers of ``gan.py``.

Activations travel between the blocks as TF32-rounded NHWC tensors whose channel count is
padded to a multiple of 32 (``forward_padded``); the public ``forward`` methods take and
return the reference's logical (B, C, H, W) shapes.
"""
from __future__ import annotations

import json
from math import floor, log2, pi
from pathlib import Path
from shutil import rmtree

import torch
import torch.distributed as dist
import torch.nn.functional as F
from torch import nn
from torch.autograd.function import once_differentiable

from . import _lib, gan, ops
from .gan import (Conv2DMod, Discriminator, GeneratorBlock, HistVectorizer, conv_weights_channels_last_,
                  leaky_relu)
from .hist import RGBuvHistBlock, hellinger_loss
from .optim import DiffGrad
from .trainer import (NanException, Trainer, _allreduce_mean_grads, cast_list, default,
                      gradient_penalty, image_noise, raise_if_nan, set_requires_grad)

EPS = 1e-8
SCALE = 1 / 2 ** 0.5          # rehistoGAN.py:58
IN_EPS = 1e-5                 # nn.InstanceNorm2d default


def _st(dev):
    return _lib.current_stream_ptr(dev)


def _nhwc(x):
    return x if x.is_contiguous(memory_format=torch.channels_last) else \
        x.contiguous(memory_format=torch.channels_last)


def _fused(x):
    return gan.USE_FUSED and x.is_cuda


# ------------------------------------------------------------------ operators ---

class _InstNormLReLU(torch.autograd.Function):
    """LeakyReLU(InstanceNorm2d(x)) on an NHWC tensor: hg_instnorm_lrelu_fwd / _bwd."""

    @staticmethod
    def forward(ctx, x, slope, round_out):
        lib = _lib.load()
        x = _nhwc(x.float())
        B, Cc, H, W = x.shape
        y = torch.empty_like(x, memory_format=torch.channels_last)
        stats = torch.empty((B, Cc, 2), dtype=torch.float64, device=x.device)
        with torch.cuda.device(x.device):
            rc = lib.hg_instnorm_lrelu_fwd(_lib.ptr(x), _lib.ptr(y), _lib.ptr(stats), B, H * W, Cc,
                                           IN_EPS, float(slope), int(round_out), _st(x.device))
        _lib.check(rc, "hg_instnorm_lrelu_fwd")
        ctx.save_for_backward(x, stats)
        ctx.slope = slope
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        x, stats = ctx.saved_tensors
        lib = _lib.load()
        B, Cc, H, W = x.shape
        dy = _nhwc(dy)
        dx = torch.empty_like(x, memory_format=torch.channels_last)
        ws = torch.empty_like(stats)
        with torch.cuda.device(x.device):
            rc = lib.hg_instnorm_lrelu_bwd(_lib.ptr(dy), _lib.ptr(x), _lib.ptr(stats), _lib.ptr(dx),
                                           _lib.ptr(ws), B, H * W, Cc, IN_EPS, float(ctx.slope), 1,
                                           _st(x.device))
        _lib.check(rc, "hg_instnorm_lrelu_bwd")
        return dx, None, None


def instnorm_lrelu(x, slope=0.2, round_out=False):
    _lib.require_cuda(x, "instnorm_lrelu")
    return _InstNormLReLU.apply(x, slope, round_out)


class _Upsample2x(torch.autograd.Function):
    """nn.Upsample(scale_factor=2, mode='bilinear', align_corners=False) on an NHWC tensor,
    result TF32-rounded (its consumers are convolutions): the generator's
    hg_upsample_modulate_round / _bwd kernels with a unit modulation."""

    @staticmethod
    def forward(ctx, x):
        lib = _lib.load()
        x = _nhwc(x.float())
        B, Cc, H, W = x.shape
        ones = torch.ones((B, Cc), dtype=torch.float32, device=x.device)
        y = torch.empty((B, Cc, 2 * H, 2 * W), dtype=torch.float32, device=x.device,
                        memory_format=torch.channels_last)
        with torch.cuda.device(x.device):
            rc = lib.hg_upsample_modulate_round(_lib.ptr(x), _lib.ptr(ones), _lib.ptr(y), B, H, W, Cc,
                                                _st(x.device))
        _lib.check(rc, "hg_upsample_modulate_round")
        ctx.save_for_backward(x, ones)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        x, ones = ctx.saved_tensors
        lib = _lib.load()
        B, Cc, H, W = x.shape
        dy = _nhwc(dy)
        dx = torch.empty_like(x, memory_format=torch.channels_last)
        gmod = torch.empty_like(ones)
        with torch.cuda.device(x.device):
            rc = lib.hg_upsample_modulate_bwd(_lib.ptr(dy), _lib.ptr(x), _lib.ptr(ones), _lib.ptr(dx),
                                              _lib.ptr(gmod), B, H, W, Cc, _st(x.device))
        _lib.check(rc, "hg_upsample_modulate_bwd")
        return dx


class _LaplacianL1(torch.autograd.Function):
    """mean |lap(a) - lap(b)| of reconstruction_loss('2nd gradient'); gradient wrt b only
    (a is the input image batch)."""

    @staticmethod
    def forward(ctx, a, b):
        lib = _lib.load()
        a, b = a.float().contiguous(), b.float().contiguous()
        B, Cc, H, W = b.shape
        loss = torch.empty((), dtype=torch.float32, device=b.device)
        sign = torch.empty((B, H, W), dtype=torch.int8, device=b.device)
        nws = lib.hg_laplacian_l1_workspace_bytes()
        ws = torch.empty((nws,), dtype=torch.uint8, device=b.device)
        with torch.cuda.device(b.device):
            rc = lib.hg_laplacian_l1_fwd(_lib.ptr(a), _lib.ptr(b), _lib.ptr(loss), _lib.ptr(sign),
                                         _lib.ptr(ws), nws, B, H, W, _st(b.device))
        _lib.check(rc, "hg_laplacian_l1_fwd")
        ctx.save_for_backward(sign)
        return loss

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        (sign,) = ctx.saved_tensors
        lib = _lib.load()
        B, H, W = sign.shape
        db = torch.empty((B, 3, H, W), dtype=torch.float32, device=sign.device)
        g = g.float().contiguous()
        with torch.cuda.device(sign.device):
            rc = lib.hg_laplacian_l1_bwd(_lib.ptr(sign), _lib.ptr(g), _lib.ptr(db), B, H, W,
                                         _st(sign.device))
        _lib.check(rc, "hg_laplacian_l1_bwd")
        return None, db


class _DepthwiseConv(torch.autograd.Function):
    """the same K x K filter on every plane, no padding (gaussian_op); the backward is the same
    kernel with pad = K - 1 and the flipped filter."""

    @staticmethod
    def forward(ctx, x, kern):
        x = x.float().contiguous()
        kern = kern.float().contiguous()
        ctx.save_for_backward(kern)
        ctx.hw = tuple(x.shape[2:])
        return _DepthwiseConv._run(x, kern, 0, 0)

    @staticmethod
    def _run(x, kern, pad, flip):
        lib = _lib.load()
        B, Cc, H, W = x.shape
        K = kern.shape[-1]
        y = torch.empty((B, Cc, H + 2 * pad - K + 1, W + 2 * pad - K + 1), dtype=torch.float32,
                        device=x.device)
        with torch.cuda.device(x.device):
            rc = lib.hg_depthwise_conv(_lib.ptr(x), _lib.ptr(kern), _lib.ptr(y), B * Cc, H, W, K, pad,
                                       flip, _st(x.device))
        _lib.check(rc, "hg_depthwise_conv")
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        (kern,) = ctx.saved_tensors
        return _DepthwiseConv._run(dy.float().contiguous(), kern, kern.shape[-1] - 1, 1), None


# ---------------------------------------- stencils / losses (rehistoGAN.py:207-326) ---

def get_gaussian_kernel(kernel_size=15, sigma=3, channels=3):
    """depth-wise gaussian filter module (rehistoGAN.py:207-225); weights as the reference"""
    x_coord = torch.arange(kernel_size)
    x_grid = x_coord.repeat(kernel_size).view(kernel_size, kernel_size)
    y_grid = x_grid.t()
    xy_grid = torch.stack([x_grid, y_grid], dim=-1).float()
    mean = (kernel_size - 1) / 2.
    variance = sigma ** 2.
    gaussian_kernel = (1. / (2. * pi * variance)) * torch.exp(
        -torch.sum((xy_grid - mean) ** 2., dim=-1) / (2 * variance))
    gaussian_kernel = gaussian_kernel / torch.sum(gaussian_kernel)
    gaussian_kernel = gaussian_kernel.view(1, 1, kernel_size, kernel_size).repeat(channels, 1, 1, 1)
    filt = nn.Conv2d(in_channels=channels, out_channels=channels, kernel_size=kernel_size,
                     groups=channels, bias=False)
    filt.weight.data = gaussian_kernel
    filt.weight.requires_grad = False
    return filt


def gaussian_op(x, kernel=None):
    """rehistoGAN.py:228-232: valid depth-wise convolution with the gaussian filter module"""
    if kernel is None:
        kernel = get_gaussian_kernel(kernel_size=15, sigma=15, channels=3).to(x.device)
    w = kernel.weight
    uniform = getattr(kernel, '_hg_uniform', None)       # same filter on every plane? (checked once)
    if uniform is None:
        uniform = bool((w == w[0:1]).all()) and w.shape[-1] <= 15 and w.shape[1] == 1
        kernel._hg_uniform = uniform
    if x.is_cuda and uniform:
        return _DepthwiseConv.apply(x, w[0, 0])
    return kernel(x)


def _stencil(x, k):
    w = torch.tensor(k, dtype=torch.float32, device=x.device).unsqueeze(0).expand(1, x.shape[1], 3, 3)
    return F.conv2d(x, w, stride=1, padding=1)


_LAPLACIAN = [[0, 1, 0], [1, -4, 1], [0, 1, 0]]
_SOBEL = ([[1, 0, -1], [2, 0, -2], [1, 0, -1]], [[1, 2, 1], [0, 0, 0], [-1, -2, -1]])


def laplacian_op(x, kernel=None):
    return F.conv2d(x, kernel, stride=1, padding=1) if kernel is not None else _stencil(x, _LAPLACIAN)


def sobel_op(x, dir=0, kernel=None):
    return F.conv2d(x, kernel, stride=1, padding=1) if kernel is not None else _stencil(x, _SOBEL[dir])


class reconstruction_loss(object):
    """rehistoGAN.py:279-326.  '2nd gradient' (the default 'laplacian' of the trainer) runs as one
    fused kernel on CUDA; 'L1' and '1st gradient' are optional variants composed from torch ops."""

    def __init__(self, loss):
        self.loss = loss

    def compute_loss(self, input, target):
        if self.loss == 'L1':
            return torch.mean(torch.abs(input - target))
        if self.loss == '1st gradient':
            ig = torch.sqrt(sobel_op(input, 0) ** 2 + sobel_op(input, 1) ** 2)
            tg = torch.sqrt(sobel_op(target, 0) ** 2 + sobel_op(target, 1) ** 2)
            return torch.mean(torch.abs(ig - tg))
        if self.loss == '2nd gradient':
            if target.is_cuda and input.shape[1] == 3 and not input.requires_grad:
                return _LaplacianL1.apply(input, target)
            return torch.mean(torch.abs(laplacian_op(input) - laplacian_op(target)))
        return None


# ------------------------------------------------------------------- networks ---

def _cba(x, m: nn.Conv2d, res=None, act=False, x_rounded=True, round_out=False):
    return ops.conv_bias_act(x, m.weight, m.bias, res, m.stride[0], m.padding[0], act=act,
                             x_rounded=x_rounded, round_out=round_out)


def _conv(m: nn.Conv2d, x):
    return ops.conv2d(x, m.weight, m.bias, m.stride[0], m.padding[0])


def _logical(x, c):
    return x if x.shape[1] == c else x[:, :c]


class EncoderBlock(nn.Module):
    """1x1 residual + 2 x [3x3 conv + InstanceNorm + LeakyReLU], sum, stride-2 3x3
    (rehistoGAN.py:485-504).  forward(x) -> (downsampled, full resolution)."""

    def __init__(self, input_channels, filters):
        super().__init__()
        self.conv_res = nn.Conv2d(input_channels, filters, 1)
        self.net = nn.Sequential(
            nn.Conv2d(input_channels, filters, 3, padding=1), nn.InstanceNorm2d(filters), leaky_relu(),
            nn.Conv2d(filters, filters, 3, padding=1), nn.InstanceNorm2d(filters), leaky_relu())
        self.downsample = nn.Conv2d(filters, filters, 3, padding=1, stride=2)
        conv_weights_channels_last_(self)

    def forward(self, x):
        if _fused(x):
            c = self.conv_res.out_channels
            xd, xu = self.forward_padded(ops.round_pad(x), round_out=False)
            return _logical(xd, c), _logical(xu, c)
        res = _conv(self.conv_res, x)
        y = F.leaky_relu(F.instance_norm(_conv(self.net[0], x), eps=IN_EPS), 0.2)
        y = F.leaky_relu(F.instance_norm(_conv(self.net[3], y), eps=IN_EPS), 0.2)
        y = y + res
        return _conv(self.downsample, y), y

    def forward_padded(self, x, round_out=True):
        """x: TF32-rounded, channel-padded NHWC.  Returns (downsampled [rounded if round_out],
        full-resolution sum [not rounded: it also feeds the modulated skip convolutions])."""
        t = instnorm_lrelu(_cba(x, self.net[0]), round_out=True)
        t = instnorm_lrelu(_cba(t, self.net[3]), round_out=False)
        y = _cba(x, self.conv_res, res=t)
        return _cba(y, self.downsample, x_rounded=False, round_out=round_out), y


class DecoderBlock(nn.Module):
    """rehistoGAN.py:507-546.  forward(x, prev_rgb, prev_latent, h=None) -> (x, rgb), both 2x
    up-sampled."""

    def __init__(self, input_channels, filters, internal_hist=False, latent_dim=None):
        super().__init__()
        self.upsample = nn.Upsample(scale_factor=2, mode='bilinear', align_corners=False)
        self.conv_res = nn.Conv2d(input_channels, filters, 1)
        self.block1 = nn.Sequential(nn.Conv2d(input_channels, input_channels, 3, padding=1), leaky_relu())
        self.block2 = nn.Sequential(nn.Conv2d(input_channels * 2, filters, 3, padding=1), leaky_relu())
        self.conv_out_latent = nn.Sequential(nn.Conv2d(filters, filters, 3, padding=1), leaky_relu())
        self.conv_out_rgb = nn.Conv2d(filters, 3, 1)
        if internal_hist:
            self.to_latent = nn.Linear(latent_dim, input_channels)
            self.conv_latent = Conv2DMod(input_channels, input_channels, 3)
        else:
            self.to_latent = None
            self.conv_latent = None
        conv_weights_channels_last_(self)

    def forward(self, x, prev_rgb, prev_latent, h=None):
        if _fused(x):
            xo, rgb = self.forward_padded(ops.round_pad(x), prev_rgb, ops.round_pad(prev_latent), h)
            return _logical(xo, self.conv_res.out_channels), rgb
        cur = F.leaky_relu(_conv(self.block1[0], x), 0.2)
        if self.to_latent is not None:
            prev_latent = self.conv_latent(prev_latent, self.to_latent(h))
        proc = F.leaky_relu(_conv(self.block2[0], torch.cat((cur, prev_latent), dim=1)), 0.2)
        xo = F.leaky_relu(_conv(self.conv_out_latent[0], _conv(self.conv_res, x) + proc), 0.2)
        rgb = _conv(self.conv_out_rgb, xo)
        if prev_rgb is not None:
            rgb = rgb + prev_rgb
        return self.upsample(xo), self.upsample(rgb)

    def forward_padded(self, x, prev_rgb, prev_latent, h=None):
        cin = self.block1[0].in_channels
        cur = _cba(x, self.block1[0], act=True, round_out=True)
        if self.to_latent is not None:
            prev_latent = ops.round_pad(self.conv_latent(_logical(prev_latent, cin), self.to_latent(h)))
        if cur.shape[1] != cin:             # padded channels would split the concatenation
            both = ops.round_pad(torch.cat((cur[:, :cin], prev_latent[:, :cin]), dim=1))
        else:
            both = _nhwc(torch.cat((cur, prev_latent), dim=1))
        proc = _cba(both, self.block2[0], act=True)
        s = _cba(x, self.conv_res, res=proc, round_out=True)
        xo = _cba(s, self.conv_out_latent[0], act=True, round_out=True)
        rgb = _cba(xo, self.conv_out_rgb)[:, :3]
        if prev_rgb is not None:
            rgb = rgb + prev_rgb
        return _Upsample2x.apply(xo), self.upsample(rgb)


class RecoloringEncoderDecoder(nn.Module):
    """rehistoGAN.py:549-634."""

    def __init__(self, image_size, network_capacity=16, hist=64, latent_dim=512, style_depth=8,
                 skip_conn_to_GAN=False, internal_hist=False):
        super().__init__()
        self.image_size = image_size
        self.encoder_num_layers = int(log2(image_size) - 2)
        self.decoder_num_layers = int(log2(image_size) - 4)
        self.skip_conn_to_GAN = skip_conn_to_GAN
        self.internal_hist = internal_hist
        encoder_filters = [network_capacity] + [network_capacity * (2 ** (i + 1))
                                                for i in range(self.encoder_num_layers)]
        encoder_pairs = list(zip(encoder_filters[:-1], encoder_filters[1:]))
        rev = encoder_filters[::-1]                 # the reference reverses the list in place
        decoder_filters = rev[:-(self.encoder_num_layers - self.decoder_num_layers)]
        decoder_pairs = list(zip(decoder_filters[:-1], decoder_filters[1:]))
        self.encoder_blocks = nn.ModuleList([])
        self.decoder_blocks = nn.ModuleList([])
        self.decoder_mapping = nn.Conv2d(decoder_filters[-1], 8 * network_capacity, 1)
        self.mapping = nn.Conv2d(3, network_capacity, 3, padding=1)
        if self.skip_conn_to_GAN:
            if not self.internal_hist:
                self.hist_projection = HistVectorizer(hist, latent_dim, int(style_depth))
            self.to_latent_1 = nn.Linear(latent_dim, rev[-3])
            self.to_latent_2 = nn.Linear(latent_dim, rev[-2])
            self.conv_latent_1 = Conv2DMod(rev[-3], 2 ** 2 * network_capacity, 3)
            self.conv_latent_2 = Conv2DMod(rev[-2], 2 ** (2 - 1) * network_capacity, 3)
        for cin, cout in encoder_pairs:
            self.encoder_blocks.append(EncoderBlock(cin, cout))
        for cin, cout in decoder_pairs:
            self.decoder_blocks.append(DecoderBlock(cin, cout, internal_hist=self.internal_hist,
                                                    latent_dim=latent_dim))
        conv_weights_channels_last_(self)

    def forward(self, x, hists=None):
        if self.skip_conn_to_GAN and not self.internal_hist:
            h_w_space = self.hist_projection(hists)
            h1, h2 = self.to_latent_1(h_w_space), self.to_latent_2(h_w_space)
        elif self.skip_conn_to_GAN and self.internal_hist:
            h1, h2 = self.to_latent_1(hists), self.to_latent_2(hists)
        fused = _fused(x)
        if fused:
            x = _cba(ops.round_pad(x), self.mapping, round_out=True)
        else:
            x = _conv(self.mapping, x)
        x_list, x_list_up = [], []
        for block in self.encoder_blocks:
            x, xup = block.forward_padded(x) if fused else block(x)
            x_list.append(x)
            x_list_up.append(xup)
        x_list.reverse()
        x_list_e = x_list[:-2]
        if self.skip_conn_to_GAN:
            eb = self.encoder_blocks
            processed_latent_1 = self.conv_latent_1(
                _logical(x_list_up[1], eb[1].conv_res.out_channels), h1)
            processed_latent_2 = self.conv_latent_2(
                _logical(x_list_up[0], eb[0].conv_res.out_channels), h2)
        rgb = None
        for prev_latent, block in zip(x_list_e, self.decoder_blocks):
            x, rgb = block.forward_padded(x, rgb, prev_latent, h=hists) if fused else \
                block(x, rgb, prev_latent, h=hists)
        if fused:
            x = _logical(_cba(x, self.decoder_mapping), self.decoder_mapping.out_channels)
        else:
            x = _conv(self.decoder_mapping, x)
        if self.skip_conn_to_GAN:
            return x, rgb, processed_latent_1, processed_latent_2
        return x, rgb


class RecoloringGAN(nn.Module):
    """the last two GeneratorBlocks of HistoGAN used as the recolouring head
    (rehistoGAN.py:449-482)."""

    def __init__(self, image_size, latent_dim, network_capacity=16, transparent=False):
        super().__init__()
        self.image_size = image_size
        self.latent_dim = latent_dim
        num_layers = int(log2(image_size) - 1)
        init_channels = 4 * network_capacity
        filters = [init_channels] + [network_capacity * (2 ** (i + 1)) for i in range(num_layers)][::-1]
        filters = filters[-3:]
        self.num_layers = 2
        self.blocks = nn.ModuleList([
            GeneratorBlock(latent_dim, cin, cout, upsample=True, upsample_rgb=ind != 1, rgba=transparent)
            for ind, (cin, cout) in enumerate(zip(filters[:-1], filters[1:]))])

    def forward(self, x, rgb, hists, input_noise, latent1=None, latent2=None):
        rgb = None                                   # (sic) the decoder's rgb is dropped, :479
        x, rgb = self.blocks[0](x, rgb, hists, input_noise, latent=latent1)
        x, rgb = self.blocks[1](x, rgb, hists, input_noise, latent=latent2)
        return rgb


class recoloringGAN(nn.Module):
    """ED + H + G + D and the two DiffGrad optimisers (rehistoGAN.py:637-718)."""

    def __init__(self, image_size, latent_dim=512, style_depth=8, network_capacity=16,
                 transparent=False, fp16=False, steps=1, lr=1e-4, fq_layers=[], fq_dict_size=256,
                 attn_layers=[], hist=64, skip_conn_to_GAN=False, fixed_gan_weights=False,
                 initialize_gan=False, internal_hist=False):
        super().__init__()
        assert not fp16, 'Apex mixed precision is not available on the sm_100a path'
        self.lr = lr
        self.steps = steps
        self.fixed_gan_weights = fixed_gan_weights
        self.internal_hist = internal_hist
        self.skip_conn_to_GAN = skip_conn_to_GAN
        self.ED = RecoloringEncoderDecoder(image_size, network_capacity=network_capacity, hist=hist,
                                           latent_dim=latent_dim, style_depth=style_depth,
                                           skip_conn_to_GAN=skip_conn_to_GAN,
                                           internal_hist=self.internal_hist)
        self.H = HistVectorizer(hist, latent_dim, int(style_depth))
        self.G = RecoloringGAN(image_size, latent_dim, network_capacity, transparent=transparent)
        self.D = Discriminator(image_size, network_capacity, fq_layers=fq_layers,
                               fq_dict_size=fq_dict_size, attn_layers=attn_layers,
                               transparent=transparent)
        for m in (self.ED, self.H, self.G, self.D):
            set_requires_grad(m, True)
        if not self.fixed_gan_weights:
            learnable = list(self.ED.parameters()) + list(self.G.parameters()) + list(self.H.parameters())
        else:
            learnable = list(self.ED.parameters())
        self.G_opt = DiffGrad(learnable, lr=self.lr, betas=(0.5, 0.9))
        self.D_opt = DiffGrad(self.D.parameters(), lr=self.lr, betas=(0.5, 0.9))
        self._init_weights(initializeGAN=bool(initialize_gan))
        self.cuda()

    def _init_weights(self, initializeGAN=False):
        if initializeGAN:
            for block in self.G.blocks:
                for t in (block.to_noise1, block.to_noise2):
                    nn.init.zeros_(t.weight)
                    nn.init.zeros_(t.bias)
            mods = list(self.H.modules())
        else:
            mods = []
        for m in mods + list(self.ED.modules()) + list(self.D.modules()):
            if type(m) in {nn.Conv2d, nn.Linear}:
                nn.init.kaiming_normal_(m.weight, a=0, mode='fan_in', nonlinearity='leaky_relu')

    def forward(self, x):
        return x


# -------------------------------------------------------------------- trainer ---

class recoloringTrainer():
    """train-step drop-in for rehistoGAN.py:721-1075 (constructor signature, attributes,
    ``train(alpha, beta, gamma)``, save / load of the raw state_dict).  Image evaluation /
    post-processing (``evaluate``, pyramid up-sampling) is host-side I/O and out of scope."""

    def __init__(self, name, results_dir, models_dir, image_size, network_capacity,
                 transparent=False, batch_size=4, mixed_prob=0.9, gradient_accumulate_every=1,
                 lr=2e-4, num_workers=None, save_every=1000, trunc_psi=0.6, fp16=False,
                 fq_layers=[], fq_dict_size=256, attn_layers=[], hist_method='inverse-quadratic',
                 hist_resizing='sampling', hist_sigma=0.02, hist_bin=64, hist_insz=150,
                 fixed_gan_weights=False, skip_conn_to_GAN=False, rec_loss='laplacian',
                 initialize_gan=False, variance_loss=True, internal_hist=False,
                 change_hyperparameters=False, change_hyperparameters_after=100000, *args, **kwargs):
        self.fast_rng = bool(kwargs.pop('fast_rng', False))
        # cuda_graphs=True: each phase (forward + backward) of a step is replayed as one CUDA graph
        self.cuda_graphs = bool(kwargs.pop('cuda_graphs', False))
        self._graphs = {}
        self._static = None
        self.graph_replayed_launches = 0
        self.GAN_params = [args, kwargs]
        self.GAN = None
        self.hist_method = hist_method
        self.hist_resizing = hist_resizing
        self.hist_sigma = hist_sigma
        self.hist_bin = hist_bin
        self.change_hyperparameters_after = change_hyperparameters_after
        self.hist_insz = hist_insz
        self.rec_loss = rec_loss
        self.internal_hist = internal_hist
        self.change_hyperparameters = change_hyperparameters
        self.variance_loss = variance_loss
        self.fixed_gan_weights = fixed_gan_weights
        self.skip_conn_to_GAN = skip_conn_to_GAN
        self.initialize_gan = initialize_gan
        self.histBlock = RGBuvHistBlock(insz=self.hist_insz, h=self.hist_bin, method=self.hist_method,
                                        resizing=self.hist_resizing, sigma=self.hist_sigma)
        if variance_loss is True:
            self.histBlock_input = RGBuvHistBlock(insz=self.hist_insz, h=self.hist_bin,
                                                  method=self.hist_method,
                                                  resizing=self.hist_resizing, sigma=self.hist_sigma)
            self.gaussKernel = get_gaussian_kernel(kernel_size=15, sigma=5, channels=3).to(
                device=torch.cuda.current_device())
        if self.rec_loss is None:
            self.rec_loss_func = reconstruction_loss('L1')
        elif self.rec_loss == 'sobel':
            self.rec_loss_func = reconstruction_loss('1st gradient')
        elif self.rec_loss == 'laplacian':
            self.rec_loss_func = reconstruction_loss('2nd gradient')
        else:
            raise Exception('Unknown reconstruction losst!')
        self.name = name
        self.results_dir = Path(results_dir)
        self.models_dir = Path(models_dir)
        self.config_path = self.models_dir / name / '.config.json'
        assert log2(image_size).is_integer(), 'image size must be a power of 2 (64, 128, 256, 512, 1024)'
        self.image_size = image_size
        self.network_capacity = network_capacity
        self.transparent = transparent
        self.fq_layers = cast_list(fq_layers)
        self.fq_dict_size = fq_dict_size
        self.attn_layers = cast_list(attn_layers)
        self.lr = lr
        self.batch_size = batch_size
        self.num_workers = num_workers
        self.mixed_prob = mixed_prob
        self.save_every = save_every
        self.steps = 0
        self.av = None
        self.trunc_psi = trunc_psi
        self.gradient_accumulate_every = gradient_accumulate_every
        assert not fp16, 'Apex mixed precision is not available on the sm_100a path'
        self.fp16 = fp16
        self.d_loss = self.g_loss = self.r_loss = self.h_loss = 0
        self.last_gp_loss = 0
        self.last_cr_loss = 0
        self.q_loss = 0
        if self.variance_loss is True:
            self.var_loss = 0
        self.init_folders()
        self.loader = None
        self.loader_evaluate = None

    def init_GAN(self):
        args, kwargs = self.GAN_params
        self._graphs, self._static = {}, None       # captured graphs belong to the old GAN
        self.GAN = recoloringGAN(lr=self.lr, image_size=self.image_size,
                                 network_capacity=self.network_capacity, transparent=self.transparent,
                                 fq_layers=self.fq_layers, fq_dict_size=self.fq_dict_size,
                                 attn_layers=self.attn_layers, fp16=self.fp16, hist=self.hist_bin,
                                 fixed_gan_weights=self.fixed_gan_weights,
                                 skip_conn_to_GAN=self.skip_conn_to_GAN,
                                 initialize_gan=self.initialize_gan, internal_hist=self.internal_hist,
                                 *args, **kwargs)
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            for p in self.GAN.parameters():
                dist.broadcast(p.data, src=0)

    def config(self):
        return {'image_size': self.image_size, 'network_capacity': self.network_capacity,
                'transparent': self.transparent, 'fq_layers': self.fq_layers,
                'fq_dict_size': self.fq_dict_size, 'attn_layers': self.attn_layers}

    def write_config(self):
        self.config_path.write_text(json.dumps(self.config()))

    def _is_main(self):
        return not (dist.is_available() and dist.is_initialized()) or dist.get_rank() == 0

    # ------------------------------------------------------------------ train --
    def _recolor(self, image_batch, hist_batch, noise):
        """the four forward variants of :935-957 / :987-1006"""
        GAN = self.GAN
        h_w_space = GAN.H(hist_batch)
        ed_hist = h_w_space if self.internal_hist else hist_batch
        if self.skip_conn_to_GAN:
            # (sic) the reference unpacks ED's (.., processed_latent_1, processed_latent_2) into
            # the swapped names and passes them swapped again: latent1 = conv_latent_1's output
            image_latent, rgb, lat1, lat2 = GAN.ED(image_batch, ed_hist)
            return GAN.G(image_latent, rgb, h_w_space, noise, lat1, lat2)
        image_latent, rgb = GAN.ED(image_batch, ed_hist)
        return GAN.G(image_latent, rgb, h_w_space, noise)

    def g_losses(self, image_batch, hist_batch, generated_images, alpha, beta, gamma):
        """the generator objective of :1003-1026 -> (d_loss, histogram_loss, reconstruction_loss,
        var_loss or None)"""
        fake_output, _ = self.GAN.D(generated_images)
        d_loss = gamma * fake_output.mean()
        generated_histograms = self.histBlock(F.relu(generated_images))
        histogram_loss = hellinger_loss(hist_batch, generated_histograms, alpha)
        rec = beta * self.rec_loss_func.compute_loss(image_batch, generated_images)
        var_loss = None
        if self.variance_loss is True:
            with torch.no_grad():
                input_histograms = self.histBlock_input(F.relu(hist_batch))
                input_gauss = gaussian_op(image_batch, kernel=self.gaussKernel)
                hist_term = torch.sum(torch.abs(hist_batch - input_histograms))
                input_std = torch.std(torch.std(input_gauss, dim=2), dim=2)
            generated_gauss = gaussian_op(generated_images, kernel=self.gaussKernel)
            var_loss = -1 * (beta / 10) * hist_term * torch.mean(torch.abs(
                input_std - torch.std(torch.std(generated_gauss, dim=2), dim=2)))
        return d_loss, histogram_loss, rec, var_loss

    # ---------------------------------------------------------- CUDA-graph path --
    _graphed = Trainer._graphed          # capture once / replay / re-attach the graph's gradients
    _capture = Trainer._capture
    _replay = Trainer._replay

    def _phase_d(self, apply_gp):
        GAN, st = self.GAN, self._static
        GAN.D_opt.zero_grad(set_to_none=True)
        noise = torch.rand(self.batch_size, GAN.G.image_size, GAN.G.image_size, 1, device='cuda')
        with torch.no_grad():
            fake = self._recolor(st['images'], st['hists'], noise)
        images = st['images'].detach().requires_grad_(True) if apply_gp else st['images']
        fake_out, _ = GAN.D(fake)
        real_out, _ = GAN.D(images)
        divergence = (F.relu(1 + real_out) + F.relu(1 - fake_out)).mean()
        loss, gp = divergence, None
        if apply_gp:
            gp = gradient_penalty(images, real_out)
            loss = loss + gp
        loss.backward()
        return divergence.detach(), (gp.detach() if gp is not None else None)

    def _phase_g(self, alpha, beta, gamma):
        GAN, st = self.GAN, self._static
        GAN.G_opt.zero_grad(set_to_none=True)
        noise = torch.rand(self.batch_size, GAN.G.image_size, GAN.G.image_size, 1, device='cuda')
        generated = self._recolor(st['images'], st['hists'], noise)
        set_requires_grad(GAN.D, False)
        try:
            d_loss, h_loss, r_loss, v_loss = self.g_losses(st['images'], st['hists'], generated,
                                                           alpha, beta, gamma)
            total = d_loss + h_loss + r_loss
            if v_loss is not None:
                total = total + v_loss
            total.backward()
        finally:
            set_requires_grad(GAN.D, True)
        return (d_loss.detach(), h_loss.detach(), r_loss.detach(),
                v_loss.detach() if v_loss is not None else None)

    def _train_graphed(self, alpha, beta, gamma, apply_gp):
        GAN = self.GAN
        B, S_ = self.batch_size, GAN.G.image_size
        if self._static is None:
            self._static = {'images': torch.zeros(B, 3, S_, S_, device='cuda'),
                            'hists': torch.zeros(B, 3, self.hist_bin, self.hist_bin, device='cuda')}
        st = self._static

        def stage(batch):
            st['images'].copy_(batch['images'], non_blocking=True)
            st['hists'].copy_(batch['histograms'], non_blocking=True)

        d_params = list(GAN.D.parameters())
        g_params = [p for grp in GAN.G_opt.param_groups for p in grp['params']]
        stage(next(self.loader))
        divergence, gp = self._graphed(('D', apply_gp), lambda: self._phase_d(apply_gp), d_params)
        _allreduce_mean_grads(d_params)
        GAN.D_opt.step()
        stage(next(self.loader))
        key = ('G', float(alpha), float(beta), float(gamma))
        d_loss, h_loss, r_loss, v_loss = self._graphed(key, lambda: self._phase_g(alpha, beta, gamma),
                                                       g_params)
        _allreduce_mean_grads([p for p in g_params if p.grad is not None])
        GAN.G_opt.step()
        if gp is not None:
            self.last_gp_loss = gp.item()
        vals = torch.stack((divergence, d_loss, r_loss, h_loss,
                            v_loss if v_loss is not None else torch.zeros_like(d_loss))).tolist()
        self.d_loss, self.g_loss, self.r_loss, self.h_loss = vals[:4]
        if self.variance_loss is True:
            self.var_loss = vals[4]
        self.q_loss = 0.0
        return divergence.clone(), d_loss.clone()

    def _finish_step(self, total_disc_loss, total_gen_loss):
        checkpoint_num = floor(self.steps / self.save_every)
        nan_flag = torch.isnan(total_gen_loss) | torch.isnan(total_disc_loss)
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            f = nan_flag.float()
            dist.all_reduce(f, op=dist.ReduceOp.MAX)
            nan_flag = f > 0
        if bool(nan_flag):
            print(f'NaN detected for generator or discriminator. Loading from checkpoint '
                  f'#{checkpoint_num}')
            self.load(checkpoint_num)
            raise NanException
        if self.steps % self.save_every == 0 and self._is_main():
            self.save(checkpoint_num)
        self.steps += 1
        self.av = None

    def train(self, alpha=32, beta=1.5, gamma=4):
        assert self.loader is not None, ('You must first initialize the data source with '
                                         '`. set_data_src(<folder of images>)`')
        if self.steps >= self.change_hyperparameters_after and self.change_hyperparameters:
            self.alpha, self.gamma, self.beta = 8, 2, 1         # (sic) sets attributes only, :901-905
        torch.autograd.set_detect_anomaly(False)
        if self.GAN is None:
            self.init_GAN()
        GAN = self.GAN
        if not GAN.training:
            GAN.train()
        if self.cuda_graphs and self.gradient_accumulate_every == 1:
            return self._finish_step(*self._train_graphed(alpha, beta, gamma, self.steps % 4 == 0))
        dev = torch.device('cuda', torch.cuda.current_device())
        total_disc_loss = torch.zeros((), device=dev)
        total_gen_loss = torch.zeros((), device=dev)
        total_rec_loss = torch.zeros((), device=dev)
        total_hist_loss = torch.zeros((), device=dev)
        total_var_loss = torch.zeros((), device=dev)
        batch_size, image_size = self.batch_size, GAN.G.image_size
        accum = self.gradient_accumulate_every
        apply_gradient_penalty = self.steps % 4 == 0

        # ---------------------------------------------------- discriminator --
        GAN.D_opt.zero_grad()
        for _ in range(accum):
            batch = next(self.loader)
            # a fresh leaf: the loader may hand out the same resident tensor every step
            image_batch = batch['images'].cuda(non_blocking=True).detach().requires_grad_()
            hist_batch = batch['histograms'].cuda(non_blocking=True)
            noise = image_noise(batch_size, image_size, self.fast_rng)
            with torch.no_grad():             # the graph of the fake is detached right away (:958)
                generated_images = self._recolor(image_batch, hist_batch, noise)
            fake_output, fake_q_loss = GAN.D(generated_images)
            real_output, real_q_loss = GAN.D(image_batch)
            divergence = (F.relu(1 + real_output) + F.relu(1 - fake_output)).mean()
            quantize_loss = (fake_q_loss + real_q_loss).mean()
            disc_loss = divergence + quantize_loss
            if apply_gradient_penalty:
                gp = gradient_penalty(image_batch, real_output)
                self.last_gp_loss = gp.clone().detach().item()
                disc_loss = disc_loss + gp
            disc_loss = disc_loss / accum
            disc_loss.register_hook(raise_if_nan)
            disc_loss.backward()
            total_disc_loss += divergence.detach() / accum
            self.q_loss = float(quantize_loss.detach().item())
        self.d_loss = float(total_disc_loss)
        _allreduce_mean_grads(list(GAN.D.parameters()))
        GAN.D_opt.step()

        # -------------------------------------------------------- generator --
        GAN.G_opt.zero_grad()
        g_params = [p for grp in GAN.G_opt.param_groups for p in grp['params']]
        set_requires_grad(GAN.D, False)       # D's parameter gradients of this phase are never used
        for _ in range(accum):
            batch = next(self.loader)
            image_batch = batch['images'].cuda(non_blocking=True).detach()
            hist_batch = batch['histograms'].cuda(non_blocking=True).detach()
            noise = image_noise(batch_size, image_size, self.fast_rng)
            generated_images = self._recolor(image_batch, hist_batch, noise)
            d_loss, histogram_loss, rec, var_loss = self.g_losses(
                image_batch, hist_batch, generated_images, alpha, beta, gamma)
            gen_loss = d_loss + histogram_loss + rec
            if var_loss is not None:
                gen_loss = gen_loss + var_loss
                total_var_loss += var_loss.detach() / accum
            gen_loss = gen_loss / accum
            gen_loss.register_hook(raise_if_nan)
            gen_loss.backward()
            total_rec_loss += rec.detach() / accum
            total_gen_loss += d_loss.detach() / accum
            total_hist_loss += histogram_loss.detach() / accum
        set_requires_grad(GAN.D, True)
        # one host read for all the logged scalars
        g, r, h, v = torch.stack((total_gen_loss, total_rec_loss, total_hist_loss, total_var_loss)).tolist()
        self.g_loss, self.r_loss, self.h_loss = g, r, h
        if self.variance_loss is True:
            self.var_loss = v
        _allreduce_mean_grads([p for p in g_params if p.grad is not None])
        GAN.G_opt.step()

        self._finish_step(total_disc_loss, total_gen_loss)

    # --------------------------------------------------------------- storage --
    def print_log(self):
        print(f'\nG: {self.g_loss:.2f} | D: {self.d_loss:.2f} | GP: {self.last_gp_loss:.2f} | '
              f'R: {self.r_loss:.2f} | H: {self.h_loss:.2f}' +
              (f' | V: {self.var_loss:.2f}' if self.variance_loss is True else ''))

    def model_name(self, num):
        return str(self.models_dir / self.name / f'model_{num}.pt')

    def init_folders(self):
        (self.results_dir / self.name).mkdir(parents=True, exist_ok=True)
        (self.models_dir / self.name).mkdir(parents=True, exist_ok=True)

    def clear(self):
        rmtree(f'./models/{self.name}', True)
        rmtree(f'./results/{self.name}', True)
        rmtree(str(self.config_path), True)
        self.init_folders()

    def save(self, num):
        torch.save(self.GAN.state_dict(), self.model_name(num))
        self.write_config()

    def load(self, num=-1):
        name = num
        if num == -1:
            file_paths = [p for p in Path(self.models_dir / self.name).glob('model_*.pt')]
            saved_nums = sorted(map(lambda x: int(x.stem.split('_')[1]), file_paths))
            if len(saved_nums) == 0:
                return
            name = saved_nums[-1]
            print(f'continuing from previous epoch - {name}')
        self.steps = name * self.save_every
        if self.GAN is None:
            self.init_GAN()
        self.GAN.load_state_dict(torch.load(self.model_name(name), map_location='cuda'))
