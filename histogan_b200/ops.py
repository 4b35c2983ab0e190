"""Differentiable convolution built on the tcgen05 kernels.

Three autograd Functions -- conv, its input gradient and its weight gradient --
each of which is linear in its two tensor arguments and whose backward is
expressed through the other two, so gradients of any order exist (the gradient
penalty of ``histoGAN/histoGAN.py:156-163`` back-propagates through
``d D(x) / d x``).  All three run on hand-written sm_100a kernels:

    conv        hg_conv2d_fwd                     (conv_tc.cu)
    grad input  hg_conv2d_fwd with flipped/transposed packed weights; a stride-2
                conv's input gradient is the stride-1 conv of the zero-dilated dy
    grad weight hg_conv2d_wgrad                   (conv_wgrad_tc.cu)

Operands are rounded to TF32 (round-to-nearest-even) right before the MMA and
accumulated in fp32, which keeps the generator within ~3e-4 of the fp32
reference (DESIGN.md "precision").  Channel counts that are not multiples of 32
(the RGB ends of G and D, D's 16-channel block) are zero-padded to 32 for the
kernel call and sliced back.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from . import conv as _conv

_CH = 32


def _round_up(n, m=_CH):
    return (n + m - 1) // m * m


def _round_tf32_nhwc(x: torch.Tensor) -> torch.Tensor:
    """fp32 channels_last copy of x, channels zero-padded to a multiple of 32, TF32-rounded."""
    lib = _lib.load()
    B, Cc, H, W = x.shape
    Cp = _round_up(Cc)
    if x.dtype != torch.float32:
        x = x.float()
    if Cp != Cc:
        src = x.new_zeros((B, Cp, H, W)).contiguous(memory_format=torch.channels_last)
        src[:, :Cc] = x
    else:
        src = x if x.is_contiguous(memory_format=torch.channels_last) else \
            x.contiguous(memory_format=torch.channels_last)
    out = torch.empty_like(src, memory_format=torch.channels_last)
    if src.numel():
        with torch.cuda.device(x.device):
            rc = lib.hg_modulate_round(_lib.ptr(src), None, _lib.ptr(out), B, H * W, Cp, 1,
                                       _lib.current_stream_ptr(x.device))
        _lib.check(rc, "hg_modulate_round")
    return out


class _PackCache:
    """Packed (K-major, TF32) copies of a weight, kept ON the weight tensor object and
    reused until the tensor is modified in place (optimizer step / load_state_dict bump
    ``_version``).  Nothing is keyed by address, so a freed-and-reallocated tensor can
    never hit a stale entry."""

    ATTR = "_hg_packed"

    def get(self, w: torch.Tensor, mode: int) -> torch.Tensor:
        ver = w._version
        store = getattr(w, self.ATTR, None)
        if store is not None:
            hit = store.get(mode)
            if hit is not None and hit[0] == ver:
                return hit[1]
        co, ci, kh, kw = w.shape
        cop, cip = _round_up(co), _round_up(ci)
        wd = w.detach().float()
        if (cop, cip) != (co, ci):
            wp = wd.new_zeros((cop, cip, kh, kw))
            wp[:co, :ci] = wd
            wd = wp
        packed = _conv.pack_weight(wd, mode)
        if w.is_leaf:                       # parameters persist; temporaries are not worth caching
            if store is None:
                store = {}
                try:
                    setattr(w, self.ATTR, store)
                except AttributeError:
                    return packed
            store[mode] = (ver, packed)
        return packed


_packs = _PackCache()


def _slice_channels(y: torch.Tensor, c: int) -> torch.Tensor:
    if y.shape[1] == c:
        return y
    return y[:, :c].contiguous(memory_format=torch.channels_last)


def _raw_conv(x, w, stride, pad):
    y = _conv.conv2d_nhwc(_round_tf32_nhwc(x), _packs.get(w, 0), stride, pad)
    return _slice_channels(y, w.shape[0])


def _raw_grad_input(dy, w, stride, pad, in_hw):
    k = w.shape[2]
    if stride == 1:
        g = dy
    else:   # zero-dilate: dy'[2i, 2j] = dy[i, j]  (see module docstring)
        g = dy.new_zeros((dy.shape[0], dy.shape[1], in_hw[0], in_hw[1])).contiguous(
            memory_format=torch.channels_last)
        g[:, :, ::stride, ::stride][:, :, :dy.shape[2], :dy.shape[3]] = dy
    dx = _conv.conv2d_nhwc(_round_tf32_nhwc(g), _packs.get(w, 1), 1, k - 1 - pad)
    assert dx.shape[2:] == tuple(in_hw), (dx.shape, in_hw)
    return _slice_channels(dx, w.shape[1])


def _raw_grad_weight(dy, x, k, stride, pad):
    dw = _conv.conv2d_wgrad_nhwc(_round_tf32_nhwc(dy), _round_tf32_nhwc(x), k, stride, pad)
    return dw[:dy.shape[1], :x.shape[1]].contiguous()


class _Conv2d(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, stride, pad):
        ctx.save_for_backward(x, w)
        ctx.cfg = (stride, pad)
        return _raw_conv(x, w, stride, pad)

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        stride, pad = ctx.cfg
        dx = dw = None
        if ctx.needs_input_grad[0]:
            dx = _Conv2dGradInput.apply(dy, w, stride, pad, tuple(x.shape[2:]))
        if ctx.needs_input_grad[1]:
            dw = _Conv2dGradWeight.apply(dy, x, w.shape[2], stride, pad)
        return dx, dw, None, None


class _Conv2dGradInput(torch.autograd.Function):
    @staticmethod
    def forward(ctx, dy, w, stride, pad, in_hw):
        ctx.save_for_backward(dy, w)
        ctx.cfg = (stride, pad, in_hw)
        return _raw_grad_input(dy, w, stride, pad, in_hw)

    @staticmethod
    def backward(ctx, ddx):
        dy, w = ctx.saved_tensors
        stride, pad, in_hw = ctx.cfg
        g_dy = g_w = None
        if ctx.needs_input_grad[0]:
            g_dy = _Conv2d.apply(ddx, w, stride, pad)
        if ctx.needs_input_grad[1]:
            g_w = _Conv2dGradWeight.apply(dy, ddx, w.shape[2], stride, pad)
        return g_dy, g_w, None, None, None


class _Conv2dGradWeight(torch.autograd.Function):
    @staticmethod
    def forward(ctx, dy, x, k, stride, pad):
        ctx.save_for_backward(dy, x)
        ctx.cfg = (k, stride, pad)
        return _raw_grad_weight(dy, x, k, stride, pad)

    @staticmethod
    def backward(ctx, ddw):
        dy, x = ctx.saved_tensors
        k, stride, pad = ctx.cfg
        g_dy = g_x = None
        if ctx.needs_input_grad[0]:
            g_dy = _Conv2d.apply(x, ddw, stride, pad)
        if ctx.needs_input_grad[1]:
            g_x = _Conv2dGradInput.apply(dy, ddw, stride, pad, tuple(x.shape[2:]))
        return g_dy, g_x, None, None, None


def conv2d(x: torch.Tensor, weight: torch.Tensor, bias=None, stride: int = 1,
           padding: int = 0) -> torch.Tensor:
    """``F.conv2d(x, weight, bias, stride, padding)`` on the tcgen05 kernels.
    x (B,Cin,H,W) CUDA float; weight OIHW square kernel; returns channels_last."""
    _lib.require_cuda(x, "conv2d")
    assert weight.shape[2] == weight.shape[3], "square kernels only"
    y = _Conv2d.apply(x, weight, int(stride), int(padding))
    if bias is not None:
        y = y + bias.view(1, -1, 1, 1)
    return y
