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

On top of them ``conv_bias_act`` fuses what a DiscriminatorBlock layer does around
its convolution (bias, LeakyReLU, residual sum, histoGAN/histoGAN.py:520-526) into
the conv epilogue, with a backward that is itself built from differentiable pieces.

Operands are rounded to TF32 (round-to-nearest-even) right before the MMA and
accumulated in fp32, which keeps the networks within ~3e-4 of the fp32 reference
(DESIGN.md "precision").  Producers that already emit TF32-rounded NHWC tensors
say so (``*_rounded`` flags) and the extra rounding pass is skipped.  The kernels take
channel counts that are multiples of 4 (their 32-channel boxes are completed by TMA's
out-of-bounds zero fill and zero-padded packed weights); this module pads activations
to multiples of 32 (see ``_CH``), which only affects D's 3- and 16-channel tensors.
"""
from __future__ import annotations

import ctypes as C

import contextlib

import torch

from . import _lib
from . import conv as _conv

# Activation channel padding.  The kernels accept any multiple of 4 (16-byte TMA rows; their
# 32-channel boxes are completed by TMA's out-of-bounds zero fill), but measured on B200 a
# 16-channel 256^2 conv takes exactly as long as the 32-channel one (the kernel is bound by the
# number of TMA boxes / MMAs per tile, not by bytes) and the 4-channel image path is slower, so
# D's 3/16-channel tensors are carried zero-padded to 32 channels.
import os as _os

# measured (round 2, bench_train_r2e*.json): padding D's 3/16-channel tensors to 16 instead of 32 leaves
# the conv kernels' times unchanged and halves every element-wise pass over the 256^2 tensors:
# 31.9 -> 30.3 ms per step
_CH = int(_os.environ.get("HG_CH_PAD", "16"))
SMALL_CIN = _os.environ.get("HG_SMALL_CIN", "1") != "0"   # image-input convs on the direct kernels


def _round_up(n, m=_CH):
    return (n + m - 1) // m * m


def _is_nhwc(x):
    return x.is_contiguous(memory_format=torch.channels_last)


def round_tf32_nhwc(x: torch.Tensor, already_rounded: bool = False) -> torch.Tensor:
    """fp32 channels_last copy of x, channels zero-padded to a multiple of _CH, TF32-rounded
    (no-op when the producer already delivered exactly that)."""
    lib = _lib.load()
    B, Cc, H, W = x.shape
    Cp = _round_up(Cc)
    if already_rounded and Cp == Cc and x.dtype == torch.float32 and _is_nhwc(x):
        return x
    if x.dtype != torch.float32:
        x = x.float()
    if Cp != Cc:
        # one pass: strided read of the Cc real channels, zero padding, rounding, NHWC write
        out = torch.empty((B, Cp, H, W), dtype=torch.float32, device=x.device,
                          memory_format=torch.channels_last)
        if out.numel():
            with torch.cuda.device(x.device):
                rc = lib.hg_pad_round_nhwc(_lib.ptr(x), _lib.ptr(out), B, Cc, H, W, Cp, *x.stride(),
                                           _lib.current_stream_ptr(x.device))
            _lib.check(rc, "hg_pad_round_nhwc")
        return out
    else:
        src = x if _is_nhwc(x) else x.contiguous(memory_format=torch.channels_last)
    out = torch.empty_like(src, memory_format=torch.channels_last)
    if src.numel():
        with torch.cuda.device(x.device):
            rc = lib.hg_modulate_round(_lib.ptr(src), None, _lib.ptr(out), B, H * W, Cp, 1,
                                       _lib.current_stream_ptr(x.device))
        _lib.check(rc, "hg_modulate_round")
    return out


class _RoundPad(torch.autograd.Function):
    """differentiable face of round_tf32_nhwc: the TF32 rounding is treated as identity
    (as it is inside the conv Functions), the channel padding as a zero-extension."""

    @staticmethod
    def forward(ctx, x):
        ctx.c = x.shape[1]
        return round_tf32_nhwc(x)

    @staticmethod
    def backward(ctx, g):
        return g if g.shape[1] == ctx.c else g[:, :ctx.c]


def round_pad(x: torch.Tensor) -> torch.Tensor:
    return _RoundPad.apply(x)


class _PackCache:
    """Packed (K-major, TF32) copies of a weight, kept ON the weight tensor object: a dict
    {mode: [version, packed]} under ``w._hg_packed``.  Nothing is keyed by address, so a
    freed-and-reallocated tensor can never hit a stale entry.

    A packed tensor is allocated ONCE per (weight, mode) and afterwards re-filled IN PLACE:
    its address is stable for the lifetime of the parameter, so captured CUDA graphs read it
    directly and contain no packing kernels.  Whoever modifies a parameter in place keeps the
    packs current: the fused DiffGrad kernel writes the forward operand itself and calls
    ``refresh`` for the other forms; any other modification (load_state_dict, a foreign
    optimiser) bumps ``_version`` and is caught by the next ``get`` / ``refresh_stale``."""

    ATTR = "_hg_packed"

    def _pack(self, w, mode, out=None):
        if mode == 'wsq':                   # (Cout,Cin) sum of squares over the taps: demodulation
            from . import fused
            return fused.weight_sqsum(w, out)
        if isinstance(mode, tuple):         # ('s2', py, px): one parity class of a stride-2 dgrad
            return _conv.pack_weight(_stride2_class_weight(w, mode[1], mode[2]), 0, out)
        return _conv.pack_weight(w, mode, out)     # zero-pads both extents to multiples of 32

    def get(self, w: torch.Tensor, mode) -> torch.Tensor:
        store = getattr(w, self.ATTR, None)
        hit = store.get(mode) if store is not None else None
        if hit is not None and hit[0] == w._version:
            return hit[1]
        packed = self._pack(w, mode, hit[1] if hit is not None else None)
        if w.is_leaf:                       # parameters persist; temporaries are not worth caching
            if store is None:
                store = {}
                try:
                    setattr(w, self.ATTR, store)
                except AttributeError:
                    return packed
            store[mode] = [w._version, packed]
        return packed

    def refresh(self, w: torch.Tensor, done=()):
        """`w` was just modified in place: refill every cached form (modes in `done` were already
        written by the caller) and stamp them with the new version."""
        store = getattr(w, self.ATTR, None)
        if not store:
            return
        for mode, hit in store.items():
            if mode not in done and hit[0] != w._version:
                self._pack(w, mode, hit[1])
            hit[0] = w._version

    def refresh_stale(self, weights):
        """bring the cached forms of `weights` up to date (cheap version check per tensor); called
        before a CUDA-graph replay, whose kernels read the packed tensors without going through get"""
        for w in weights:
            store = getattr(w, self.ATTR, None)
            if store and any(hit[0] != w._version for hit in store.values()):
                self.refresh(w)

    def fused_forward_target(self, w: torch.Tensor):
        """the cached forward operand of `w` if an optimiser kernel may write it directly
        (a plain TF32-rounded copy in the parameter's own linear order), else None"""
        store = getattr(w, self.ATTR, None)
        if store and 0 in store and w.dim() == 4 and _conv.pack_is_plain_copy(w) \
                and store[0][1].numel() == w.numel():
            return store[0][1]
        return None


# below 64x64 the four launches cost more than the dilated single conv (measured, B=32:
# 64 vs 55 us at 32x32, 82 vs 45 us at 16x16; 175 vs 539 us at 256x256)
_S2_MIN_PIXELS = 64 * 64
_S2_TAPS = ((1,), (2, 0))        # input parity 0 <- tap 1;  parity 1 <- taps 2 (offset 0), 0 (offset +1)


def _stride2_class_weight(w, py, px):
    """weight of the small convolution over dy that yields dx[:, :, py::2, px::2] for a 3x3 /
    stride 2 / pad 1 convolution with weight w (Cout,Cin,3,3): y[o] = sum_k x[2o+k-1] w[k]  =>
    dx[2a] = dy[a] w[1],  dx[2a+1] = dy[a] w[2] + dy[a+1] w[0]  (per axis)."""
    t = w.detach().permute(1, 0, 2, 3)
    # plain slices only (no index tensors: this also runs under CUDA-graph capture)
    t = torch.cat([t[:, :, k:k + 1] for k in _S2_TAPS[py]], dim=2)
    return torch.cat([t[:, :, :, k:k + 1] for k in _S2_TAPS[px]], dim=3)


_packs = _PackCache()


def _match_channels(y: torch.Tensor, c: int) -> torch.Tensor:
    """drop (or keep) the zero padding channels: `c` is what the caller's tensors carry"""
    if y.shape[1] == c:
        return y
    return y[:, :c].contiguous(memory_format=torch.channels_last)


# the three raw primitives (tests/emulation.py swaps these for torch stand-ins on CPU) ------

def _small(x_channels, w, stride, pad):
    """image-input conv (Cin <= 4, e.g. DiscriminatorBlock 0): direct CUDA-core kernels, no padding"""
    return x_channels == w.shape[1] and _conv.small_ok(w.shape[1], w.shape[0], w.shape[2], stride, pad)


def _raw_conv(x, w, stride, pad, x_rounded=False, padded_io=False):
    """padded_io: x / y carry round_up(C, 32) channels (zeros in the padding) instead of C"""
    if _small(x.shape[1], w, stride, pad):
        return _conv.conv_small_fwd(x, w, _round_up(w.shape[0]) if padded_io else w.shape[0])
    y = _conv.conv2d_nhwc(round_tf32_nhwc(x, x_rounded), _packs.get(w, 0), stride, pad,
                          cout=_round_up(w.shape[0]))
    return y if padded_io else _match_channels(y, w.shape[0])


def _raw_grad_input(dy, w, stride, pad, in_hw, dy_rounded=False, padded_io=False):
    k = w.shape[2]
    if _conv.small_ok(w.shape[1], w.shape[0], k, stride, pad) and tuple(in_hw) == tuple(dy.shape[2:]):
        return _conv.conv_small_dgrad(dy, w, w.shape[1])          # true Cin channels, planar
    if (stride == 2 and k == 3 and pad == 1 and in_hw[0] == 2 * dy.shape[2]
            and in_hw[1] == 2 * dy.shape[3] and in_hw[0] * in_hw[1] >= _S2_MIN_PIXELS):
        # four parity classes of dx, each a 1..4-tap convolution over dy written straight into
        # its interleaved positions: 9 tap-GEMMs instead of the 36 of a zero-dilated dy, and no
        # dilated copy of dy
        g = round_tf32_nhwc(dy, dy_rounded)
        cin_p = _round_up(w.shape[1])
        dx = torch.empty((dy.shape[0], cin_p, in_hw[0], in_hw[1]), dtype=torch.float32,
                         device=dy.device, memory_format=torch.channels_last)
        for py in (0, 1):
            for px in (0, 1):
                _conv.conv2d_nhwc(g, _packs.get(w, ('s2', py, px)), 1, 0, cout=cin_p,
                                  out_hw=(dy.shape[2], dy.shape[3]), into=(dx, py, px, 2))
        return dx if padded_io else _match_channels(dx, w.shape[1])
    if stride == 1:
        g = dy
    else:   # zero-dilate: dy'[2i, 2j] = dy[i, j]  (see module docstring)
        g = dy.new_zeros((dy.shape[0], dy.shape[1], in_hw[0], in_hw[1])).contiguous(
            memory_format=torch.channels_last)
        g[:, :, ::stride, ::stride][:, :, :dy.shape[2], :dy.shape[3]] = dy
    dx = _conv.conv2d_nhwc(round_tf32_nhwc(g, dy_rounded), _packs.get(w, 1), 1, k - 1 - pad,
                           cout=_round_up(w.shape[1]))
    assert dx.shape[2:] == tuple(in_hw), (dx.shape, in_hw)
    return dx if padded_io else _match_channels(dx, w.shape[1])


GRAD_SLOT = "_hg_grad_slot"      # trainer.GradArena: where this parameter's gradient should be written


_slot_epoch = [0]


def new_backward():
    """a new backward pass begins (the gradients were reset to None): slots may be handed out again"""
    _slot_epoch[0] += 1


def grad_slot(w):
    """the flat-arena view a weight's gradient should be produced in, or None (outside DDP, or when
    this backward already produced a gradient for `w` there: a second use of the weight must come in
    its own tensor and be ADDED by autograd)"""
    slot = getattr(w, GRAD_SLOT, None)
    if slot is None or getattr(w, "_hg_grad_slot_epoch", -1) == _slot_epoch[0]:
        return None
    w._hg_grad_slot_epoch = _slot_epoch[0]
    return slot


def _raw_grad_weight(dy, x, wshape, stride, pad, dy_rounded=False, x_rounded=False, out=None):
    k = wshape[2]
    if x.shape[1] == wshape[1] and _conv.small_ok(wshape[1], wshape[0], k, stride, pad):
        return _conv.conv_small_wgrad(dy, x, wshape).contiguous(memory_format=torch.channels_last)
    dw = _conv.conv2d_wgrad_nhwc(round_tf32_nhwc(dy, dy_rounded), round_tf32_nhwc(x, x_rounded), k,
                                 stride, pad, out=out)
    if tuple(dw.shape[:2]) != tuple(wshape[:2]):
        dw = dw[:wshape[0], :wshape[1]].contiguous(memory_format=torch.channels_last)
    return dw


# torch.autograd.grad(outputs, inputs=images, create_graph=True) -- the first half of the gradient penalty
# (histoGAN.py:156-163) -- needs d out / d x of every layer and nothing else, but a Python autograd.Function
# only sees the static `needs_input_grad` (the weights DO require grad) and would compute a weight and a bias
# gradient per layer that the engine then throws away: one discarded weight-gradient pass of the whole
# discriminator per penalty step.  The caller says so:
_INPUT_GRADS_ONLY = False


@contextlib.contextmanager
def input_grads_only():
    """inside: the backward of the conv ops returns gradients for their activations only"""
    global _INPUT_GRADS_ONLY
    prev, _INPUT_GRADS_ONLY = _INPUT_GRADS_ONLY, True
    try:
        yield
    finally:
        _INPUT_GRADS_ONLY = prev


class _Conv2d(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, stride, pad, x_rounded=False, padded_io=False):
        ctx.save_for_backward(x, w)
        ctx.cfg = (stride, pad, x_rounded, padded_io)
        return _raw_conv(x, w, stride, pad, x_rounded, padded_io)

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        stride, pad, x_rounded, padded_io = ctx.cfg
        dx = dw = None
        if ctx.needs_input_grad[0]:
            dx = _Conv2dGradInput.apply(dy, w, stride, pad, tuple(x.shape[2:]), False, padded_io)
        if ctx.needs_input_grad[1] and not _INPUT_GRADS_ONLY:
            dw = _Conv2dGradWeight.apply(dy, x, tuple(w.shape), stride, pad, False, x_rounded,
                                         None if torch.is_grad_enabled() else grad_slot(w))
        return dx, dw, None, None, None, None


class _Conv2dGradInput(torch.autograd.Function):
    @staticmethod
    def forward(ctx, dy, w, stride, pad, in_hw, dy_rounded=False, padded_io=False):
        ctx.save_for_backward(dy, w)
        ctx.cfg = (stride, pad, in_hw, dy_rounded, padded_io)
        return _raw_grad_input(dy, w, stride, pad, in_hw, dy_rounded, padded_io)

    @staticmethod
    def backward(ctx, ddx):
        dy, w = ctx.saved_tensors
        stride, pad, in_hw, dy_rounded, padded_io = ctx.cfg
        g_dy = g_w = None
        if ctx.needs_input_grad[0]:
            g_dy = _Conv2d.apply(ddx, w, stride, pad, False, padded_io)
        if ctx.needs_input_grad[1]:
            g_w = _Conv2dGradWeight.apply(dy, ddx, tuple(w.shape), stride, pad, dy_rounded, False)
        return g_dy, g_w, None, None, None, None, None


class _Conv2dGradWeight(torch.autograd.Function):
    @staticmethod
    def forward(ctx, dy, x, wshape, stride, pad, dy_rounded=False, x_rounded=False, out=None):
        ctx.save_for_backward(dy, x)
        ctx.cfg = (wshape, stride, pad, dy_rounded, x_rounded)
        if out is not None:
            return _raw_grad_weight(dy, x, wshape, stride, pad, dy_rounded, x_rounded, out=out)
        return _raw_grad_weight(dy, x, wshape, stride, pad, dy_rounded, x_rounded)

    @staticmethod
    def backward(ctx, ddw):
        dy, x = ctx.saved_tensors
        wshape, stride, pad, dy_rounded, x_rounded = ctx.cfg
        g_dy = g_x = None
        padded_io = (dy.shape[1], x.shape[1]) != (wshape[0], wshape[1])
        if ctx.needs_input_grad[0]:
            g_dy = _Conv2d.apply(x, ddw, stride, pad, x_rounded, padded_io)
        if ctx.needs_input_grad[1]:
            g_x = _Conv2dGradInput.apply(dy, ddw, stride, pad, tuple(x.shape[2:]), dy_rounded,
                                         padded_io)
        return g_dy, g_x, None, None, None, None, None, None


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


# ------------------------------------------------- fused conv + bias + act ------

class _BiasActBwd(torch.autograd.Function):
    """(dy, y) -> dpre = tf32_round(dy * lrelu'(y)), db = sum dpre   (hg_bias_act_bwd).
    Linear in dy, so its own backward is the same op."""

    @staticmethod
    def forward(ctx, dy, y, slope, want_db):
        lib = _lib.load()
        dy = dy if _is_nhwc(dy) else dy.contiguous(memory_format=torch.channels_last)
        B, Cc, H, W = dy.shape
        dpre = torch.empty_like(dy, memory_format=torch.channels_last)
        db = torch.empty((Cc,), dtype=torch.float32, device=dy.device) if want_db else None
        with torch.cuda.device(dy.device):
            rc = lib.hg_bias_act_bwd(_lib.ptr(dy), _lib.ptr(y), _lib.ptr(dpre), _lib.ptr(db), B, H * W,
                                     Cc, float(slope), _lib.current_stream_ptr(dy.device))
        _lib.check(rc, "hg_bias_act_bwd")
        ctx.save_for_backward(y)
        ctx.slope = slope
        ctx.shape = tuple(dy.shape)
        return dpre, db

    @staticmethod
    def backward(ctx, g_dpre, g_db):
        (y,) = ctx.saved_tensors
        g = g_dpre
        if g_db is not None:
            gb = g_db.view(1, -1, 1, 1)
            g = gb.expand(ctx.shape) if g is None else g + gb
        if g is None:
            return None, None, None, None
        ddy, _ = _BiasActBwd.apply(g, y, ctx.slope, False)
        return ddy, None, None, None


class _ConvBiasAct(torch.autograd.Function):
    """y = [LeakyReLU](conv(x, w) + b) [+ residual], optionally stored TF32-rounded --
    one kernel (hg_conv2d_fwd with fused epilogue).  Works on channel-padded tensors:
    x and y carry round_up(C, _CH) channels, the padding channels stay exactly zero."""

    @staticmethod
    def forward(ctx, x, w, b, res, stride, pad, act, slope, x_rounded, round_out):
        cout_p = _round_up(w.shape[0])
        if x.is_cuda and _small(x.shape[1], w, stride, pad):
            # the raw image: direct kernel with the same fused epilogue, nothing padded or rounded
            y = _conv.conv_small_fwd(x, w, cout_p, bias=b, residual=res, lrelu=act, slope=slope,
                                     round_tf32=round_out)
            ctx.save_for_backward(x, w, y if act else None)
            ctx.cfg = (stride, pad, act, slope, b is not None, res is not None, tuple(x.shape))
            return y
        xr = round_tf32_nhwc(x, x_rounded)
        bp = None
        if b is not None:
            bp = b.detach().float()
            if cout_p != w.shape[0]:
                bp = torch.nn.functional.pad(bp, (0, cout_p - w.shape[0]))
        y = _conv.conv2d_nhwc(xr, _packs.get(w, 0), stride, pad, cout=cout_p, bias=bp, residual=res,
                              lrelu=act, slope=slope, round_tf32=round_out)
        ctx.save_for_backward(xr, w, y if act else None)
        ctx.cfg = (stride, pad, act, slope, b is not None, res is not None, tuple(x.shape))
        return y

    @staticmethod
    def backward(ctx, dy):
        xr, w, y = ctx.saved_tensors
        stride, pad, act, slope, has_b, has_res, xshape = ctx.cfg
        need_x, need_w, need_b = ctx.needs_input_grad[0], ctx.needs_input_grad[1], \
            has_b and ctx.needs_input_grad[2]
        if _INPUT_GRADS_ONLY:
            need_w = need_b = False
        dx = dw = db = None
        if need_x or need_w or need_b:
            dpre, dbp = _BiasActBwd.apply(dy, y if act else None, slope if act else 1.0, need_b)
            if need_b:
                db = dbp[:w.shape[0]]
            if need_x:
                dx = _Conv2dGradInput.apply(dpre, w, stride, pad, tuple(xr.shape[2:]), True, True)
                if dx.shape[1] != xshape[1]:
                    dx = dx[:, :xshape[1]]
            if need_w:      # (under create_graph the result is part of a graph: never an arena slot)
                dw = _Conv2dGradWeight.apply(dpre, xr, tuple(w.shape), stride, pad, True, True,
                                             None if torch.is_grad_enabled() else grad_slot(w))
        dres = dy if (has_res and ctx.needs_input_grad[3]) else None
        return dx, dw, db, dres, None, None, None, None, None, None


def conv_bias_act(x, weight, bias=None, residual=None, stride=1, padding=0, act=False, slope=0.2,
                  x_rounded=False, round_out=False):
    """fused ``[leaky_relu](conv2d(x, weight, bias, stride, padding)) [+ residual]``.
    Returns a channels_last tensor with round_up(Cout, _CH) channels (padding channels are
    zero); `x` may itself be such a padded tensor."""
    _lib.require_cuda(x, "conv_bias_act")
    return _ConvBiasAct.apply(x, weight, bias, residual, int(stride), int(padding), bool(act),
                              float(slope), bool(x_rounded), bool(round_out))
