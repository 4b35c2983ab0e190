"""TEST-ONLY stand-ins for the three raw convolution primitives of
histogan_b200.ops, used on the CPU to check the HOST logic (module structure,
activation-side modulation algebra, the any-order autograd formulas) without a
GPU, and to measure how much of a GPU-vs-fp32 difference is TF32 rounding.

They are monkeypatched into ops for the duration of a test; the product never
imports this file and has no CPU path of its own."""
import contextlib

import torch
import torch.nn.functional as F

from histogan_b200 import _lib, ops


def tf32(t):
    tc = t.contiguous()
    i = tc.view(torch.int32)
    r = ((i + 0x0FFF + ((i >> 13) & 1)) & ~0x1FFF).view(torch.float32)
    return torch.where(torch.isfinite(tc), r, tc).view_as(t)


@contextlib.contextmanager
def emulated_conv(round_operands: bool):
    r = tf32 if round_operands else (lambda t: t)

    def raw_conv(x, w, stride, pad, x_rounded=False, padded_io=False):
        return F.conv2d(r(x.float()), r(w.detach().float()), stride=stride, padding=pad)

    def raw_grad_input(dy, w, stride, pad, in_hw, dy_rounded=False, padded_io=False):
        k = w.shape[2]
        out_pad = (in_hw[0] + 2 * pad - k) % stride, (in_hw[1] + 2 * pad - k) % stride
        return F.conv_transpose2d(r(dy.float()), r(w.detach().float()), stride=stride, padding=pad,
                                  output_padding=out_pad)

    def raw_grad_weight(dy, x, wshape, stride, pad, dy_rounded=False, x_rounded=False):
        return torch.nn.grad.conv2d_weight(r(x.float()), tuple(wshape), r(dy.float()),
                                           stride=stride, padding=pad)

    saved = (ops._raw_conv, ops._raw_grad_input, ops._raw_grad_weight, _lib.require_cuda)
    ops._raw_conv, ops._raw_grad_input, ops._raw_grad_weight = raw_conv, raw_grad_input, raw_grad_weight
    _lib.require_cuda = lambda t, what: None
    try:
        yield
    finally:
        ops._raw_conv, ops._raw_grad_input, ops._raw_grad_weight, _lib.require_cuda = saved
