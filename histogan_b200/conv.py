"""NHWC convolution primitives on the tcgen05 kernels (hg_conv2d_fwd).

Tensors are float32 torch tensors in channels_last memory format: logically
(B, C, H, W) like the reference, physically NHWC -- the layout the TMA boxes of
the kernels walk.  These are raw (non-autograd) calls; the differentiable
operators are built on top of them in ops.py.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib

CONV_LRELU = 1
CONV_ROUND_TF32 = 2


def tf32_round(t: torch.Tensor) -> torch.Tensor:
    """round-to-nearest-even to TF32 (10-bit mantissa); torch-side twin of
    hg::tf32_round used for small host-prepared tensors and in tests."""
    tc = t.contiguous()
    i = tc.view(torch.int32)
    r = (i + 0x0FFF + ((i >> 13) & 1)) & ~0x1FFF
    return torch.where(torch.isfinite(tc), r.view(torch.float32), tc).view_as(t)       # Inf / NaN pass through


def as_nhwc(x: torch.Tensor) -> torch.Tensor:
    return x.contiguous(memory_format=torch.channels_last)


def _up32(n):
    return (n + 31) // 32 * 32


PACK_FROM_OHWI = 2


def is_ohwi(w: torch.Tensor) -> bool:
    """is the (Cout,Cin,kh,kw) tensor stored channels_last (memory [Cout][kh][kw][Cin])?"""
    return w.dim() == 4 and w.is_contiguous(memory_format=torch.channels_last)


def pack_is_plain_copy(w: torch.Tensor) -> bool:
    """is the forward (mode 0) packed operand of `w` just its memory, TF32-rounded?  (channels_last
    parameter, both channel counts multiples of 32: nothing to transpose or pad)"""
    return is_ohwi(w) and w.shape[0] % 32 == 0 and w.shape[1] % 32 == 0


def pack_weight(w_oihw: torch.Tensor, mode: int = 0, out: torch.Tensor = None) -> torch.Tensor:
    """hg_pack_conv_weight: (Cout,Cin,kh,kw) -> [Np][KH][KW][Kp] K-major TF32, N/K zero-padded
    to multiples of 32.  mode 1 = dgrad (N = Cin, K = Cout, taps flipped).  The weight may be
    stored channels_last (what this package's modules do) or plain contiguous.  `out`: write
    into this existing packed tensor (same shape) instead of allocating."""
    lib = _lib.load()
    _lib.require_cuda(w_oihw, "pack_weight")
    w = w_oihw.detach().float()
    ohwi = is_ohwi(w)
    if not ohwi:
        w = w.contiguous()
    co, ci, kh, kw = w.shape
    n, k = (ci, co) if mode else (co, ci)
    shape = (_up32(n), kh, kw, _up32(k))
    if out is None:
        out = torch.empty(shape, dtype=torch.float32, device=w.device)
    else:
        assert tuple(out.shape) == shape and out.is_contiguous() and out.device == w.device
    if ohwi and mode == 0 and co % 32 == 0 and ci % 32 == 0:
        # channels_last parameter, nothing to pad: the packed forward operand IS the parameter's
        # memory, TF32-rounded -- one vectorised copy (hg_modulate_round without a modulation)
        with torch.cuda.device(w.device):
            rc = lib.hg_modulate_round(_lib.ptr(w), None, _lib.ptr(out), 1, co * kh * kw, ci, 1,
                                       _lib.current_stream_ptr(w.device))
        _lib.check(rc, "hg_modulate_round")
        return out
    with torch.cuda.device(w.device):
        rc = lib.hg_pack_conv_weight(_lib.ptr(w), _lib.ptr(out), co, ci, kh, kw,
                                     int(mode) | (PACK_FROM_OHWI if ohwi else 0),
                                     _lib.current_stream_ptr(w.device))
    _lib.check(rc, "hg_pack_conv_weight")
    return out


def conv2d_nhwc(x: torch.Tensor, w_packed: torch.Tensor, stride: int = 1, pad: int = 1, *,
                cout=None, scale=None, bias=None, noise=None, noise_w=None, noise_b=None,
                residual=None, lrelu: bool = False, slope: float = 0.2,
                round_tf32: bool = False, out_hw=None, into=None) -> torch.Tensor:
    """y = epilogue(conv(x, w)); x (B,Cin,H,W) channels_last (Cin % 4 == 0), w_packed
    [Cout_p][KH][KW][Cin_p] from pack_weight; `cout` = true number of output channels
    (% 4 == 0, default Cout_p).  Returns (B,cout,OH,OW) channels_last.
    out_hw: produce this output extent instead of the natural one (the excess windows read
    zeros); into=(t, oy, ox, step): write output pixel (oh, ow) to t[:, :, oy + step*oh,
    ox + step*ow] of an existing channels_last tensor t (B,cout,*,*) and return t."""
    lib = _lib.load()
    _lib.require_cuda(x, "conv2d_nhwc")
    assert x.dtype == torch.float32 and x.dim() == 4
    if not x.is_contiguous(memory_format=torch.channels_last):
        x = x.contiguous(memory_format=torch.channels_last)
    B, Cin, H, W = x.shape
    Cout_p, KH, KW, Cin_p = w_packed.shape
    assert Cin_p == _up32(Cin), (Cin_p, Cin)
    Cout = Cout_p if cout is None else int(cout)
    assert Cout <= Cout_p and _up32(Cout) == Cout_p, (Cout, Cout_p)
    OH = (H + 2 * pad - KH) // stride + 1
    OW = (W + 2 * pad - KW) // stride + 1
    if out_hw is not None:
        OH, OW = out_hw
    strides, y_ptr = (0, 0, 0), None
    if into is not None:
        y, oy, ox, step = into
        assert residual is None and y.is_contiguous(memory_format=torch.channels_last)
        assert y.shape[0] == B and y.shape[1] == Cout and y.dtype == torch.float32
        TH_, TW_ = y.shape[2], y.shape[3]
        assert oy + step * (OH - 1) < TH_ and ox + step * (OW - 1) < TW_
        strides = (TH_ * TW_ * Cout, step * TW_ * Cout, step * Cout)
        y_ptr = C.c_void_p(y.data_ptr() + 4 * (oy * TW_ + ox) * Cout)
    else:
        y = torch.empty((B, Cout, OH, OW), dtype=torch.float32, device=x.device,
                        memory_format=torch.channels_last)
    if residual is not None:
        residual = as_nhwc(residual)
        assert residual.shape == y.shape
    p = _lib.ConvParams(B, H, W, Cin, Cout, KH, KW, stride, pad, OH, OW)
    flags = (CONV_LRELU if lrelu else 0) | (CONV_ROUND_TF32 if round_tf32 else 0)
    keep = [t.contiguous() if t is not None else None for t in (scale, bias, noise, noise_w, noise_b)]
    ep = _lib.ConvEpilogue(
        keep[0].data_ptr() if keep[0] is not None else None,
        keep[1].data_ptr() if keep[1] is not None else None,
        keep[2].data_ptr() if keep[2] is not None else None,
        keep[3].data_ptr() if keep[3] is not None else None,
        keep[4].data_ptr() if keep[4] is not None else None,
        residual.data_ptr() if residual is not None else None,
        int(noise.shape[1]) if noise is not None else 0, flags, float(slope), 0, *strides)
    with torch.cuda.device(x.device):
        rc = lib.hg_conv2d_fwd(_lib.ptr(x), _lib.ptr(w_packed), y_ptr or _lib.ptr(y), C.byref(p),
                               C.byref(ep), _lib.current_stream_ptr(x.device))
    _lib.check(rc, "hg_conv2d_fwd")
    return y


def conv2d_wgrad_nhwc(dy: torch.Tensor, x: torch.Tensor, ksize: int, stride: int = 1,
                      pad: int = 1, out: torch.Tensor = None) -> torch.Tensor:
    """dW (Cout,Cin,k,k) from dy (B,Cout,OH,OW) and x (B,Cin,H,W), both channels_last float32
    (hg_conv2d_wgrad).  The result is stored channels_last -- the kernel's own
    [Cout][k][k][Cin] output viewed as (Cout,Cin,k,k) -- which is also how the modules store
    their weights, so no layout conversion happens anywhere on the gradient path."""
    lib = _lib.load()
    _lib.require_cuda(x, "conv2d_wgrad_nhwc")
    dy, x = as_nhwc(dy), as_nhwc(x)
    B, Cin, H, W = x.shape
    _, Cout, OH, OW = dy.shape
    p = _lib.ConvParams(B, H, W, Cin, Cout, ksize, ksize, stride, pad, OH, OW)
    # `out`: a (Cout,Cin,k,k) channels_last tensor to write into (a slot of the trainer's flat
    # gradient arena); usable when the kernel's [Cout][k][k][Cin_p] output has no channel padding
    dwp = None
    if out is not None and _up32(Cin) == Cin and tuple(out.shape) == (Cout, Cin, ksize, ksize) \
            and out.is_contiguous(memory_format=torch.channels_last) and out.dtype == torch.float32:
        dwp = out.permute(0, 2, 3, 1)
    if dwp is None:
        dwp = torch.empty((Cout, ksize, ksize, _up32(Cin)), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        st = _lib.current_stream_ptr(x.device)
        _lib.check(lib.hg_conv2d_wgrad(_lib.ptr(dy), _lib.ptr(x), _lib.ptr(dwp), C.byref(p), st),
                   "hg_conv2d_wgrad")
    dw = dwp.permute(0, 3, 1, 2)                     # (Cout, Cin_p, k, k), channels_last strides
    if dw.shape[1] != Cin:                           # drop the K padding (small-channel layers only)
        dw = dw[:, :Cin].contiguous(memory_format=torch.channels_last)
    return dw


# ---------------------------------------------------------------- image-input convolutions ----
def small_ok(cin: int, cout: int, k: int, stride: int, pad: int) -> bool:
    """can this conv run on the direct (CUDA-core) kernels for 3/4-channel inputs (conv_small.cu)?"""
    return cin <= 4 and cout % 4 == 0 and cout <= 64 and k in (1, 3) and stride == 1 and pad == k // 2


def conv_small_fwd(x, w, cp, bias=None, residual=None, lrelu=False, slope=0.2, round_tf32=False):
    """y (B,cp,H,W) channels_last = [round]([lrelu](conv(x, w) + bias) [+ residual]); x (B,Cin<=4,H,W)
    float32 with any strides (the planar image as it is), w (Cout,Cin,k,k)."""
    lib = _lib.load()
    B, Cin, H, W = x.shape
    Cout, _, k, _ = w.shape
    wc = w.detach().float().contiguous()                      # OIHW (a few hundred floats)
    y = torch.empty((B, cp, H, W), dtype=torch.float32, device=x.device, memory_format=torch.channels_last)
    res = as_nhwc(residual) if residual is not None else None
    if res is not None:
        assert tuple(res.shape) == tuple(y.shape)
    flags = (CONV_LRELU if lrelu else 0) | (CONV_ROUND_TF32 if round_tf32 else 0)
    b = bias.detach().float().contiguous() if bias is not None else None
    with torch.cuda.device(x.device):
        rc = lib.hg_conv_small_fwd(_lib.ptr(x), _lib.ptr(wc), _lib.ptr(b), _lib.ptr(res), _lib.ptr(y), B, Cin, H, W,
                                   Cout, cp, k, *x.stride(), flags, float(slope), _lib.current_stream_ptr(x.device))
    _lib.check(rc, "hg_conv_small_fwd")
    return y


def conv_small_dgrad(dy, w, cin):
    """dx (B,cin,H,W) contiguous = conv^T(dy, w); dy (B,Cp,H,W) channels_last, Cp >= Cout"""
    lib = _lib.load()
    dy = as_nhwc(dy)
    B, Cp, H, W = dy.shape
    Cout, _, k, _ = w.shape
    wc = w.detach().float().contiguous()
    dx = torch.empty((B, cin, H, W), dtype=torch.float32, device=dy.device)
    with torch.cuda.device(dy.device):
        rc = lib.hg_conv_small_dgrad(_lib.ptr(dy), _lib.ptr(wc), _lib.ptr(dx), B, cin, H, W, Cout, Cp, k,
                                     *dx.stride(), _lib.current_stream_ptr(dy.device))
    _lib.check(rc, "hg_conv_small_dgrad")
    return dx


def conv_small_wgrad(dy, x, wshape):
    """dw (Cout,Cin,k,k) from dy (B,Cp,H,W) channels_last and the image x (B,Cin,H,W) (any strides)"""
    lib = _lib.load()
    dy = as_nhwc(dy)
    B, Cp, H, W = dy.shape
    Cout, Cin, k, _ = wshape
    dw = torch.empty((Cout, Cin, k, k), dtype=torch.float32, device=dy.device)
    nws = lib.hg_conv_small_wgrad_workspace_bytes(Cin, Cout, k)
    ws = torch.empty(nws, dtype=torch.uint8, device=dy.device)
    xf = x if x.dtype == torch.float32 else x.float()
    with torch.cuda.device(dy.device):
        rc = lib.hg_conv_small_wgrad(_lib.ptr(dy), _lib.ptr(xf), _lib.ptr(dw), _lib.ptr(ws), nws, B, Cin, H, W,
                                     Cout, Cp, k, *xf.stride(), _lib.current_stream_ptr(dy.device))
    _lib.check(rc, "hg_conv_small_wgrad")
    return dw
