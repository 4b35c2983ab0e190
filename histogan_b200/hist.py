"""RGB-uv histogram block and Hellinger histogram loss -- host-side mirror of the
reference operator API on top of the C ABI.

``RGBuvHistBlock`` keeps the constructor kwargs, ``forward`` signature, output
shape/dtype and error behaviour of ``histogram_classes/RGBuvHistBlock.py:28-228``
so that ``histoGAN/histoGAN.py:737-741,955`` and ``ReHistoGAN/rehistoGAN.py``
can use it unchanged.  All arithmetic happens in
``libhistogan_b200.so``; there is no PyTorch/CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn
from torch.autograd.function import once_differentiable

from . import _lib

EPS = 1e-6


def _make_params(x: torch.Tensor, h, insz, resizing_id, method_id, sigma, lo, hi,
                 intensity_scale, green_only, projection=0) -> _lib.HistParams:
    B, Cc, H, W = x.shape
    sb, sc, sh, sw = x.stride()
    return _lib.HistParams(B, Cc, H, W, sb, sc, sh, sw, int(h), int(insz), resizing_id,
                           method_id, float(sigma), float(lo), float(hi),
                           int(bool(intensity_scale)), int(bool(green_only)), int(projection))


def _workspace(nbytes: int, device) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)


class _RGBuvHistFn(torch.autograd.Function):
    """hg_hist_fwd / hg_hist_bwd."""

    @staticmethod
    def forward(ctx, x, cfg):
        lib = _lib.load()
        params = _make_params(x, *cfg)
        nc = 1 if (cfg[8] or (len(cfg) > 9 and cfg[9])) else 3     # green_only or rg-chroma
        h = int(cfg[0])
        hist = torch.empty((x.shape[0], nc, h, h), dtype=torch.float32, device=x.device)
        hist_sum = torch.empty((x.shape[0],), dtype=torch.float32, device=x.device)
        if x.shape[0] > 0:
            ws_bytes = lib.hg_hist_fwd_workspace_bytes(C.byref(params))
            ws = _workspace(ws_bytes, x.device)
            with torch.cuda.device(x.device):
                rc = lib.hg_hist_fwd(_lib.ptr(x), C.byref(params), _lib.ptr(hist),
                                     _lib.ptr(hist_sum), _lib.ptr(ws), ws.numel(),
                                     _lib.current_stream_ptr(x.device))
            _lib.check(rc, "hg_hist_fwd")
        ctx.cfg = cfg
        ctx.save_for_backward(x, hist, hist_sum)
        return hist

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_hist):
        x, hist, hist_sum = ctx.saved_tensors
        lib = _lib.load()
        params = _make_params(x, *ctx.cfg)
        grad_hist = grad_hist.contiguous().float()
        # grad_x shares x's strides so one hg_hist_params describes both
        grad_x = torch.empty_strided(x.shape, x.stride(), dtype=torch.float32, device=x.device)
        if x.shape[0] > 0:
            ws_bytes = lib.hg_hist_bwd_workspace_bytes(C.byref(params))
            ws = _workspace(ws_bytes, x.device)
            with torch.cuda.device(x.device):
                rc = lib.hg_hist_bwd(_lib.ptr(x), C.byref(params), _lib.ptr(hist),
                                     _lib.ptr(hist_sum), _lib.ptr(grad_hist), _lib.ptr(grad_x),
                                     _lib.ptr(ws), ws.numel(), _lib.current_stream_ptr(x.device))
            _lib.check(rc, "hg_hist_bwd")
        return grad_x, None


def _kernel_addressable(x: torch.Tensor) -> bool:
    """The kernels take arbitrary element strides, but grad_x is allocated with x's
    strides, so they must address distinct elements: NCHW or NHWC dense layouts."""
    return x.is_contiguous() or x.is_contiguous(memory_format=torch.channels_last)


class RGBuvHistBlock(nn.Module):
    """Drop-in for ``histogram_classes.RGBuvHistBlock.RGBuvHistBlock``.

    Args (identical to RGBuvHistBlock.py:29-57): h, insz, resizing
    ('interpolation' | 'sampling'), method ('thresholding' | 'RBF' |
    'inverse-quadratic'), sigma, intensity_scale, hist_boundary, green_only,
    device ('cuda', 'cuda:N' or an int; a CPU device is rejected at forward
    time -- the CPU restatement is test-only, see oracle/).

    forward(x): float (B, C>=3, H, W) -> float32 (B, 3 or 1, h, h), normalised
    per image; differentiable w.r.t. x unless method == 'thresholding'.
    """

    def __init__(self, h=64, insz=150, resizing='interpolation',
                 method='inverse-quadratic', sigma=0.02, intensity_scale=True,
                 hist_boundary=None, green_only=False, device='cuda'):
        super().__init__()
        self.h = h
        self.insz = insz
        self.device = device
        self.resizing = resizing
        self.method = method
        self.intensity_scale = intensity_scale
        self.green_only = green_only
        if hist_boundary is None:
            hist_boundary = [-3, 3]
        hist_boundary.sort()                    # in place, as RGBuvHistBlock.py:68
        self.hist_boundary = hist_boundary
        if self.method == 'thresholding':
            self.eps = (abs(hist_boundary[0]) + abs(hist_boundary[1])) / h
        else:
            self.sigma = sigma

    def _torch_device(self) -> torch.device:
        d = self.device
        if isinstance(d, int):
            return torch.device('cuda', d)
        return torch.device(d)

    def forward(self, x):
        dev = self._torch_device()
        if dev.type != 'cuda':
            raise RuntimeError(
                f"histogan_b200.RGBuvHistBlock runs on CUDA (sm_100a) only; got device="
                f"{self.device!r}. The CPU restatement is test infrastructure (oracle/).")
        _lib.require_cuda(x, "RGBuvHistBlock.forward")
        if x.dim() != 4 or x.shape[1] < 3:
            raise RuntimeError(f"expected (B, C>=3, H, W) input, got {tuple(x.shape)}")
        needs_resize = x.shape[2] > self.insz or x.shape[3] > self.insz
        if needs_resize and self.resizing not in _lib.RESIZE_IDS:
            raise Exception(
                f'Wrong resizing method. It should be: interpolation or sampling. '
                f'But the given value is {self.resizing}.')
        if self.method not in _lib.METHOD_IDS:
            raise Exception(
                f'Wrong kernel method. It should be either thresholding, RBF,'
                f' inverse-quadratic. But the given value is {self.method}.')
        if x.dtype != torch.float32:
            x = x.float()
        if not _kernel_addressable(x):
            x = x.contiguous()
        if dev.index is not None and x.device != dev:
            # the reference allocates the output on self.device (RGBuvHistBlock.py:101-102)
            raise RuntimeError(f"input on {x.device} but block constructed for {dev}")
        cfg = (self.h, self.insz, _lib.RESIZE_IDS.get(self.resizing, 0),
               _lib.METHOD_IDS[self.method], getattr(self, 'sigma', 1.0),
               self.hist_boundary[0], self.hist_boundary[1], self.intensity_scale,
               self.green_only, getattr(self, 'PROJECTION', 0))
        return _RGBuvHistFn.apply(x, cfg)


class rgChromaHistBlock(RGBuvHistBlock):
    """Drop-in for ``histogram_classes.rgChromaHistBlock.rgChromaHistBlock``:
    one-channel soft histogram of the rg chromaticity u = R/(R+G+B+eps), v = G/(R+G+B+eps)
    (rgChromaHistBlock.py:53-127); default boundary [0, 1], intensity_scale False.  Shares the
    generic CUDA kernels of the RGB-uv block (float64 soft-binning like the reference).
    forward(x): (B, C>=3, H, W) -> (B, 1, h, h)."""

    PROJECTION = 1

    def __init__(self, h=64, insz=150, resizing='interpolation', method='inverse-quadratic',
                 sigma=0.02, intensity_scale=False, hist_boundary=None, device='cuda'):
        if hist_boundary is None:
            hist_boundary = [0, 1]
        super().__init__(h=h, insz=insz, resizing=resizing, method=method, sigma=sigma,
                         intensity_scale=intensity_scale, hist_boundary=hist_boundary,
                         green_only=False, device=device)


class LabHistBlock(rgChromaHistBlock):
    """Drop-in for ``histogram_classes.LabHistBlock.LabHistBlock``: one-channel soft
    histogram over the (a, b) planes of an image already in Lab scaled to [0, 1], weighted by L
    when ``intensity_scale`` (LabHistBlock.py:73-145).  forward(x) -> (B, 1, h, h)."""

    PROJECTION = 2


class _HellingerFn(torch.autograd.Function):
    """hg_hellinger_fwd / hg_hellinger_bwd."""

    @staticmethod
    def forward(ctx, target, generated, alpha):
        lib = _lib.load()
        loss = torch.empty((), dtype=torch.float32, device=generated.device)
        q = torch.empty((), dtype=torch.float32, device=generated.device)
        ws = _workspace(lib.hg_hellinger_workspace_bytes(), generated.device)
        with torch.cuda.device(generated.device):
            rc = lib.hg_hellinger_fwd(_lib.ptr(target), _lib.ptr(generated), generated.numel(),
                                      generated.shape[0], float(alpha), _lib.ptr(loss),
                                      _lib.ptr(q), _lib.ptr(ws), ws.numel(),
                                      _lib.current_stream_ptr(generated.device))
        _lib.check(rc, "hg_hellinger_fwd")
        ctx.alpha = float(alpha)
        ctx.save_for_backward(target, generated, q)
        return loss

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_loss):
        target, generated, q = ctx.saved_tensors
        lib = _lib.load()
        need_t, need_g = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        g_t = torch.empty_like(target) if need_t else None
        g_g = torch.empty_like(generated) if need_g else None
        grad_loss = grad_loss.contiguous().float()
        with torch.cuda.device(generated.device):
            rc = lib.hg_hellinger_bwd(_lib.ptr(target), _lib.ptr(generated), generated.numel(),
                                      generated.shape[0], ctx.alpha, _lib.ptr(q),
                                      _lib.ptr(grad_loss), _lib.ptr(g_g), _lib.ptr(g_t),
                                      _lib.current_stream_ptr(generated.device))
        _lib.check(rc, "hg_hellinger_bwd")
        return g_t, g_g, None


def hellinger_loss(target: torch.Tensor, generated: torch.Tensor, alpha: float = 2.0):
    """``alpha * SCALE * sqrt(sum((sqrt(target) - sqrt(generated))**2)) / B`` --
    the histogram loss of histoGAN/histoGAN.py:957-960 as one fused op."""
    _lib.require_cuda(generated, "hellinger_loss")
    _lib.require_cuda(target, "hellinger_loss")
    if target.shape != generated.shape:
        raise RuntimeError(f"shape mismatch {tuple(target.shape)} vs {tuple(generated.shape)}")
    return _HellingerFn.apply(target.contiguous().float(), generated.contiguous().float(), alpha)


# ------------------------------------------------------------- test hooks ----

def hist_preprocess(x: torch.Tensor, h=64, insz=150, resizing='interpolation') -> torch.Tensor:
    """(B,3,N) pre-processed pixels as the kernels see them (hg_hist_preprocess)."""
    lib = _lib.load()
    _lib.require_cuda(x, "hist_preprocess")
    x = x.float()
    params = _make_params(x, h, insz, _lib.RESIZE_IDS[resizing], 2, 0.02, -3.0, 3.0, True, False)
    n = lib.hg_hist_num_pixels(C.byref(params))
    if n < 0:
        _lib.check(int(n), "hg_hist_num_pixels")
    out = torch.empty((x.shape[0], 3, int(n)), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        rc = lib.hg_hist_preprocess(_lib.ptr(x), C.byref(params), _lib.ptr(out),
                                    _lib.current_stream_ptr(x.device))
    _lib.check(rc, "hg_hist_preprocess")
    return out


def device_logf(v: torch.Tensor) -> torch.Tensor:
    """The float32 log the kernels use (hg_debug_logf)."""
    lib = _lib.load()
    _lib.require_cuda(v, "device_logf")
    v = v.contiguous().float()
    out = torch.empty_like(v)
    with torch.cuda.device(v.device):
        rc = lib.hg_debug_logf(_lib.ptr(v), _lib.ptr(out), v.numel(),
                               _lib.current_stream_ptr(v.device))
    _lib.check(rc, "hg_debug_logf")
    return out
