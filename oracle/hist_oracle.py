"""CPU oracle for the RGB-uv histogram block and the Hellinger histogram loss.

TEST INFRASTRUCTURE -- see oracle/__init__.py.  A restatement (not a copy) of

* ``histogram_classes/RGBuvHistBlock.py:75-228``  (forward)
* ``histoGAN/histoGAN.py:955-960``                (relu + Hellinger loss)

in vectorised torch-CPU code.  It keeps every numerically observable quirk of
the reference:

* clamp to [0,1] first (``:76``), resize only when H or W > insz (``:77``);
  'interpolation' = bilinear, align_corners=False (``:78-80``); 'sampling'
  keeps ``h`` (not insz!) rows/cols at ``int(linspace(0, H, h, endpoint=False))``
  (``:81-89``)
* channels >= 3 dropped (``:98-99``)
* ``Iy = sqrt(R^2+G^2+B^2+EPS)`` (``:105-110``), logs and their differences in
  float32 (``:112-115``)
* bin centres ``np.linspace(lo, hi, h)`` are float64, so the difference, the
  square, the division by sigma^2 and the kernel run in float64 and are
  rounded to float32 afterwards (``:116-146``)
* ``hist[c] = (Iy * Ku).T @ Kv`` as a float32 matmul (``:147-148``)
* joint normalisation over (c,u,v) with ``+EPS`` (``:225-226``)
* ``green_only`` stores the G-channel histogram at index 0 (``:184-187``)

Gradients come from autograd on this restatement (the reference has no
hand-written backward either).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

EPS = 1e-6  # RGBuvHistBlock.py:25

METHODS = ("thresholding", "RBF", "inverse-quadratic")
RESIZINGS = ("interpolation", "sampling")


def bin_centres(lo: float, hi: float, h: int) -> np.ndarray:
    """float64 bin centres exactly as the reference builds them (``:117-118``)."""
    return np.linspace(lo, hi, num=h)


def _soft_assign(u: torch.Tensor, centres: torch.Tensor, method: str,
                 sigma: float, thr: float) -> torch.Tensor:
    """(N,) float32 log-chroma -> (N,h) float32 kernel matrix (``:116-146``)."""
    d = (u.unsqueeze(1) - centres.unsqueeze(0)).abs()        # float64 (N,h)
    if method == "thresholding":
        k = d <= thr / 2
    elif method == "RBF":
        k = torch.exp(-(torch.pow(d, 2) / sigma ** 2))
    elif method == "inverse-quadratic":
        k = 1 / (1 + torch.pow(d, 2) / sigma ** 2)
    else:
        raise Exception(
            f"Wrong kernel method. It should be either thresholding, RBF,"
            f" inverse-quadratic. But the given value is {method}.")
    return k.type(torch.float32)


def preprocess(x: torch.Tensor, h: int, insz: int, resizing: str) -> torch.Tensor:
    """clamp + optional resize + channel truncation (``:76-99``)."""
    x = torch.clamp(x, 0, 1)
    if x.shape[2] > insz or x.shape[3] > insz:
        if resizing == "interpolation":
            x = F.interpolate(x, size=(insz, insz), mode="bilinear",
                              align_corners=False)
        elif resizing == "sampling":
            i1 = torch.LongTensor(np.linspace(0, x.shape[2], h, endpoint=False))
            i2 = torch.LongTensor(np.linspace(0, x.shape[3], h, endpoint=False))
            x = x.index_select(2, i1).index_select(3, i2)
        else:
            raise Exception(
                f"Wrong resizing method. It should be: interpolation or sampling. "
                f"But the given value is {resizing}.")
    if x.shape[1] > 3:
        x = x[:, :3]
    return x


# (u-partner, v-partner) of each output channel: R:(G,B)  G:(R,B)  B:(R,G)
_PARTNERS = ((1, 2), (0, 2), (0, 1))


def rgb_uv_hist_raw(x: torch.Tensor, h: int = 64, insz: int = 150,
                    resizing: str = "interpolation",
                    method: str = "inverse-quadratic", sigma: float = 0.02,
                    intensity_scale: bool = True, hist_boundary=None,
                    green_only: bool = False, log_fn=torch.log,
                    preprocessed: bool = False) -> torch.Tensor:
    """Un-normalised histograms (B, 3|1, h, h), float32.

    ``log_fn`` / ``preprocessed`` exist for the parity tests only: they let the
    test feed the oracle the float32 logs / resized pixels produced on the GPU so
    that kernel arithmetic can be checked separately from libm / resize rounding
    (DESIGN.md "precision").  Defaults reproduce the reference exactly."""
    if hist_boundary is None:
        hist_boundary = [-3, 3]
    lo, hi = sorted(hist_boundary)
    thr = (abs(lo) + abs(hi)) / h                           # ``:70-71``
    centres = torch.tensor(bin_centres(lo, hi, h))          # float64
    xs = x if preprocessed else preprocess(x, h, insz, resizing)
    chans = (1,) if green_only else (0, 1, 2)
    out = []
    for img in torch.unbind(xs, dim=0):
        I = img.reshape(3, -1)                               # (3,N)
        II = torch.pow(I, 2)
        if intensity_scale:
            Iy = torch.sqrt(II[0] + II[1] + II[2] + EPS).unsqueeze(1)
        else:
            Iy = 1
        logI = log_fn(I + EPS)
        hs = []
        for c in chans:
            a, b = _PARTNERS[c]
            Ku = _soft_assign(logI[c] - logI[a], centres, method, sigma, thr)
            Kv = _soft_assign(logI[c] - logI[b], centres, method, sigma, thr)
            hs.append(torch.mm(torch.t(Iy * Ku), Kv))
        out.append(torch.stack(hs, dim=0))
    return torch.stack(out, dim=0)


def rgb_uv_hist(x: torch.Tensor, **kw) -> torch.Tensor:
    """Normalised histograms, == ``RGBuvHistBlock(**kw, device='cpu')(x)``."""
    raw = rgb_uv_hist_raw(x, **kw)
    # same reduction order as ``:225-226`` (three successive sums)
    return raw / (raw.sum(dim=1).sum(dim=1).sum(dim=1).view(-1, 1, 1, 1) + EPS)


def rg_chroma_hist(x: torch.Tensor, h: int = 64, insz: int = 150, resizing: str = "interpolation",
                   method: str = "inverse-quadratic", sigma: float = 0.02,
                   intensity_scale: bool = False, hist_boundary=None) -> torch.Tensor:
    """Restatement of ``histogram_classes/rgChromaHistBlock.py:75-131``: one-channel soft
    histogram of (R, G) / (R+G+B+EPS); same float64 soft-binning / float32 matmul structure."""
    if hist_boundary is None:
        hist_boundary = [0, 1]
    lo, hi = sorted(hist_boundary)
    thr = (abs(lo) + abs(hi)) / h
    centres = torch.tensor(bin_centres(lo, hi, h))
    xs = preprocess(x, h, insz, resizing)
    out = []
    for img in torch.unbind(xs, dim=0):
        I = torch.t(img.reshape(3, -1))                      # (N,3)
        II = torch.pow(I, 2)
        Iy = torch.sqrt(II[:, 0] + II[:, 1] + II[:, 2] + EPS).unsqueeze(1) if intensity_scale else 1
        ssum = torch.sum(I, dim=-1) + EPS
        Ku = _soft_assign(I[:, 0] / ssum, centres, method, sigma, thr)
        Kv = _soft_assign(I[:, 1] / ssum, centres, method, sigma, thr)
        out.append(torch.mm(torch.t(Iy * Ku), Kv).unsqueeze(0))
    raw = torch.stack(out, dim=0)
    return raw / (raw.sum(dim=1).sum(dim=1).sum(dim=1).view(-1, 1, 1, 1) + EPS)


def lab_hist(x: torch.Tensor, h: int = 64, insz: int = 150, resizing: str = "interpolation",
             method: str = "inverse-quadratic", sigma: float = 0.02,
             intensity_scale: bool = False, hist_boundary=None) -> torch.Tensor:
    """Restatement of ``histogram_classes/LabHistBlock.py:73-145``: (a, b) soft histogram of an
    image already in Lab/[0,1], weighted by L when intensity_scale."""
    if hist_boundary is None:
        hist_boundary = [0, 1]
    lo, hi = sorted(hist_boundary)
    thr = (abs(lo) + abs(hi)) / h
    centres = torch.tensor(bin_centres(lo, hi, h))
    xs = preprocess(x, h, insz, resizing)
    out = []
    for img in torch.unbind(xs, dim=0):
        I = torch.t(img.reshape(3, -1))
        Il = I[:, 0].unsqueeze(1) if intensity_scale else 1
        Ka = _soft_assign(I[:, 1], centres, method, sigma, thr)
        Kb = _soft_assign(I[:, 2], centres, method, sigma, thr)
        out.append(torch.mm(torch.t(Il * Ka), Kb).unsqueeze(0))
    raw = torch.stack(out, dim=0)
    return raw / (raw.sum(dim=1).sum(dim=1).sum(dim=1).view(-1, 1, 1, 1) + EPS)


SCALE = 1 / np.sqrt(2.0)  # histoGAN/histoGAN.py:54


def hellinger_loss(target: torch.Tensor, generated: torch.Tensor,
                   alpha: float = 2.0) -> torch.Tensor:
    """``alpha*SCALE*sqrt(sum((sqrt(T)-sqrt(H))^2))/B`` (histoGAN.py:957-960)."""
    return alpha * SCALE * (torch.sqrt(torch.sum(torch.pow(
        torch.sqrt(target) - torch.sqrt(generated), 2)))) / target.shape[0]


def hist_loss_and_grad(x: torch.Tensor, target: torch.Tensor, alpha: float = 2.0,
                       apply_relu: bool = True, **kw):
    """The G-phase snippet ``histBlock(F.relu(img))`` + loss + d loss/d img
    (histoGAN.py:955-960), returned as (hist, loss, grad_x)."""
    x = x.detach().clone().requires_grad_(True)
    hist = rgb_uv_hist(F.relu(x) if apply_relu else x, **kw)
    loss = hellinger_loss(target, hist, alpha)
    (gx,) = torch.autograd.grad(loss, x)
    return hist.detach(), loss.detach(), gx


def hist_linear_grad(x: torch.Tensor, weights: torch.Tensor, apply_relu: bool = True, **kw):
    """d <hist, weights> / dx -- a finite upstream gradient for every method."""
    x = x.detach().clone().requires_grad_(True)
    hist = rgb_uv_hist(F.relu(x) if apply_relu else x, **kw)
    (gx,) = torch.autograd.grad((hist * weights).sum(), x)
    return gx


# ---------------------------------------------------------------------------
# synthetic inputs shared by tests / bench
# ---------------------------------------------------------------------------

def synth_uniform(B, S, seed=0, C=3):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(B, C, S, S, generator=g)


def synth_generator_like(B, S, seed=0, C=3):
    """relu(randn*0.5+0.3): ~27 % exact zeros, ~8 % above 1."""
    g = torch.Generator().manual_seed(seed)
    return torch.relu(torch.randn(B, C, S, S, generator=g) * 0.5 + 0.3)


def synth_signed(B, S, seed=0, C=3):
    """randn*0.5+0.4: exercises both clamp sides and the outer relu."""
    g = torch.Generator().manual_seed(seed)
    return torch.randn(B, C, S, S, generator=g) * 0.5 + 0.4


def synth_random_target(B, h=64, seed=1, nc=3):
    g = torch.Generator().manual_seed(seed)
    t = torch.rand(B, nc, h, h, generator=g)
    return t / t.sum(dim=(1, 2, 3), keepdim=True)
