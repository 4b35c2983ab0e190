"""Fused generator-layer operators: one autograd Function per reference sub-module call
chain, each a handful of kernel launches.

  mod_conv_layer   x -> LeakyReLU(Conv2DMod(x, style) + noise)      (histoGAN.py:465-476)
                   = hg_modulate_round -> hg_conv2d_fwd (demod / noise / lrelu epilogue)
                   backward = hg_modconv_epilogue_bwd -> hg_conv2d_fwd (dgrad) ->
                              hg_modulate_bwd ; hg_conv2d_wgrad
  to_rgb           RGBBlock's 1x1 modulated conv + skip              (histoGAN.py:380-386)
                   = hg_torgb_fwd / hg_torgb_bwd

First-order gradients only (the generator is never differentiated twice: the gradient
penalty acts on the discriminator, the path-length regulariser is first order).
"""
from __future__ import annotations

import ctypes as C

import torch
from torch.autograd.function import once_differentiable

from . import _lib
from . import conv as _conv
from . import ops


def _st(dev):
    return _lib.current_stream_ptr(dev)


def _modulate_round(x, mod):
    lib = _lib.load()
    x = x if x.is_contiguous(memory_format=torch.channels_last) else \
        x.contiguous(memory_format=torch.channels_last)
    B, Cc, H, W = x.shape
    out = torch.empty_like(x, memory_format=torch.channels_last)
    with torch.cuda.device(x.device):
        rc = lib.hg_modulate_round(_lib.ptr(x), _lib.ptr(mod), _lib.ptr(out), B, H * W, Cc, 1,
                                   _st(x.device))
    _lib.check(rc, "hg_modulate_round")
    return x, out


def _upsample_modulate_round(x, mod):
    lib = _lib.load()
    x = x if x.is_contiguous(memory_format=torch.channels_last) else \
        x.contiguous(memory_format=torch.channels_last)
    B, Cc, H, W = x.shape
    out = torch.empty((B, Cc, 2 * H, 2 * W), dtype=torch.float32, device=x.device,
                      memory_format=torch.channels_last)
    with torch.cuda.device(x.device):
        rc = lib.hg_upsample_modulate_round(_lib.ptr(x), _lib.ptr(mod), _lib.ptr(out), B, H, W, Cc,
                                            _st(x.device))
    _lib.check(rc, "hg_upsample_modulate_round")
    return x, out


class _ModConvLayer(torch.autograd.Function):
    """wsq is None: `d` is a differentiable input (its gradient is returned).  wsq given:
    d = rsqrt(mod^2 wsq^T + eps) was computed outside autograd (demod_all) and its dependence on
    (mod, w) is differentiated here analytically: hg_demod_bwd adds d loss/d d * d d/d mod to the
    modulation gradient and 2 w (sum_b t mod^2) to the weight gradient, in place."""

    @staticmethod
    def forward(ctx, x, mod, w, d, inoise, nw, nb, slope, upsample, act=True, wsq=None):
        k = w.shape[2]
        pad = (k - 1) // 2
        mod = mod.contiguous()
        x, xm = (_upsample_modulate_round if upsample else _modulate_round)(x.float(), mod)
        ctx.upsample = upsample
        y = _conv.conv2d_nhwc(xm, ops._packs.get(w, 0), 1, pad, cout=w.shape[0], scale=d,
                              noise=inoise, noise_w=nw, noise_b=nb, lrelu=act, slope=slope)
        ctx.save_for_backward(x, xm, mod, w, d, inoise, nw, nb, y, wsq)
        ctx.slope = slope if act else 1.0       # no activation == LeakyReLU with slope 1
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        x, xm, mod, w, d, inoise, nw, nb, y, wsq = ctx.saved_tensors
        lib = _lib.load()
        B, Cout, H, W = y.shape
        Cin, k = w.shape[1], w.shape[2]
        dev = y.device
        dy = dy if dy.is_contiguous(memory_format=torch.channels_last) else \
            dy.contiguous(memory_format=torch.channels_last)
        dz = torch.empty_like(y, memory_format=torch.channels_last)
        gd = torch.empty((B, Cout), dtype=torch.float32, device=dev) if d is not None else None
        gnw = torch.empty((Cout,), dtype=torch.float32, device=dev) if inoise is not None else None
        gnb = torch.empty((Cout,), dtype=torch.float32, device=dev) if inoise is not None else None
        with torch.cuda.device(dev):
            rc = lib.hg_modconv_epilogue_bwd(
                _lib.ptr(dy), _lib.ptr(y), _lib.ptr(d), _lib.ptr(inoise), _lib.ptr(nw), _lib.ptr(nb),
                _lib.ptr(dz), _lib.ptr(gd), _lib.ptr(gnw), _lib.ptr(gnb), B, H, W, Cout,
                int(inoise.shape[1]) if inoise is not None else 0, float(ctx.slope), _st(dev))
        _lib.check(rc, "hg_modconv_epilogue_bwd")
        dx = gmod = dw = None
        if ctx.needs_input_grad[0] or ctx.needs_input_grad[1]:
            dx = _conv.conv2d_nhwc(dz, ops._packs.get(w, 1), 1, k - 1 - (k - 1) // 2, cout=Cin)
            gmod = torch.empty((B, Cin), dtype=torch.float32, device=dev)
            if ctx.upsample:
                dxm = dx
                dx = torch.empty_like(x, memory_format=torch.channels_last)
                with torch.cuda.device(dev):
                    rc = lib.hg_upsample_modulate_bwd(_lib.ptr(dxm), _lib.ptr(x), _lib.ptr(mod),
                                                      _lib.ptr(dx), _lib.ptr(gmod), B, H // 2, W // 2,
                                                      Cin, _st(dev))
                _lib.check(rc, "hg_upsample_modulate_bwd")
            else:
                with torch.cuda.device(dev):
                    rc = lib.hg_modulate_bwd(_lib.ptr(dx), _lib.ptr(x), _lib.ptr(mod), _lib.ptr(gmod),
                                             B, H * W, Cin, _st(dev))
                _lib.check(rc, "hg_modulate_bwd")
        if ctx.needs_input_grad[2]:
            dw = _conv.conv2d_wgrad_nhwc(dz, xm, k, 1, (k - 1) // 2, out=ops.grad_slot(w))
        if wsq is not None and d is not None:
            analytic = dw is None or (dw.is_contiguous(memory_format=torch.channels_last) and
                                      w.is_contiguous(memory_format=torch.channels_last))
            assert analytic, "analytic demodulation needs channels_last weights"
            t_ws = torch.empty_like(gd)
            with torch.cuda.device(dev):
                rc = lib.hg_demod_bwd(_lib.ptr(gd), _lib.ptr(d), _lib.ptr(mod), _lib.ptr(wsq), _lib.ptr(w),
                                      _lib.ptr(gmod), _lib.ptr(dw), _lib.ptr(t_ws), B, Cout, k * k, Cin,
                                      _st(dev))
            _lib.check(rc, "hg_demod_bwd")
            gd = None
        return dx, gmod, dw, gd, None, gnw, gnb, None, None, None, None


def fusable(x, w):
    """can conv `w` (OIHW) run as a fused layer on activation x?  (channels multiples of 4)"""
    return x.is_cuda and w.shape[0] % 4 == 0 and w.shape[1] % 4 == 0


def mod_conv_layer(x, style, weight, demod, inoise, noise_lin, slope=0.2, eps=1e-8, upsample=False,
                   act=True):
    """LeakyReLU(Conv2DMod(up(x), style) + to_noise(inoise).permute(0,3,2,1)) in one fused op.
    style (B,Cin); inoise (B,S,S,1) image noise or None; noise_lin = the nn.Linear(1, Cout);
    upsample=True folds the block's 2x bilinear nn.Upsample of x into the op; act=False drops
    the LeakyReLU (a bare Conv2DMod.forward, histoGAN.py:420-440)."""
    mod = style + 1                                                    # histoGAN.py:423-425
    d = None
    if demod:                                                          # :427-429
        wsq = weight.pow(2).sum(dim=(2, 3))
        d = torch.rsqrt(mod.pow(2) @ wsq.t() + eps)
    nz = nw = nb = None
    if inoise is not None:
        nz = inoise.reshape(inoise.shape[0], inoise.shape[1], inoise.shape[2])
        nw = noise_lin.weight.reshape(-1)
        nb = noise_lin.bias
    return _ModConvLayer.apply(x, mod, weight, d, nz, nw, nb, slope, bool(upsample), bool(act))


class _ToRGB(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, wmod, prev):
        lib = _lib.load()
        x = x if x.is_contiguous(memory_format=torch.channels_last) else \
            x.contiguous(memory_format=torch.channels_last)
        B, Cc, H, W = x.shape
        wmod = wmod.contiguous()
        prev_c = prev.contiguous() if prev is not None else None
        rgb = torch.empty((B, 3, H, W), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            rc = lib.hg_torgb_fwd(_lib.ptr(x), _lib.ptr(wmod), _lib.ptr(prev_c), _lib.ptr(rgb), B,
                                  H * W, Cc, _st(x.device))
        _lib.check(rc, "hg_torgb_fwd")
        ctx.save_for_backward(x, wmod)
        ctx.has_prev = prev is not None
        return rgb

    @staticmethod
    @once_differentiable
    def backward(ctx, drgb):
        x, wmod = ctx.saved_tensors
        lib = _lib.load()
        B, Cc, H, W = x.shape
        drgb = drgb.contiguous()
        dx = torch.empty_like(x, memory_format=torch.channels_last)
        gw = torch.empty_like(wmod)
        with torch.cuda.device(x.device):
            rc = lib.hg_torgb_bwd(_lib.ptr(drgb), _lib.ptr(x), _lib.ptr(wmod), _lib.ptr(dx),
                                  _lib.ptr(gw), B, H * W, Cc, 0, _st(x.device))
        _lib.check(rc, "hg_torgb_bwd")
        return dx, gw, (drgb if ctx.has_prev else None)


def to_rgb(x, style, weight, prev_rgb):
    """Conv2DMod(C -> 3, k=1, demod=False)(x, style) + prev_rgb, planar NCHW result."""
    wmod = weight[None, :, :, 0, 0] * (style[:, None, :] + 1)           # (B,3,C)
    return _ToRGB.apply(x, wmod, prev_rgb)


class _Upsample2xPlanar(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        lib = _lib.load()
        x = x.contiguous()
        B, Cc, H, W = x.shape
        y = torch.empty((B, Cc, 2 * H, 2 * W), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            rc = lib.hg_upsample2x_planar(_lib.ptr(x), _lib.ptr(y), B * Cc, H, W, 0, _st(x.device))
        _lib.check(rc, "hg_upsample2x_planar")
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        lib = _lib.load()
        dy = dy.contiguous()
        B, Cc, OH, OW = dy.shape
        dx = torch.empty((B, Cc, OH // 2, OW // 2), dtype=torch.float32, device=dy.device)
        with torch.cuda.device(dy.device):
            rc = lib.hg_upsample2x_planar(_lib.ptr(dy), _lib.ptr(dx), B * Cc, OH // 2, OW // 2, 1,
                                          _st(dy.device))
        _lib.check(rc, "hg_upsample2x_planar")
        return dx


def upsample2x_planar(x):
    """nn.Upsample(scale_factor=2, mode='bilinear', align_corners=False) of a planar (NCHW)
    float32 CUDA tensor -- RGBBlock's skip path (histoGAN.py:377-378,388-389)."""
    return _Upsample2xPlanar.apply(x)


# ------------------------------------------------------------ style path (style.cu) ------
LIN_ADD_ONE, LIN_LRELU, LIN_SQUARE_INPUT, LIN_RSQRT_EPS, LIN_POST_2X, LIN_ACCUMULATE = 1, 2, 4, 8, 16, 32


def _ptr_table(ts):
    return (C.c_void_p * len(ts))(*[t.data_ptr() if t is not None else None for t in ts])


def _int_table(vs):
    return (C.c_int32 * len(vs))(*[int(v) for v in vs])


def grouped_linear(xs, ws, bs, flags=0, slope=0.2, eps=1e-8):
    """raw hg_grouped_linear_fwd: y_g = f(x_g W_g^T + b_g) for every group in ONE launch.
    xs[g] (B,K_g), ws[g] (J_g,K_g), bs[g] (J_g) or None -> list of (B,J_g)."""
    lib = _lib.load()
    B, dev = xs[0].shape[0], xs[0].device
    ys = [torch.empty((B, w.shape[0]), dtype=torch.float32, device=dev) for w in ws]
    with torch.cuda.device(dev):
        rc = lib.hg_grouped_linear_fwd(len(xs), _ptr_table(xs), _ptr_table(ws), _ptr_table(bs), _ptr_table(ys),
                                       _int_table([w.shape[0] for w in ws]), _int_table([w.shape[1] for w in ws]),
                                       B, flags, float(slope), float(eps), _st(dev))
    _lib.check(rc, "hg_grouped_linear_fwd")
    return ys


def grouped_linear_bwd(xs, ws, gys, gws, gbs, gxs, flags=0):
    """raw hg_grouped_linear_bwd into the given output tensors (entries may be None)"""
    lib = _lib.load()
    B, dev = xs[0].shape[0], xs[0].device
    with torch.cuda.device(dev):
        rc = lib.hg_grouped_linear_bwd(len(xs), _ptr_table(xs), _ptr_table(ws), _ptr_table(gys),
                                       _ptr_table(gws), _ptr_table(gbs), _ptr_table(gxs),
                                       _int_table([w.shape[0] for w in ws]), _int_table([w.shape[1] for w in ws]),
                                       B, flags, _st(dev))
    _lib.check(rc, "hg_grouped_linear_bwd")


MAX_GROUPS = 24          # style.cu kMaxGroups (a multiple of 3: a chunk holds whole blocks)


class _GroupedStyleLinear(torch.autograd.Function):
    """mods = [to_style_g(istyle[block_g]) + 1 for g]: the 3 Linears (to_style1, to_style2,
    to_rgb.to_style; histoGAN.py:451,455,372) of EVERY generator block and the `y + 1` of
    Conv2DMod (:423-425) in one launch; backward = one weight/bias-gradient launch + one
    input-gradient launch."""

    @staticmethod
    def forward(ctx, per_block, block_of, *wb):
        ws, bs = list(wb[0::2]), list(wb[1::2])
        xs = [per_block[i] for i in block_of]
        ctx.save_for_backward(per_block, *ws)
        ctx.block_of = block_of
        return tuple(grouped_linear(xs, ws, bs, LIN_ADD_ONE))

    @staticmethod
    @once_differentiable
    def backward(ctx, *gys):
        per_block, *ws = ctx.saved_tensors
        block_of = ctx.block_of
        xs = [per_block[i] for i in block_of]
        B = per_block.shape[1]
        gys = [g.contiguous() if g is not None else torch.zeros((B, w.shape[0]), device=w.device)
               for g, w in zip(gys, ws)]
        gws = [torch.empty_like(w) for w in ws]
        gbs = [torch.empty((w.shape[0],), dtype=torch.float32, device=w.device) for w in ws]
        gx = torch.empty((len(ws),) + tuple(per_block.shape[1:]), dtype=torch.float32, device=per_block.device)
        grouped_linear_bwd(xs, ws, gys, gws, gbs, list(gx.unbind(0)), 0)
        # the groups come as consecutive triples per block (style_mods): sum each triple, zero-pad
        # the blocks this call did not touch (no index tensors: this runs under graph capture)
        n = len(block_of) // 3
        assert block_of == tuple(block_of[0] + i // 3 for i in range(3 * n)), block_of
        g = gx.view(n, 3, *per_block.shape[1:]).sum(dim=1)
        if n != per_block.shape[0]:
            lo, hi = block_of[0], per_block.shape[0] - block_of[0] - n
            g = torch.cat([g.new_zeros((lo,) + tuple(g.shape[1:])), g, g.new_zeros((hi,) + tuple(g.shape[1:]))])
        out = [g, None]
        for gw, gb in zip(gws, gbs):
            out += [gw, gb]
        return tuple(out)


def style_mods(per_block, linears):
    """per_block (L,B,latent); linears = [(block index, nn.Linear)]; returns [Linear(istyle)+1]."""
    per_block = per_block.contiguous().float()
    mods = []
    for i in range(0, len(linears), MAX_GROUPS):
        chunk = linears[i:i + MAX_GROUPS]
        wb = []
        for _, lin in chunk:
            wb += [lin.weight, lin.bias]
        mods += list(_GroupedStyleLinear.apply(per_block, tuple(b for b, _ in chunk), *wb))
    return mods


def weight_sqsum(w, out=None):
    """Wsq (Cout,Cin) = sum_taps w^2 of a channels_last (Cout,Cin,k,k) weight (hg_weight_sqsum)"""
    lib = _lib.load()
    co, ci, kh, kw = w.shape
    assert w.is_contiguous(memory_format=torch.channels_last) and ci % 4 == 0
    if out is None:
        out = torch.empty((co, ci), dtype=torch.float32, device=w.device)
    with torch.cuda.device(w.device):
        rc = lib.hg_weight_sqsum(_lib.ptr(w.detach()), _lib.ptr(out), co, kh * kw, ci, _st(w.device))
    _lib.check(rc, "hg_weight_sqsum")
    return out


@torch.no_grad()
def demod_all(mods, wsqs, eps=1e-8):
    """d_g = rsqrt(mod_g^2 Wsq_g^T + eps) (histoGAN.py:427-429) for every conv layer of a generator
    pass in one launch.  NOT differentiated by autograd: _ModConvLayer's backward carries the
    dependence on (mod, w) analytically."""
    out = []
    for i in range(0, len(mods), MAX_GROUPS):
        out += grouped_linear([m.detach() for m in mods[i:i + MAX_GROUPS]], wsqs[i:i + MAX_GROUPS],
                              [None] * len(mods[i:i + MAX_GROUPS]), LIN_SQUARE_INPUT | LIN_RSQRT_EPS, eps=eps)
    return out


def mod_conv_layer_pre(x, mod, weight, d, wsq, inoise, noise_lin, slope=0.2, upsample=False, act=True):
    """mod_conv_layer with the modulation (style + 1), the demodulation factor d and Wsq already
    computed by style_mods / demod_all (one launch each for the whole generator)."""
    nz = nw = nb = None
    if inoise is not None:
        nz = inoise.reshape(inoise.shape[0], inoise.shape[1], inoise.shape[2])
        nw = noise_lin.weight.reshape(-1)
        nb = noise_lin.bias
    return _ModConvLayer.apply(x, mod, weight, d, nz, nw, nb, slope, bool(upsample), bool(act), wsq)


def to_rgb_mod(x, mod, weight, prev_rgb):
    """to_rgb with the modulation (style + 1) already formed"""
    wmod = weight[None, :, :, 0, 0] * mod[:, None, :]                    # (B,3,C)
    return _ToRGB.apply(x, wmod, prev_rgb)


class _LinearLReLU(torch.autograd.Function):
    """LeakyReLU(Linear(x)) of a skinny batch (<= 32 rows): StyleVectorizer / HistVectorizer layers
    (histoGAN.py:335-365) on the grouped-linear kernels (weights read once, batch on the lanes)."""

    @staticmethod
    def forward(ctx, x, w, b, slope):
        x = x.contiguous().float()
        (y,) = grouped_linear([x], [w], [b], LIN_LRELU, slope=slope)
        ctx.save_for_backward(x, w, y)
        ctx.slope = slope
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, gy):
        x, w, y = ctx.saved_tensors
        gpre = torch.where(y > 0, gy, gy * ctx.slope).contiguous()
        gw = torch.empty_like(w) if ctx.needs_input_grad[1] else None
        gb = torch.empty((w.shape[0],), dtype=torch.float32, device=w.device) if ctx.needs_input_grad[2] else None
        gx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        grouped_linear_bwd([x], [w], [gpre], [gw], [gb], [gx], 0)
        return gx, gw, gb, None


def mlp_lrelu(x, linears, slope=0.2):
    """[Linear -> LeakyReLU(slope)] x n on the grouped-linear kernels; x (B <= 32, K), K % 4 == 0"""
    for lin in linears:
        x = _LinearLReLU.apply(x, lin.weight, lin.bias, slope)
    return x


def mlp_ok(x, linears):
    return (x.is_cuda and x.dim() == 2 and x.shape[0] <= 32 and x.dtype == torch.float32
            and all(l.weight.shape[1] % 4 == 0 and l.weight.is_contiguous() for l in linears))
