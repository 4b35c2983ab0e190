"""GPU: the fused generator-layer operators against the same math written with plain
PyTorch ops (fp32; conv operands TF32-rounded identically on both sides)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def _torch_layer(x, style, w, inoise, lin, upsample):
    from histogan_b200 import conv
    if upsample:
        x = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False)
    mod = style + 1
    xm = conv.tf32_round(x * mod[:, :, None, None])
    # straight-through rounding so autograd sees x * mod
    xm = x * mod[:, :, None, None] + (xm - x * mod[:, :, None, None]).detach()
    z = F.conv2d(xm.double(), conv.tf32_round(w).double().detach() + (w - w.detach()).double(), padding=1).float()
    d = torch.rsqrt(mod.pow(2) @ w.pow(2).sum(dim=(2, 3)).t() + 1e-8)
    nz = inoise[:, :x.shape[2], :x.shape[3], :]
    noise = lin(nz).permute(0, 3, 2, 1)
    return F.leaky_relu(z * d[:, :, None, None] + noise, 0.2)


@pytest.mark.parametrize("upsample", [False, True])
def test_mod_conv_layer_fwd_bwd(upsample, cuda_device):
    from histogan_b200 import fused
    torch.manual_seed(0)
    B, Cin, Cout, S = 3, 64, 96, 8
    x = torch.randn(B, Cin, S, S, device="cuda").contiguous(memory_format=torch.channels_last).requires_grad_(True)
    style = torch.randn(B, Cin, device="cuda", requires_grad=True)
    w = (torch.randn(Cout, Cin, 3, 3, device="cuda") / 24).requires_grad_(True)
    lin = torch.nn.Linear(1, Cout).cuda()
    inoise = torch.rand(B, 32, 32, 1, device="cuda")
    So = S * (2 if upsample else 1)
    gy = torch.randn(B, Cout, So, So, device="cuda")

    y = fused.mod_conv_layer(x, style, w, True, inoise, lin, upsample=upsample)
    grads = torch.autograd.grad(y, [x, style, w, lin.weight, lin.bias], gy)
    yr = _torch_layer(x, style, w, inoise, lin, upsample)
    gref = torch.autograd.grad(yr, [x, style, w, lin.weight, lin.bias], gy)
    assert _rel(y, yr) < 2e-5
    # backward: dz is TF32-rounded before dgrad/wgrad in the fused op, not in the torch graph
    for name, a, b in zip(("dx", "dstyle", "dw", "dnoise_w", "dnoise_b"), grads, gref):
        assert _rel(a, b) < 2e-3, (name, _rel(a, b))


@pytest.mark.parametrize("shape", [(3, 64, 16), (2, 32, 64), (2, 2048, 4), (3, 1024, 8), (2, 512, 16),
                                   (2, 256, 32), (1, 128, 5), (2, 48, 7), (33, 32, 128), (40, 256, 16), (2, 8, 6)])
def test_to_rgb_fwd_bwd(shape, cuda_device):
    """every lane-group width of torgb_fwd (8..256 threads per pixel, 1 and 4 pixels per thread), odd
    sizes, a channel count that is not a power of two"""
    from histogan_b200 import fused
    torch.manual_seed(1)
    B, Cc, S = shape
    x = torch.randn(B, Cc, S, S, device="cuda").contiguous(memory_format=torch.channels_last).requires_grad_(True)
    style = torch.randn(B, Cc, device="cuda", requires_grad=True)
    w = torch.randn(3, Cc, 1, 1, device="cuda", requires_grad=True)
    prev = torch.randn(B, 3, S, S, device="cuda", requires_grad=True)
    g = torch.randn(B, 3, S, S, device="cuda")
    y = fused.to_rgb(x, style, w, prev)
    grads = torch.autograd.grad(y, [x, style, w, prev], g)
    wm = w[None, :, :, 0, 0] * (style[:, None, :] + 1)
    yr = torch.einsum("bchw,boc->bohw", x, wm) + prev
    gref = torch.autograd.grad(yr, [x, style, w, prev], g)
    assert _rel(y, yr) < 1e-5
    for name, a, b in zip(("dx", "dstyle", "dw", "dprev"), grads, gref):
        assert _rel(a, b) < 1e-5, (name, _rel(a, b))


def test_conv_bias_act_and_double_backward(cuda_device):
    """fused D layer (conv + bias + lrelu [+ residual]) incl. the second-order path."""
    from histogan_b200 import conv, ops
    torch.manual_seed(2)
    B, Cin, Cout, S = 2, 32, 64, 16
    x = conv.tf32_round(torch.randn(B, Cin, S, S)).cuda().contiguous(memory_format=torch.channels_last).requires_grad_(True)
    w = conv.tf32_round(torch.randn(Cout, Cin, 3, 3) / 17).cuda().requires_grad_(True)
    b = torch.randn(Cout, device="cuda", requires_grad=True)
    res = torch.randn(B, Cout, S, S, device="cuda").contiguous(memory_format=torch.channels_last)

    def ours():
        t = ops.conv_bias_act(x, w, b, None, 1, 1, act=True, x_rounded=True)
        return ops.conv_bias_act(x, w, b, t, 1, 1, act=False, x_rounded=True)

    def ref():
        t = F.leaky_relu(F.conv2d(x, w, b, padding=1), 0.2)
        return F.conv2d(x, w, b, padding=1) + t

    torch.backends.cudnn.allow_tf32 = False
    for fn, tol in ((ours, None),):
        y, yr = ours(), ref()
        assert _rel(y, yr) < 2e-5
        # first order
        g = torch.randn_like(yr)
        ga = torch.autograd.grad(y, [x, w, b], g, create_graph=True)
        gb = torch.autograd.grad(yr, [x, w, b], g, create_graph=True)
        for a_, b_ in zip(ga, gb):
            assert _rel(a_, b_) < 2e-3
        # second order: d/dw of || d y / d x ||^2   (the gradient-penalty pattern)
        (ha,) = torch.autograd.grad(ga[0].pow(2).sum(), w)
        (hb,) = torch.autograd.grad(gb[0].pow(2).sum(), w)
        assert _rel(ha, hb) < 5e-3, _rel(ha, hb)


def test_upsample2x_planar_matches_torch(cuda_device):
    """RGBBlock's skip up-sampling (histoGAN.py:377-378): hand-written planar kernel == nn.Upsample,
    forward and adjoint, incl. odd / non-square / 1-pixel extents."""
    import torch.nn.functional as F
    from histogan_b200 import fused
    torch.manual_seed(0)
    for shape in [(2, 3, 4, 4), (3, 3, 16, 16), (1, 3, 7, 5), (2, 4, 1, 1), (2, 3, 128, 128)]:
        x = torch.randn(shape, device="cuda", requires_grad=True)
        y = fused.upsample2x_planar(x)
        ref = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False)
        assert y.shape == ref.shape
        assert (y - ref).abs().max().item() <= 1e-6 * ref.abs().max().item() + 1e-7
        g = torch.randn_like(ref)
        (gx,) = torch.autograd.grad(y, x, g)
        (gr,) = torch.autograd.grad(ref, x, g)
        assert (gx - gr).abs().max().item() <= 1e-5 * gr.abs().max().item() + 1e-7, shape


# ------------------------------------------------------------- style path (style.cu) -----

def test_grouped_linear_forward_and_backward(cuda_device):
    """hg_grouped_linear_fwd/bwd against torch: ragged groups, K over several shared-memory chunks,
    J not a multiple of the CTA row count, every epilogue flag."""
    import torch.nn.functional as F
    from histogan_b200 import fused as fz
    torch.manual_seed(0)
    for B in (32, 5):
        shapes = [(64, 512), (2048, 512), (100, 512), (3, 512), (130, 2048), (17, 36)]
        xs = [torch.randn(B, k, device="cuda") for _, k in shapes]
        ws = [torch.randn(j, k, device="cuda") / k ** 0.5 for j, k in shapes]
        bs = [torch.randn(j, device="cuda") if i % 2 == 0 else None for i, (j, _) in enumerate(shapes)]
        ys = fz.grouped_linear(xs, ws, bs, fz.LIN_ADD_ONE)
        for x, w, b, y in zip(xs, ws, bs, ys):
            ref = F.linear(x.double(), w.double(), b.double() if b is not None else None) + 1
            assert (y.double() - ref).abs().max().item() < 2e-5, (B, tuple(w.shape))
        ys = fz.grouped_linear(xs, ws, bs, fz.LIN_LRELU, slope=0.2)
        for x, w, b, y in zip(xs, ws, bs, ys):
            ref = F.leaky_relu(F.linear(x.double(), w.double(), b.double() if b is not None else None), 0.2)
            assert (y.double() - ref).abs().max().item() < 2e-5
        wp = [w.abs() for w in ws]
        ys = fz.grouped_linear(xs, wp, [None] * len(xs), fz.LIN_SQUARE_INPUT | fz.LIN_RSQRT_EPS, eps=1e-8)
        for x, w, y in zip(xs, wp, ys):
            ref = torch.rsqrt(x.double().pow(2) @ w.double().t() + 1e-8)
            assert ((y.double() - ref).abs() / ref).max().item() < 2e-5
        # backward: gW, gb, gx (+ accumulate, + 2x post-factor)
        gys = [torch.randn(B, j, device="cuda") for j, _ in shapes]
        gws = [torch.empty_like(w) for w in ws]
        gbs = [torch.empty(j, device="cuda") for j, _ in shapes]
        gxs = [torch.empty_like(x) for x in xs]
        fz.grouped_linear_bwd(xs, ws, gys, gws, gbs, gxs, 0)
        for x, w, gy, gw, gb, gx in zip(xs, ws, gys, gws, gbs, gxs):
            assert (gw.double() - gy.double().t() @ x.double()).abs().max().item() < 1e-4
            assert (gb.double() - gy.double().sum(0)).abs().max().item() < 1e-4
            assert (gx.double() - gy.double() @ w.double()).abs().max().item() < 1e-4
        acc = [torch.randn_like(x) for x in xs]
        acc0 = [a.clone() for a in acc]
        fz.grouped_linear_bwd(xs, ws, gys, [None] * 6, [None] * 6, acc, fz.LIN_POST_2X | fz.LIN_ACCUMULATE)
        for x, w, gy, a, a0 in zip(xs, ws, gys, acc, acc0):
            ref = a0.double() + 2 * x.double() * (gy.double() @ w.double())
            assert (a.double() - ref).abs().max().item() < 2e-4


def test_grouped_linear_long_k_split(cuda_device):
    """one layer with a long K (the 12288-wide first layer of the histogram MLP, histoGAN.py:384-398):
    the K chunks are divided over several CTAs per row block, ordered finish with the epilogue;
    deterministic."""
    import torch.nn.functional as F
    from histogan_b200 import fused as fz
    torch.manual_seed(1)
    for B, J, K in [(32, 1024, 12288), (7, 100, 4100), (32, 512, 4096)]:
        x = torch.randn(B, K, device="cuda")
        w = torch.randn(J, K, device="cuda") / K ** 0.5
        b = torch.randn(J, device="cuda")
        y, = fz.grouped_linear([x], [w], [b], fz.LIN_LRELU, slope=0.2)
        ref = F.leaky_relu(F.linear(x.double(), w.double(), b.double()), 0.2)
        assert (y.double() - ref).abs().max().item() < 2e-5, (B, J, K)
        y2, = fz.grouped_linear([x], [w], [b], fz.LIN_LRELU, slope=0.2)
        assert torch.equal(y, y2)


def test_analytic_demodulation_matches_autograd(cuda_device):
    """_ModConvLayer with the demodulation differentiated analytically (hg_demod_bwd: style_mods /
    demod_all path of Generator.forward) == the same layer with torch ops + autograd for
    d = rsqrt((style+1)^2 Wsq^T + eps) (histoGAN.py:423-429)."""
    from histogan_b200 import fused as fz, ops
    torch.manual_seed(0)
    for B, cin, cout, s, up in [(3, 64, 128, 8, False), (2, 32, 32, 16, True), (4, 128, 64, 4, False)]:
        x = torch.randn(B, cin, s, s, device="cuda").contiguous(memory_format=torch.channels_last).requires_grad_(True)
        style = (torch.randn(B, cin, device="cuda") * 0.5).requires_grad_(True)
        w = torch.nn.Parameter((torch.randn(cout, cin, 3, 3, device="cuda") / (cin * 9) ** 0.5)
                               .contiguous(memory_format=torch.channels_last))
        lin = torch.nn.Linear(1, cout).cuda()
        so = s * (2 if up else 1)
        nz = torch.rand(B, so, so, 1, device="cuda")
        gy = torch.randn(B, cout, so, so, device="cuda")

        y0 = fz.mod_conv_layer(x, style, w, True, nz, lin, upsample=up)
        g0 = torch.autograd.grad(y0, (x, style, w, lin.weight, lin.bias), gy)

        mod = style + 1
        wsq = ops._packs.get(w, 'wsq')
        assert (wsq - w.detach().pow(2).sum(dim=(2, 3))).abs().max().item() < 1e-6
        (d,) = fz.demod_all([mod], [wsq])
        y1 = fz.mod_conv_layer_pre(x, mod, w, d, wsq, nz, lin, upsample=up)
        g1 = torch.autograd.grad(y1, (x, style, w, lin.weight, lin.bias), gy)
        assert (y1 - y0).abs().max().item() <= 2e-5 * y0.abs().max().item()
        for name, a, b in zip(("x", "style", "w", "noise_w", "noise_b"), g1, g0):
            err = ((a - b).norm() / b.norm()).item()
            assert err < 2e-4, (name, err, (B, cin, cout, s, up))


def test_style_and_hist_vectorizers_match_torch(cuda_device, monkeypatch):
    """StyleVectorizer / HistVectorizer (histoGAN.py:335-365) on the grouped-linear kernels ==
    the nn.Sequential of nn.Linear + LeakyReLU they wrap, values and every gradient."""
    from histogan_b200 import gan
    torch.manual_seed(0)
    for mod, x in ((gan.StyleVectorizer(512, 8).cuda(), torch.randn(32, 512, device="cuda")),
                   (gan.HistVectorizer(64, 512, 8).cuda(), torch.rand(7, 3, 64, 64, device="cuda"))):
        x.requires_grad_(True)
        y = mod(x)
        g = torch.randn_like(y)
        grads = torch.autograd.grad(y, [x] + list(mod.parameters()), g)
        monkeypatch.setattr(gan, "USE_FUSED", False)
        y0 = mod(x)
        grads0 = torch.autograd.grad(y0, [x] + list(mod.parameters()), g)
        monkeypatch.setattr(gan, "USE_FUSED", True)
        assert ((y - y0).norm() / y0.norm()).item() < 1e-5
        for a, b in zip(grads, grads0):
            assert ((a - b).norm() / b.norm().clamp_min(1e-20)).item() < 1e-4
