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


def test_to_rgb_fwd_bwd(cuda_device):
    from histogan_b200 import fused
    torch.manual_seed(1)
    B, Cc, S = 3, 64, 16
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
