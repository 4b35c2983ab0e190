"""GPU: the tcgen05 implicit-GEMM convolution against a plain PyTorch fp32
reference of the same op (inputs pre-rounded to TF32, so the only difference
left is fp32 accumulation order: tolerance 2e-5 relative to the output scale)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

CASES = [
    # B, Cin, H, W, Cout, k, stride
    (8, 64, 4, 4, 128, 3, 1),      # generator block 0 shape family (4x4, several images per tile)
    (4, 128, 8, 8, 64, 3, 1),      # 8x8: two images per tile
    (2, 64, 16, 16, 64, 3, 1),     # one image row-band per tile
    (2, 32, 64, 64, 32, 3, 1),     # BLOCK_N = 32
    (3, 96, 32, 32, 192, 3, 1),    # BLOCK_N = 64, odd batch, K = 9*3 k-blocks
    (2, 64, 32, 32, 128, 1, 1),    # 1x1 == plain GEMM
    (2, 64, 32, 32, 64, 3, 2),     # stride 2 (DiscriminatorBlock.downsample)
    (5, 256, 4, 4, 256, 3, 1),     # batch tail: 5 images, TB = 8
    (1, 32, 24, 40, 32, 3, 1),     # non power-of-two spatial size
    (2, 16, 32, 32, 16, 3, 1),     # 16-channel D block: 32-wide boxes completed by TMA zero fill
    (2, 4, 32, 32, 16, 3, 1),      # image padded 3 -> 4 channels
    (2, 16, 32, 32, 32, 1, 1),
    (2, 32, 32, 32, 16, 3, 2),
    (2, 48, 16, 16, 80, 3, 1),     # channel tails inside a k-block / an n-tile
    (8, 32, 128, 128, 32, 3, 1),   # >= 592 tiles, one n-tile, 36 KB filter: resident-filter halo variant
    (4, 64, 128, 160, 32, 3, 1),   # resident filter, two k-chunks per tap
    (4, 32, 128, 160, 64, 3, 1),   # resident filter, BLOCK_N = 64
    (5, 20, 120, 136, 24, 3, 1),   # resident filter with channel tails and ragged tiles
    (5, 32, 128, 128, 32, 3, 1),   # single-box halo (8x16 tiles, all nine taps from one TMA box)
    (3, 64, 160, 152, 32, 3, 1),   # ... two k-chunks per tap, ragged tiles in both directions
    (6, 32, 96, 104, 64, 3, 1),    # ... BLOCK_N = 64
    (32, 16, 64, 64, 16, 3, 1),    # ... channel tails (D block 0 shape family)
]


def _ref(x, w, stride, pad):
    return F.conv2d(x.double(), w.double(), stride=stride, padding=pad).float()


@pytest.mark.parametrize("case", CASES, ids=[str(c) for c in CASES])
def test_conv_matches_torch(case, cuda_device):
    from histogan_b200 import conv
    B, Cin, H, W, Cout, k, stride = case
    g = torch.Generator().manual_seed(0)
    x = conv.tf32_round(torch.randn(B, Cin, H, W, generator=g)).cuda()
    w = conv.tf32_round(torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5).cuda()
    pad = k // 2
    wp = conv.pack_weight(w, 0)
    y = conv.conv2d_nhwc(x, wp, stride, pad, cout=Cout)
    ref = _ref(x, w, stride, pad)
    assert y.shape == ref.shape
    err = (y - ref).abs().max().item() / ref.abs().max().item()
    assert err < 2e-5, err


def test_tf32_rounding_keeps_nan_and_inf(cuda_device):
    """hg::tf32_round and its torch twin: round-to-nearest-even on finite values; Inf / NaN -- including the
    hardware's canonical NaN 0x7FFFFFFF, which a bare integer add carries over into -0.0 -- pass through, so
    a NaN weight still reaches the loss (Trainer's NaN recovery, histoGAN.py:1003-1006)."""
    from histogan_b200 import conv, ops
    bits = torch.tensor([0x7FFFFFFF, 0x7FC00000, -1, 0x7F800000, -8388608, 0x3F801000, 0x3F803000, 0x3F800FFF],
                        dtype=torch.int32)
    x = bits.view(torch.float32).reshape(1, 8, 1, 1).repeat(2, 1, 3, 3).cuda()
    for y in (ops.round_tf32_nhwc(x), conv.tf32_round(x)):
        y = y.cpu()[0, :, 0, 0]
        assert torch.isnan(y[:3]).all() and y[3] == float("inf") and y[4] == -float("inf")
        assert y[5:8].contiguous().view(torch.int32).tolist() == [0x3F800000, 0x3F804000, 0x3F800000]   # ties to even
    # a NaN in one input element poisons its receptive field and nothing else
    w = torch.randn(32, 32, 3, 3).cuda()
    xx = conv.tf32_round(torch.randn(1, 32, 8, 8)).cuda().contiguous(memory_format=torch.channels_last)
    xx[0, 5, 4, 4] = float("nan")
    y = conv.conv2d_nhwc(xx, conv.pack_weight(w, 0), 1, 1, cout=32, round_tf32=True)
    nan = torch.isnan(y[0, 0])
    assert nan[3:6, 3:6].all() and int(nan.sum()) == 9


def test_conv_epilogue(cuda_device):
    from histogan_b200 import conv
    B, Cin, S, Cout = 2, 64, 16, 64
    g = torch.Generator().manual_seed(1)
    x = conv.tf32_round(torch.randn(B, Cin, S, S, generator=g)).cuda()
    w = conv.tf32_round(torch.randn(Cout, Cin, 3, 3, generator=g) / 24).cuda()
    scale = torch.rand(B, Cout, generator=g).cuda() + 0.5
    bias = torch.randn(Cout, generator=g).cuda()
    noise = torch.rand(B, 32, 32, generator=g).cuda()          # image noise larger than the layer
    nw, nb = torch.randn(Cout, generator=g).cuda(), torch.randn(Cout, generator=g).cuda()
    res = torch.randn(B, Cout, S, S, generator=g).cuda()
    wp = conv.pack_weight(w, 0)
    y = conv.conv2d_nhwc(x, wp, 1, 1, scale=scale, bias=bias, noise=noise, noise_w=nw, noise_b=nb,
                         residual=res, lrelu=True)
    ref = _ref(x, w, 1, 1) * scale[:, :, None, None] + bias[None, :, None, None]
    nz = noise[:, :S, :S].transpose(1, 2)                       # (b, oh, ow) <- noise[b, ow, oh]
    ref = ref + nz[:, None] * nw[None, :, None, None] + nb[None, :, None, None]
    ref = F.leaky_relu(ref, 0.2) + res
    err = (y - ref).abs().max().item() / ref.abs().max().item()
    assert err < 2e-5, err
    y2 = conv.conv2d_nhwc(x, wp, 1, 1, round_tf32=True)
    assert torch.equal(y2, conv.tf32_round(conv.conv2d_nhwc(x, wp, 1, 1)))


def test_conv_epilogue_split_k(cuda_device):
    """few output tiles + a long K loop -> split-K: partial sums by atomics into a zeroed y, the
    fused epilogue chain applied by the finishing kernel."""
    from histogan_b200 import conv
    B, Cin, S, Cout = 4, 256, 4, 128
    g = torch.Generator().manual_seed(5)
    x = conv.tf32_round(torch.randn(B, Cin, S, S, generator=g)).cuda()
    w = conv.tf32_round(torch.randn(Cout, Cin, 3, 3, generator=g) / 48).cuda()
    scale = torch.rand(B, Cout, generator=g).cuda() + 0.5
    bias = torch.randn(Cout, generator=g).cuda()
    noise = torch.rand(B, 8, 8, generator=g).cuda()
    nw, nb = torch.randn(Cout, generator=g).cuda(), torch.randn(Cout, generator=g).cuda()
    res = torch.randn(B, Cout, S, S, generator=g).cuda()
    wp = conv.pack_weight(w, 0)
    y = conv.conv2d_nhwc(x, wp, 1, 1, scale=scale, bias=bias, noise=noise, noise_w=nw, noise_b=nb,
                         residual=res, lrelu=True)
    ref = _ref(x, w, 1, 1) * scale[:, :, None, None] + bias[None, :, None, None]
    nz = noise[:, :S, :S].transpose(1, 2)
    ref = ref + nz[:, None] * nw[None, :, None, None] + nb[None, :, None, None]
    ref = F.leaky_relu(ref, 0.2) + res
    err = (y - ref).abs().max().item() / ref.abs().max().item()
    assert err < 2e-5, err
    y2 = conv.conv2d_nhwc(x, wp, 1, 1, bias=bias, round_tf32=True)
    assert bool(((y2.view(torch.int32) & 0x1FFF) == 0).all())
    plain = conv.conv2d_nhwc(x, wp, 1, 1)
    assert (plain - _ref(x, w, 1, 1)).abs().max().item() / ref.abs().max().item() < 2e-5


def test_dgrad_weight_packing(cuda_device):
    """conv(dy, pack(w, mode=1)) == d/dx of conv(x, w) for stride 1."""
    from histogan_b200 import conv
    g = torch.Generator().manual_seed(2)
    x = torch.randn(2, 64, 16, 16, generator=g).cuda().requires_grad_(True)
    w = conv.tf32_round(torch.randn(96, 64, 3, 3, generator=g) / 24).cuda()
    dy = conv.tf32_round(torch.randn(2, 96, 16, 16, generator=g)).cuda()
    (ref,) = torch.autograd.grad(F.conv2d(x.double(), w.double(), padding=1), x, dy.double())
    dx = conv.conv2d_nhwc(dy, conv.pack_weight(w, 1), 1, 1, cout=64)
    err = (dx - ref.float()).abs().max().item() / ref.abs().max().item()
    assert err < 2e-5, err


WG_CASES = [
    # B, Cin, H, W, Cout, k, stride
    (8, 64, 4, 4, 128, 3, 1),
    (4, 128, 8, 8, 64, 3, 1),
    (2, 64, 16, 16, 96, 3, 1),
    (2, 32, 64, 64, 32, 3, 1),     # many pixel blocks -> split-K + atomics
    (3, 96, 32, 32, 160, 3, 1),    # Cout tail inside the 128-wide M tile
    (2, 64, 32, 32, 128, 1, 1),
    (2, 64, 32, 32, 64, 3, 2),
    (1, 32, 24, 40, 32, 3, 1),
    (2, 16, 32, 32, 16, 3, 1),
    (2, 4, 32, 32, 16, 3, 1),
    (2, 16, 32, 32, 32, 3, 2),
    (2, 48, 16, 16, 80, 3, 1),
    (2, 16, 64, 64, 16, 3, 1),     # all-taps variant (<= 32 input channels, long pixel reduction)
    (4, 32, 64, 64, 64, 3, 2),     # all-taps, stride 2
    (2, 32, 64, 64, 160, 3, 1),    # all-taps, two co tiles with a tail
    (4, 32, 128, 128, 32, 3, 1),   # column-halo variant (few output channels, >= 64K pixels)
    (2, 64, 190, 176, 48, 3, 1),   # column-halo: two ci tiles, co tail, ragged tile rows / columns
    (5, 20, 120, 112, 64, 3, 1),   # column-halo: channel tail in ci, two dy chunks
    (3, 128, 160, 144, 16, 3, 1),  # column-halo: four ci tiles, 16 output channels
    (8, 64, 96, 96, 128, 3, 1),    # column-halo: four dy chunks
    (8, 32, 96, 96, 96, 3, 1),     # 96 output channels: not a column-halo shape (generic path)
]


@pytest.mark.parametrize("case", WG_CASES, ids=[str(c) for c in WG_CASES])
def test_wgrad_matches_torch(case, cuda_device):
    from histogan_b200 import conv
    B, Cin, H, W, Cout, k, stride = case
    g = torch.Generator().manual_seed(3)
    pad = k // 2
    x = conv.tf32_round(torch.randn(B, Cin, H, W, generator=g)).cuda()
    w = torch.zeros(Cout, Cin, k, k, device="cuda", dtype=torch.float64, requires_grad=True)
    y = F.conv2d(x.double(), w, stride=stride, padding=pad)
    dy = conv.tf32_round(torch.randn(y.shape, generator=g)).cuda()
    (ref,) = torch.autograd.grad(y, w, dy.double())
    dw = conv.conv2d_wgrad_nhwc(dy, x, k, stride, pad)
    err = (dw - ref.float()).abs().max().item() / ref.abs().max().item()
    assert err < 2e-5, err


SMALL_CASES = [
    # B, Cin, H, W, Cout, k, Cp
    (2, 3, 32, 32, 16, 3, 32),     # DiscriminatorBlock 0 net[0], channel-padded output
    (2, 3, 32, 32, 16, 1, 32),     # ... conv_res
    (3, 3, 19, 23, 24, 3, 24),     # odd sizes, Cout not a power of two, dense output
    (1, 4, 16, 16, 64, 3, 64),     # transparent images (4 channels), widest supported layer
    (5, 3, 64, 64, 16, 1, 16),
    (2, 3, 48, 80, 16, 3, 16),     # 3 -> 16, dense 16-channel output: the constant-memory kernels (k = 3)
    (3, 3, 19, 23, 16, 3, 16),     # ... odd sizes (partial 32 x 8 dgrad tiles)
    (2, 3, 21, 37, 16, 1, 16),     # ... k = 1
]


@pytest.mark.parametrize("case", SMALL_CASES, ids=[str(c) for c in SMALL_CASES])
def test_image_input_conv_kernels_match_torch(case, cuda_device):
    """conv_small.cu (the discriminator's image-input convolutions on the CUDA cores, exact fp32):
    forward with the fused epilogue, input gradient, weight gradient -- against torch float64, on a
    NON-contiguous planar image (the kernels take element strides)."""
    from histogan_b200 import conv
    B, Cin, H, W, Cout, k, Cp = case
    g = torch.Generator().manual_seed(3)
    big = torch.randn(B, Cin + 1, H, W + 3, generator=g).cuda()
    x = big[:, :Cin, :, 2:2 + W]                                   # strided view
    w = (torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5).cuda()
    bias = torch.randn(Cout, generator=g).cuda()
    res = torch.randn(B, Cp, H, W, generator=g).cuda().contiguous(memory_format=torch.channels_last)
    assert conv.small_ok(Cin, Cout, k, 1, k // 2)
    y = conv.conv_small_fwd(x, w, Cp, bias=bias, residual=res, lrelu=True, slope=0.2)
    ref = F.leaky_relu(F.conv2d(x.double(), w.double(), bias.double(), padding=k // 2), 0.2) + res[:, :Cout].double()
    assert y.shape == (B, Cp, H, W) and y.is_contiguous(memory_format=torch.channels_last)
    assert (y[:, :Cout].double() - ref).abs().max().item() < 1e-5 * ref.abs().max().item()
    assert (y[:, Cout:] == 0).all()                                # padding channels are zeros
    y2 = conv.conv_small_fwd(x, w, Cp, round_tf32=True)
    assert torch.equal(y2, conv.tf32_round(conv.conv_small_fwd(x, w, Cp)))
    dy = torch.randn(B, Cp, H, W, generator=g).cuda().contiguous(memory_format=torch.channels_last)
    dx = conv.conv_small_dgrad(dy, w, Cin)
    dx_ref = torch.nn.grad.conv2d_input((B, Cin, H, W), w.double(), dy[:, :Cout].double(), padding=k // 2)
    assert (dx.double() - dx_ref).abs().max().item() < 1e-5 * dx_ref.abs().max().item()
    dw = conv.conv_small_wgrad(dy, x, (Cout, Cin, k, k))
    dw_ref = torch.nn.grad.conv2d_weight(x.double(), (Cout, Cin, k, k), dy[:, :Cout].double(), padding=k // 2)
    assert (dw.double() - dw_ref).abs().max().item() < 2e-5 * dw_ref.abs().max().item()
    assert torch.equal(dw, conv.conv_small_wgrad(dy, x, (Cout, Cin, k, k)))      # deterministic
