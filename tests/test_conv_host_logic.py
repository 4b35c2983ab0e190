"""CPU: host-side logic of the convolution path that needs no GPU -- the parity-class
decomposition of the stride-2 input gradient, and the channels_last parameter storage."""
import torch
import torch.nn.functional as F

from histogan_b200 import conv, ops


def test_stride2_input_gradient_as_parity_class_convolutions():
    """dx of a 3x3 / stride 2 / pad 1 conv == four small convolutions over dy (1, 2, 2, 4 taps)
    written to dx[:, :, py::2, px::2] -- what ops._raw_grad_input launches on the GPU."""
    g = torch.Generator().manual_seed(0)
    for B, Cin, Cout, H in ((2, 5, 7, 8), (1, 3, 4, 12)):
        w = torch.randn(Cout, Cin, 3, 3, generator=g)
        dy = torch.randn(B, Cout, H // 2, H // 2, generator=g)
        ref = F.conv_transpose2d(dy, w, stride=2, padding=1, output_padding=1)
        dx = torch.zeros(B, Cin, H, H)
        for py in (0, 1):
            for px in (0, 1):
                w2 = ops._stride2_class_weight(w, py, px)          # (Cin, Cout, th, tw)
                th, tw = w2.shape[2:]
                assert (th, tw) == (1 + py, 1 + px)
                # the kernel runs this as a stride-1, pad-0 conv whose windows may reach one
                # row / column past dy (zero-filled by TMA): same as padding bottom/right here
                dyp = F.pad(dy, (0, tw - 1, 0, th - 1))
                dx[:, :, py::2, px::2] = F.conv2d(dyp, w2)
        assert torch.allclose(dx, ref, atol=1e-5, rtol=1e-5)


def test_conv_weights_are_channels_last_with_reference_state_dict():
    from histogan_b200.gan import Conv2DMod, Discriminator, Generator
    from histogan_b200.rehistogan import RecoloringEncoderDecoder
    for m in (Generator(32, 512, 16), Discriminator(32, 16),
              RecoloringEncoderDecoder(64, 16, skip_conn_to_GAN=True)):
        for k, p in m.named_parameters():
            if p.dim() == 4:
                assert conv.is_ohwi(p), k
        # a reference checkpoint holds plain contiguous tensors: loading keeps the values, the
        # shapes and this package's layout; saving gives the same values back
        ref_sd = {k: torch.randn(v.shape) for k, v in m.state_dict().items()}
        m.load_state_dict(ref_sd)
        for k, v in m.state_dict().items():
            assert v.shape == ref_sd[k].shape and torch.equal(v, ref_sd[k]), k
        for k, p in m.named_parameters():
            if p.dim() == 4:
                assert conv.is_ohwi(p), k
    c = Conv2DMod(8, 4, 3)
    assert c.weight.shape == (4, 8, 3, 3) and c.weight.stride() == (72, 1, 24, 8)


def test_wgrad_layout_matches_parameter_layout():
    """the weight-gradient kernel's native [Cout][k][k][Cin] output viewed as (Cout,Cin,k,k) has
    exactly the strides of a channels_last parameter (so autograd accumulates without a copy)"""
    dwp = torch.empty(16, 3, 3, 32)
    dw = dwp.permute(0, 3, 1, 2)
    p = torch.empty(16, 32, 3, 3).contiguous(memory_format=torch.channels_last)
    assert dw.shape == p.shape and dw.stride() == p.stride()
    from histogan_b200.optim import _dense, _same_layout
    assert _dense(p) and _same_layout(p, dw)
    assert not _same_layout(p, torch.empty(16, 32, 3, 3))
