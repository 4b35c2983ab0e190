"""GPU parity tests of the RGB-uv histogram block + Hellinger loss: the CUDA
path (through the reference-shaped class API -> C ABI) against the CPU oracle
and the reference's golden vectors.  Tolerances: tests/parity.py."""
import pytest
import torch
import torch.nn.functional as F

from oracle import hist_oracle as ho
from tests import parity

pytestmark = pytest.mark.gpu


def _block(kwargs):
    from histogan_b200 import RGBuvHistBlock
    kw = {k: (list(v) if isinstance(v, list) else v) for k, v in kwargs.items()}
    return RGBuvHistBlock(device="cuda", **kw)


def _cuda_loss_and_grads(g, fused):
    from histogan_b200 import hellinger_loss
    x = g["x"].cuda().requires_grad_(True)
    t = g["target"].cuda()
    hist = _block(g["kwargs"])(F.relu(x))
    if fused:
        loss = hellinger_loss(t, hist, g["alpha"])
    else:   # the reference's own expression, histoGAN/histoGAN.py:957-960
        loss = g["alpha"] * ho.SCALE * torch.sqrt(torch.sum(torch.pow(
            torch.sqrt(t) - torch.sqrt(hist), 2))) / t.shape[0]
    return x, t, hist, loss


@pytest.mark.parametrize("name", parity.golden_names("hist_"))
def test_forward_against_reference_golden(name, cuda_device):
    g = parity.load_golden(name)
    hist = _block(g["kwargs"])(F.relu(g["x"].cuda()))
    assert hist.shape == g["hist"].shape and hist.dtype == torch.float32
    atol = 1e-6 if g["kwargs"].get("method") == "RBF" else 1e-9   # RBF underflows to ~0
    m = parity.assert_hist_e2e(hist, g["hist"], name, atol_frac=atol)
    print(name, m)


@pytest.mark.parametrize("fused", [True, False])
@pytest.mark.parametrize("name", parity.golden_names("hist_"))
def test_loss_and_grad_against_reference_golden(name, fused, cuda_device):
    g = parity.load_golden(name)
    # thresholding: the reference's gradient is finite only for functionals without sqrt'(0);
    # its Hellinger gradient is NaN everywhere (72 % of the hard bins are empty) -- the NaN
    # pattern must agree (parity.assert_grad)
    x, t, hist, loss = _cuda_loss_and_grads(g, fused)
    parity.assert_loss(loss.item(), g["loss"], name)
    loss.backward()
    parity.assert_grad(x.grad, g["grad_x"], name)


@pytest.mark.parametrize("name", parity.golden_names("hist_"))
def test_linear_functional_grad(name, cuda_device):
    g = parity.load_golden(name)
    x = g["x"].cuda().requires_grad_(True)
    hist = _block(g["kwargs"])(F.relu(x))
    (hist * g["target"].cuda()).sum().backward()
    # thresholding included: masks are constants, the gradient flows through Iy
    parity.assert_grad(x.grad, g["grad_x_lin"], name)
    if g["kwargs"].get("method") == "thresholding":
        assert g["grad_x_lin"].abs().max() > 0
        kw = dict(g["kwargs"], intensity_scale=False)     # reference: hist.requires_grad == False
        x2 = g["x"].cuda().requires_grad_(True)
        (_block(kw)(F.relu(x2)) * g["target"].cuda()).sum().backward()
        assert (x2.grad == 0).all()


CASES = [
    ("uniform", ho.synth_uniform, 4, 64, dict()),
    ("genlike", ho.synth_generator_like, 2, 96, dict()),
    ("genlike_interp", ho.synth_generator_like, 2, 200, dict(insz=150)),
    ("signed_interp_nonsquare", ho.synth_signed, 2, 131, dict(insz=100)),
]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_strict_kernel_arithmetic(case, cuda_device):
    """Oracle fed the GPU's resized pixels and float32 logs: isolates the kernel's
    soft-binning + accumulation arithmetic -> max element-wise rel <= 1e-5."""
    from histogan_b200 import RGBuvHistBlock, device_logf, hist_preprocess
    _, maker, B, S, kw = case
    x = maker(B, S, seed=11)
    xc = x.cuda()
    hist = RGBuvHistBlock(device="cuda", **kw)(xc)
    insz = kw.get("insz", 150)
    pix = hist_preprocess(xc, 64, insz, "interpolation")          # (B,3,N)
    n = pix.shape[-1]
    side = int(round(n ** 0.5))
    pix_img = pix.reshape(B, 3, side, side).cpu()
    ref = ho.rgb_uv_hist(pix_img, preprocessed=True,
                         log_fn=lambda v: device_logf(v.cuda()).cpu(), **kw)
    parity.assert_hist_strict(hist, ref, case[0])
    # and the hooks themselves are within float32 rounding of torch's CPU ops
    cpu_pix = ho.preprocess(x, 64, insz, "interpolation").reshape(B, 3, -1)
    # the resize reproduces torch's CPU bilinear kernel bit-for-bit
    assert (pix.cpu() != cpu_pix).float().mean().item() < 1e-4
    v = torch.rand(1 << 16, generator=torch.Generator().manual_seed(1)) + 1e-6
    dl = device_logf(v.cuda()).cpu()
    ulp = torch.abs(torch.nextafter(dl, dl + 1) - dl)
    assert ((dl - torch.log(v)).abs() <= ulp).all()
    assert (dl != torch.log(v)).float().mean().item() < 0.01


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_e2e_against_oracle(case, cuda_device):
    from histogan_b200 import RGBuvHistBlock, hellinger_loss
    _, maker, B, S, kw = case
    x = maker(B, S, seed=12)
    t = ho.synth_random_target(B, seed=13)
    # the oracle evaluates log with the device's logf (see parity.DeviceLog); the host's own
    # libm is compared at 10x the tolerance at the end
    ref_hist, ref_loss, ref_grad = ho.hist_loss_and_grad(x, t, 2.0, log_fn=parity.device_log, **kw)
    xc = x.cuda().requires_grad_(True)
    hist = RGBuvHistBlock(device="cuda", **kw)(F.relu(xc))
    loss = hellinger_loss(t.cuda(), hist, 2.0)
    loss.backward()
    print(case[0], parity.assert_hist_e2e(hist, ref_hist, case[0]))
    parity.assert_loss(loss.item(), ref_loss.item(), case[0])
    print(case[0], parity.assert_grad(xc.grad, ref_grad, case[0]))
    host_hist, _, _ = ho.hist_loss_and_grad(x, t, 2.0, **kw)
    if kw.get("method") != "thresholding":          # hard bins flip on a 1-ulp change of u
        assert parity.fro_rel(hist, host_hist) <= 10 * parity.E2E_FRO_REL


def test_layouts_and_channels(cuda_device):
    """channels_last input, >3 channels (alpha dropped, RGBuvHistBlock.py:98-99),
    int device argument (histoGAN.py:134), empty batch."""
    from histogan_b200 import RGBuvHistBlock
    x = ho.synth_generator_like(2, 48, seed=3, C=4)
    ref = ho.rgb_uv_hist(x, log_fn=parity.device_log)
    blk = RGBuvHistBlock(device=0)
    for xin in (x.cuda(), x.cuda().contiguous(memory_format=torch.channels_last),
                x.cuda()[:, :, ::1, :].transpose(2, 3).contiguous().transpose(2, 3)):
        xin = xin.detach().requires_grad_(True)
        h = blk(xin)
        parity.assert_hist_e2e(h, ref, "layout")
        h.sum().backward()
        assert xin.grad.shape == x.shape and (xin.grad[:, 3] == 0).all()
    assert blk(torch.zeros(0, 3, 8, 8, device="cuda")).shape == (0, 3, 64, 64)
    hb = blk(torch.zeros(1, 3, 8, 8, device="cuda"))              # all-black stays finite
    assert torch.isfinite(hb).all() and abs(hb.sum().item() - 1) < 1e-3


def test_error_behaviour(cuda_device):
    from histogan_b200 import RGBuvHistBlock
    x = torch.rand(1, 3, 200, 200, device="cuda")
    with pytest.raises(Exception, match="Wrong resizing method"):
        RGBuvHistBlock(resizing="nearest")(x)
    RGBuvHistBlock(resizing="nearest")(x[:, :, :64, :64])   # only raised when a resize is needed
    with pytest.raises(Exception, match="Wrong kernel method"):
        RGBuvHistBlock(method="gaussian")(x)


def test_clamp_and_relu_masks(cuda_device):
    """clamp gradient mask is inclusive at 0 and 1; values outside get zero grad."""
    from histogan_b200 import RGBuvHistBlock
    x = ho.synth_uniform(1, 16, seed=2)
    x[0, 0, 0, :5] = torch.tensor([-0.5, 0.0, 0.5, 1.0, 1.5])
    w = ho.synth_random_target(1, seed=4)
    ref = ho.hist_linear_grad(x, w, apply_relu=False)
    xc = x.cuda().requires_grad_(True)
    (RGBuvHistBlock()(xc) * w.cuda()).sum().backward()
    got = xc.grad.cpu()
    assert got[0, 0, 0, 0] == 0 and got[0, 0, 0, 4] == 0
    assert got[0, 0, 0, 1] != 0 and got[0, 0, 0, 3] != 0
    parity.assert_grad(got, ref, "clamp mask")


@pytest.mark.parametrize("name", parity.golden_names("chroma_") + parity.golden_names("lab_"))
def test_rg_chroma_block_against_reference_golden(name, cuda_device):
    """rgChromaHistBlock / LabHistBlock on the generic CUDA kernels vs the
    unmodified reference classes."""
    from histogan_b200 import rgChromaHistBlock, LabHistBlock
    if name.startswith("lab_"):
        rgChromaHistBlock = LabHistBlock
    g = parity.load_golden(name)
    kw = {k: (list(v) if isinstance(v, list) else v) for k, v in g["kwargs"].items()}
    x = g["x"].cuda().requires_grad_(True)
    hist = rgChromaHistBlock(device="cuda", **kw)(x)
    assert hist.shape == g["hist"].shape
    atol = 1e-6 if kw.get("method") == "RBF" else 1e-9
    print(name, parity.assert_hist_e2e(hist, g["hist"], name, atol_frac=atol))
    (hist * g["target"].cuda()).sum().backward()
    parity.assert_grad(x.grad, g["grad_x_lin"], name)
