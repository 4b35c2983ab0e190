"""GPU: the ReHistoGAN recolouring path -- the new kernels one by one against
plain torch fp32 references, then the whole generator phase (fused modules + product losses)
against the reference-made golden vectors and the CPU TF32-emulation noise floor."""
import math

import pytest
import torch
import torch.nn.functional as F

from tests import rehisto_checks as rc
from tests.emulation import emulated_conv
from tests.gan_checks import rel

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape", [(2, 32, 16, 16), (3, 64, 33, 20), (2, 96, 64, 64), (1, 1024, 4, 4)])
def test_instnorm_lrelu_kernel_matches_torch(shape, cuda_device):
    from histogan_b200.rehistogan import instnorm_lrelu
    torch.manual_seed(0)
    x = (torch.randn(shape, device="cuda") * 1.7 + 0.4).contiguous(memory_format=torch.channels_last)
    w = torch.randn(shape, device="cuda")
    xa = x.clone().requires_grad_(True)
    y = instnorm_lrelu(xa, 0.2)
    (y * w).sum().backward()
    xb = x.clone().requires_grad_(True)
    ref = F.leaky_relu(F.instance_norm(xb, eps=1e-5), 0.2)
    (ref * w).sum().backward()
    assert rel(y, ref) < 1e-5                       # fp32 tolerance: different summation order only
    # the kernel hands dx to a tensor-core conv: it is TF32-rounded (2^-11 relative per element)
    assert rel(xa.grad, xb.grad) < 4e-4
    yr = instnorm_lrelu(x, 0.2, round_out=True)
    assert rel(yr, ref) < 4e-4 and bool(((yr.view(torch.int32) & 0x1FFF) == 0).all())


def test_instnorm_padded_channels_stay_zero(cuda_device):
    from histogan_b200.rehistogan import instnorm_lrelu
    x = torch.zeros(2, 32, 8, 8, device="cuda").contiguous(memory_format=torch.channels_last)
    x[:, :16] = torch.randn(2, 16, 8, 8, device="cuda")
    assert float(instnorm_lrelu(x)[:, 16:].abs().max()) == 0.0


@pytest.mark.parametrize("B,S", [(2, 64), (3, 50), (1, 256)])
def test_laplacian_l1_kernel_matches_torch(B, S, cuda_device):
    from histogan_b200.rehistogan import reconstruction_loss
    from oracle import rehisto_oracle as ro
    torch.manual_seed(1)
    a = torch.rand(B, 3, S, S, device="cuda")
    b = (a + 0.1 * torch.randn(B, 3, S, S, device="cuda")).requires_grad_(True)
    loss = 1.5 * reconstruction_loss("2nd gradient").compute_loss(a, b)
    loss.backward()
    bc = b.detach().cpu().requires_grad_(True)
    ref = 1.5 * ro.reconstruction_loss(a.cpu(), bc, "laplacian")
    ref.backward()
    assert abs(loss.item() - ref.item()) <= 1e-5 * abs(ref.item())         # fp32 tolerance
    assert rel(b.grad, bc.grad) < 1e-6
    for kind, name in (("1st gradient", "sobel"), ("L1", None)):
        got = reconstruction_loss(kind).compute_loss(a, b.detach())
        want = ro.reconstruction_loss(a.cpu(), b.detach().cpu(), name)
        assert abs(got.item() - want.item()) <= 1e-5 * abs(want.item())


@pytest.mark.parametrize("B,S", [(2, 64), (1, 256), (2, 37)])
def test_gaussian_filter_kernel_matches_torch(B, S, cuda_device):
    from histogan_b200.rehistogan import gaussian_op, get_gaussian_kernel
    torch.manual_seed(2)
    k = get_gaussian_kernel(kernel_size=15, sigma=5, channels=3)
    x = torch.rand(B, 3, S, S)
    w = torch.randn(B, 3, S - 14, S - 14)
    xa = x.cuda().requires_grad_(True)
    y = gaussian_op(xa, kernel=get_gaussian_kernel(kernel_size=15, sigma=5, channels=3).cuda())
    (y * w.cuda()).sum().backward()
    xb = x.clone().requires_grad_(True)
    ref = k(xb)
    (ref * w).sum().backward()
    assert y.shape == ref.shape
    assert rel(y, ref) < 1e-5 and rel(xa.grad, xb.grad) < 1e-5              # fp32 tolerance


def test_upsample2x_matches_torch(cuda_device):
    from histogan_b200.rehistogan import _Upsample2x
    torch.manual_seed(3)
    x = torch.randn(2, 64, 8, 8, device="cuda").contiguous(memory_format=torch.channels_last)
    w = torch.randn(2, 64, 16, 16, device="cuda")
    xa = x.clone().requires_grad_(True)
    y = _Upsample2x.apply(xa)
    (y * w).sum().backward()
    xb = x.clone().requires_grad_(True)
    ref = F.interpolate(xb, scale_factor=2, mode="bilinear", align_corners=False)
    (ref * w).sum().backward()
    assert rel(y, ref) < 4e-4 and rel(xa.grad, xb.grad) < 1e-5              # y is TF32-rounded


def _product_losses(images, hists, gen, D):
    from histogan_b200.rehistogan import recoloringTrainer
    from oracle import make_golden_rehisto as mr
    t = _product_losses.trainer
    if t is None:
        t = recoloringTrainer("t", "gpurun_out/rehisto_results", "gpurun_out/rehisto_models", mr.IMAGE_SIZE,
                              mr.CAPACITY, batch_size=mr.B, skip_conn_to_GAN=True, **{
                                  "hist_" + k if k != "h" else "hist_bin": v for k, v in mr.HIST_KW.items()})
        _product_losses.trainer = t

    class _G:          # g_losses only touches GAN.D
        pass
    t.GAN = _G()
    t.GAN.D = D
    d, h, r, v = t.g_losses(images, hists, gen, mr.ALPHA, mr.BETA, mr.GAMMA)
    return dict(d_loss=d, hist_loss=h, rec_loss=r, var_loss=v)


_product_losses.trainer = None


def test_generator_phase_against_reference_golden(cuda_device, monkeypatch):
    """fused ED / recolouring head / discriminator + the product's own loss kernels."""
    from histogan_b200 import rehistogan as rh
    calls = {"n": 0}
    real = rh._InstNormLReLU.apply
    monkeypatch.setattr(rh._InstNormLReLU, "apply",
                        staticmethod(lambda *a, **k: (calls.__setitem__("n", calls["n"] + 1), real(*a, **k))[1]))
    e_gpu, out_gpu = rc.g_phase_errors("cuda", _product_losses)
    assert calls["n"] > 0, "the fused encoder path must be the one running"
    with emulated_conv(round_operands=True):
        e_emu, out_emu = rc.g_phase_errors("cpu", rc.oracle_losses)
    print("rehisto vs golden (GPU):", {k: f"{v:.2e}" for k, v in e_gpu.items()})
    print("rehisto vs golden (CPU TF32 emulation):", {k: f"{v:.2e}" for k, v in e_emu.items()})
    # activations: the encoder-decoder puts 7..22 TF32 convolutions (with instance norms) in
    # series, whose rounding alone is 0.9e-3 .. 1.5e-3 (the emulation row); the CUDA path must
    # (a) stay within 1e-3 (north_star) of that same-arithmetic emulation and (b) be no
    # further from the fp32 golden than the emulation's floor allows
    for k in ("latent", "p1", "p2", "generated"):
        assert rel(out_gpu[k], out_emu[k]) < 1e-3, (k, rel(out_gpu[k], out_emu[k]))
    for k in ("latent", "p1", "p2", "generated", "ed_rgb"):
        assert e_gpu[k] < 2 * e_emu[k] + 1e-3, (k, e_gpu[k], e_emu[k])
    # loss terms evaluated on the GPU's own generated image; the Hellinger term amplifies the
    # TF32 difference of that image, the others are well conditioned
    for k in ("d_loss", "rec_loss", "var_loss"):
        assert e_gpu[k] < 2 * e_emu[k] + 2e-3, (k, e_gpu[k], e_emu[k])
    assert e_gpu["hist_loss"] < 2 * e_emu["hist_loss"] + 1e-2
    # loss kernels differentiated at the reference's generated image: fp32-level agreement
    assert e_gpu["dgen_rec"] < 1e-5 and e_gpu["dgen_var"] < 1e-4 and e_gpu["dgen_hist"] < 1e-3
    assert e_gpu["dgen_d"] < 2 * e_emu["dgen_d"] + 2e-3
    assert e_gpu["param_grads_max"] < 2 * e_emu["param_grads_max"] + 2e-3, \
        (e_gpu["param_grads_max"], e_emu["param_grads_max"])


def test_recoloring_train_steps_run_and_learn(cuda_device, tmp_path):
    """two full recoloringTrainer.train steps (D phase with gradient penalty + G phase) at the
    CLI defaults, small image size: finite losses, every trainable parameter moves."""
    from histogan_b200.rehistogan import recoloringTrainer
    from histogan_b200.trainer import SyntheticLoader
    torch.manual_seed(0)
    t = recoloringTrainer("r", str(tmp_path / "res"), str(tmp_path / "mod"), 64, 16, batch_size=4,
                          skip_conn_to_GAN=True, initialize_gan=True, save_every=10 ** 9, fast_rng=True)
    t.loader = SyntheticLoader(4, 64, seed=0)
    t.steps = 1                                           # no checkpoint at step 0
    t.init_GAN()
    before = {k: v.detach().clone() for k, v in t.GAN.named_parameters()}
    for _ in range(4):
        t.train()
    assert t.steps == 5
    for v in (t.d_loss, t.g_loss, t.r_loss, t.h_loss, t.var_loss, t.last_gp_loss):
        assert math.isfinite(v)
    unmoved = [k for k, v in t.GAN.named_parameters()
               if torch.equal(v, before[k]) and "conv_out_rgb" not in k]
    assert not unmoved, unmoved


def test_recoloring_graph_path(cuda_device, tmp_path, monkeypatch):
    """cuda_graphs=True: captured D / G phases reproduce the eager phases (same noise) and a few
    full steps run (both D variants: with and without the gradient penalty)."""
    from histogan_b200 import rehistogan as rh
    from histogan_b200.trainer import SyntheticLoader
    torch.manual_seed(0)
    t = rh.recoloringTrainer("g", str(tmp_path / "res"), str(tmp_path / "mod"), 64, 16, batch_size=4,
                             skip_conn_to_GAN=True, initialize_gan=True, save_every=10 ** 9,
                             fast_rng=True, cuda_graphs=True)
    t.loader = SyntheticLoader(4, 64, seed=0)
    t.init_GAN()
    t.GAN.train()
    batch = next(t.loader)
    t._static = {'images': batch['images'].clone(), 'hists': batch['histograms'].clone()}
    fixed = torch.rand(4, 64, 64, 1, device='cuda')
    monkeypatch.setattr(rh.torch, "rand", lambda *a, **k: fixed)
    d_params = list(t.GAN.D.parameters())
    # (biases in front of an instance norm have exactly-zero gradients: rounding noise only)
    g_params = [p for k, p in t.GAN.named_parameters()
                if not k.startswith("D.") and not rc.is_prenorm_bias(k)]
    for key, fn, params in ((('D', True), lambda: t._phase_d(True), d_params),
                            (('D', False), lambda: t._phase_d(False), d_params),
                            (('G', 32.0, 1.5, 4.0), lambda: t._phase_g(32.0, 1.5, 4.0), g_params)):
        out_e = [o.clone() for o in fn() if o is not None]
        g_e = [p.grad.detach().clone() for p in params if p.grad is not None]
        out_g = [o.clone() for o in t._graphed(key, fn, params) if o is not None]
        g_g = [p.grad.detach().clone() for p in params if p.grad is not None]
        for a, b in zip(out_e, out_g):
            assert abs(a.item() - b.item()) <= 1e-3 * abs(a.item()) + 1e-6, (key, a.item(), b.item())
        assert len(g_e) == len(g_g) > 0
        worst = max(((a - b).norm() / b.norm().clamp_min(1e-20)).item() for a, b in zip(g_g, g_e))
        print(key, "max grad rel diff graph vs eager", worst)
        assert worst < 2e-2, (key, worst)
    monkeypatch.undo()
    t._graphs.clear()
    t._static = None
    t.steps = 3
    before = {k: v.detach().clone() for k, v in t.GAN.named_parameters()}
    for _ in range(4):                       # steps 3, 4 (gradient penalty), 5, 6
        t.train()
    assert t.steps == 7 and t.graph_replayed_launches > 0
    for v in (t.d_loss, t.g_loss, t.r_loss, t.h_loss, t.var_loss, t.last_gp_loss):
        assert math.isfinite(v)
    unmoved = [k for k, v in t.GAN.named_parameters()
               if torch.equal(v, before[k]) and "conv_out_rgb" not in k]
    assert not unmoved, unmoved
