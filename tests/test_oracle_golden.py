"""CPU: the oracle restatement reproduces the reference's golden vectors
(tests/golden/*.npz were produced by oracle/make_golden.py running the
unmodified reference)."""
import pytest
import torch

from oracle import hist_oracle as ho
from tests import parity


@pytest.mark.parametrize("name", parity.golden_names("hist_"))
def test_hist_oracle_matches_golden(name):
    g = parity.load_golden(name)
    hist, loss, grad = ho.hist_loss_and_grad(g["x"], g["target"], g["alpha"], **g["kwargs"])
    # same ops, same machine class: identical up to a last-bit reduction-order effect
    assert parity.rel_err(hist, g["hist"]).max().item() < 2e-6
    assert parity.fro_rel(hist, g["hist"]) < 1e-7
    if loss is not None:
        assert abs(float(loss) - g["loss"]) <= 1e-6 * abs(g["loss"])
        parity.assert_grad(grad, g["grad_x"], name)
        glin = ho.hist_linear_grad(g["x"], g["target"], **g["kwargs"])
        parity.assert_grad(glin, g["grad_x_lin"], name + " (linear)")
    else:
        l2 = ho.hellinger_loss(g["target"], hist, g["alpha"])
        assert abs(float(l2) - g["loss"]) <= 1e-6 * abs(g["loss"])


def test_hist_invariants():
    x = ho.synth_generator_like(3, 40)
    h = ho.rgb_uv_hist(x)
    assert h.shape == (3, 3, 64, 64) and h.dtype == torch.float32
    s = h.sum(dim=(1, 2, 3))
    assert torch.allclose(s, torch.ones(3), atol=1e-4)
    assert (h > 0).all()
    # pixel-permutation invariance
    perm = torch.randperm(40 * 40, generator=torch.Generator().manual_seed(3))
    xp = x.reshape(3, 3, -1)[:, :, perm].reshape(3, 3, 40, 40)
    assert parity.fro_rel(ho.rgb_uv_hist(xp), h) < 1e-6
    # all-black image stays finite
    hb = ho.rgb_uv_hist(torch.zeros(1, 3, 8, 8))
    assert torch.isfinite(hb).all()


@pytest.mark.parametrize("name", parity.golden_names("chroma_") + parity.golden_names("lab_"))
def test_chroma_oracle_matches_golden(name):
    g = parity.load_golden(name)
    fn = ho.lab_hist if name.startswith("lab_") else ho.rg_chroma_hist
    hist = fn(g["x"], **g["kwargs"])
    assert hist.shape == g["hist"].shape
    assert parity.rel_err(hist, g["hist"], 1e-7).max().item() < 2e-6
    x = g["x"].clone().requires_grad_(True)
    (gx,) = torch.autograd.grad((fn(x, **g["kwargs"]) * g["target"]).sum(), x)
    parity.assert_grad(gx, g["grad_x_lin"], name)
