"""CPU: the functional G/D oracle reproduces the reference's golden vectors, and
(in the build container) the live reference classes."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import gan_oracle as go
from oracle import make_golden_gan as mg
from oracle import ref_shim
from tests import parity


def _load(name):
    z = np.load(os.path.join(parity.GOLDEN_DIR, name))
    return {k: z[k] for k in z.files}


def _rel(a, b):
    a = torch.as_tensor(np.asarray(a)).double()
    b = torch.as_tensor(np.asarray(b)).double()
    return ((a - b).norm() / b.norm().clamp_min(1e-300)).item()


def test_generator_oracle_matches_golden():
    g = _load("gan_generator_32.npz")
    sd = go.seeded_state_dict(json.loads(str(g["shapes"])), seed=1)
    inp = mg.gan_inputs()
    styles = inp["styles"].clone().requires_grad_(True)
    hists = inp["hists"].clone().requires_grad_(True)
    sd = {k: v.requires_grad_(True) for k, v in sd.items()}
    rgb, acts = go.generator(sd, styles, hists, inp["noise"], mg.IMAGE_SIZE, return_all=True)
    assert _rel(rgb.detach(), g["rgb"]) < 1e-5
    assert np.allclose([a.norm().item() for a in acts], g["act_norms"], rtol=1e-5)
    (rgb * inp["w_rgb"]).sum().backward()
    assert _rel(styles.grad, g["g_styles"]) < 1e-4
    assert _rel(hists.grad, g["g_hists"]) < 1e-4
    assert _rel(sd["initial_block"].grad, g["g_initial_block"]) < 1e-4
    norms = json.loads(str(g["param_grad_norms"]))
    for k, v in norms.items():
        assert abs(sd[k].grad.norm().item() - v) <= 1e-4 * max(v, 1e-6), k


def test_discriminator_oracle_matches_golden():
    g = _load("gan_discriminator_32.npz")
    sd = go.seeded_state_dict(json.loads(str(g["shapes"])), seed=2)
    sd = {k: v.requires_grad_(True) for k, v in sd.items()}
    images = mg.gan_inputs()["images"].clone().requires_grad_(True)
    logits = go.discriminator(sd, images, mg.IMAGE_SIZE)
    assert _rel(logits.detach(), g["logits"]) < 1e-5
    gp = mg.gradient_penalty(images, logits)
    assert abs(gp.item() - float(g["gp"])) < 1e-4 * float(g["gp"])
    (logits.sum() + gp * mg.GP_WEIGHT_IN_TEST).backward()
    assert _rel(images.grad, g["g_images"]) < 1e-3
    assert _rel(sd["to_logit.weight"].grad, g["g_to_logit_w"]) < 1e-3


@pytest.mark.skipif(not ref_shim.available(), reason="reference not mounted")
def test_oracle_vs_live_reference_conv2dmod():
    gm = ref_shim.ref_gan_module()
    torch.manual_seed(0)
    for cin, cout, k, demod in [(8, 12, 3, True), (16, 3, 1, False)]:
        m = gm.Conv2DMod(cin, cout, k, demod=demod)
        x, y = torch.randn(3, cin, 6, 6), torch.randn(3, cin)
        assert _rel(go.mod_conv(x, y, m.weight.detach(), demod), m(x, y).detach()) < 1e-6


@pytest.mark.skipif(not ref_shim.available(), reason="reference not mounted")
def test_oracle_vs_live_reference_mlps():
    gm = ref_shim.ref_gan_module()
    torch.manual_seed(0)
    S = gm.StyleVectorizer(32, 3)
    H = gm.HistVectorizer(4, 16, 3)
    z, h = torch.randn(2, 32), torch.rand(2, 3, 4, 4)
    assert _rel(go.mlp(S.state_dict(), "net", 3, z), S(z).detach()) < 1e-6
    assert _rel(go.mlp(H.state_dict(), "fcs", 3, h.reshape(2, -1)), H(h).detach()) < 1e-6
