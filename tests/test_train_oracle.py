"""CPU: the train-step oracle (oracle/train_oracle.py) against the golden produced by the
UNMODIFIED reference ``Trainer.train`` (histoGAN/histoGAN.py:853-1020) -- loss composition,
random-draw order, and every parameter gradient of both phases, for a plain step, a
gradient-penalty step and a gradient-penalty + path-length step."""
import math

import pytest
import torch

from oracle import gan_oracle as go
from oracle import make_golden_step as mgs
from oracle import train_oracle as to
from tests import step_checks as sc


def _state_dicts():
    from histogan_b200.gan import Discriminator, Generator, HistVectorizer, StyleVectorizer
    with torch.device("meta"):
        mods = {"G": Generator(mgs.IMAGE_SIZE, 512, mgs.CAPACITY), "D": Discriminator(mgs.IMAGE_SIZE, mgs.CAPACITY),
                "S": StyleVectorizer(512, 8), "H": HistVectorizer(64, 512, 8)}
    out = {}
    for name, m in mods.items():
        shapes = {k: list(v.shape) for k, v in m.state_dict().items()}
        out[name] = {k: v.requires_grad_(True) for k, v in go.seeded_state_dict(shapes, mgs.SEEDS[name]).items()}
    return out


@pytest.mark.parametrize("alpha", [2.0, 0.0])
@pytest.mark.parametrize("case", mgs.CASES)
def test_step_oracle_matches_reference_trainer(case, alpha):
    g = sc.load_golden()
    sd = _state_dicts()
    ck = sc.key(case, alpha)
    ref = g[f"{ck}_scalars"]
    images, hists = mgs.step_inputs(case)
    layers = int(math.log2(mgs.IMAGE_SIZE) - 1) - 2
    mgs.seed_step(case)
    draws = to.draw_step_inputs(mgs.BATCH, layers, 512, mgs.IMAGE_SIZE, path_penalty=case % 32 == 0)
    d = to.d_phase(sd["G"], sd["D"], sd["S"], sd["H"], images, hists[0], draws["d_style"], draws["d_noise"],
                   mgs.IMAGE_SIZE, apply_gp=case % 4 == 0)
    assert sc.rel(d["divergence"], ref["d_loss"]) < 1e-5
    if case % 4 == 0:
        assert sc.rel(d["gp"], ref["gp"]) < 1e-4
    td = sc.compare_grads(d["grads"], g["names_d"], g[f"{ck}_d_norms"], g[f"{ck}_d_samples"])
    print("D grads worst:", sc.worst(td))
    assert all(v[0] < 1e-3 and v[1] > 1 - 1e-6 for v in td.values()), sc.worst(td)

    gph = to.g_phase(sd["G"], sd["D"], sd["S"], sd["H"], hists[1], draws["g_style"], draws["g_noise"],
                     mgs.IMAGE_SIZE, alpha, hist_kw=dict(insz=150, resizing="interpolation"),
                     pl_noise=draws["pl_noise"], pl_mean=0)
    assert sc.rel(gph["loss"], ref["g_loss"]) < 1e-5
    assert sc.rel(gph["hist_loss"], ref["h_loss"]) < 1e-5 if alpha else float(gph["hist_loss"]) == 0.0
    if case % 32 == 0:
        assert sc.rel(0.01 * gph["avg_pl"], ref["pl_mean"]) < 1e-4      # EMA(0.99) from pl_mean = 0
    tg = sc.compare_grads(gph["grads"], g["names_g"], g[f"{ck}_g_norms"], g[f"{ck}_g_samples"])
    print("G grads worst:", sc.worst(tg))
    assert all(v[0] < 1e-3 and v[1] > 1 - 1e-6 for v in tg.values()), sc.worst(tg)


def test_diffgrad_scalar_oracle_matches_foreach_restatement():
    """the from-the-paper scalar DiffGrad (oracle) against optim.DiffGrad's torch path"""
    from histogan_b200.optim import DiffGrad
    torch.manual_seed(0)
    p = torch.nn.Parameter(torch.randn(37, dtype=torch.float64))
    opt = DiffGrad([p], lr=2e-4, betas=(0.5, 0.9))
    ref, state = p.detach().tolist(), {}
    for _ in range(5):
        gr = torch.randn(37, dtype=torch.float64)
        p.grad = gr.clone()
        opt.step()
        ref = to.diffgrad_step(ref, gr.tolist(), state, lr=2e-4, betas=(0.5, 0.9))
    assert torch.allclose(p.detach(), torch.tensor(ref, dtype=torch.float64), rtol=1e-12, atol=1e-15)
