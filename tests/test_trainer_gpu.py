"""GPU: Trainer.train end to end at a small configuration (image 32, capacity 16):
both phases, gradient penalty (step 0), path-length regulariser (step 0), checkpoint
save/load with the reference's state_dict layout."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _trainer(tmp_path, **kw):
    from histogan_b200.trainer import SyntheticLoader, Trainer
    t = Trainer("t", str(tmp_path / "results"), str(tmp_path / "models"), image_size=32,
                network_capacity=16, batch_size=4, hist_insz=150, hist_resizing="interpolation",
                save_every=1000, **kw)
    t.loader = SyntheticLoader(4, 32, seed=0)
    t.loader_evaluate = SyntheticLoader(4, 32, seed=1, eval_batch=4)
    return t


def test_train_steps(tmp_path, cuda_device):
    torch.manual_seed(0)
    t = _trainer(tmp_path)
    t.train(alpha=2)                  # step 0: GP + PL + save + evaluate
    sd0 = {k: v.clone() for k, v in t.GAN.state_dict().items()}
    assert (tmp_path / "models" / "t" / "model_0.pt").exists()
    assert (tmp_path / "results" / "t" / "0-ema.jpg").exists()
    for _ in range(3):
        t.train(alpha=2)
    assert t.steps == 4
    for name in ("d_loss", "g_loss", "h_loss", "last_gp_loss"):
        v = getattr(t, name)
        assert v == v and abs(v) < 1e6, (name, v)
    changed = [k for k, v in t.GAN.state_dict().items()
               if k.split(".")[0] in ("G", "D", "S", "H") and not torch.equal(v, sd0[k])]
    assert len(changed) > 50
    # reference checkpoint layout (histoGAN.py:1120-1122): raw HistoGAN.state_dict()
    ck = torch.load(t.model_name(0), map_location="cpu")
    assert {"G.initial_block", "G.blocks.0.conv1.weight", "G.blocks.0.to_rgb.conv.weight",
            "D.blocks.0.conv_res.weight", "D.blocks.0.net.0.bias", "D.to_logit.weight",
            "S.net.0.weight", "H.fcs.0.weight", "GE.initial_block", "SE.net.0.weight",
            "HE.fcs.0.weight"} <= set(ck)
    assert ck["G.blocks.0.conv1.weight"].shape == (256, 64, 3, 3)
    t.load(0)
    assert t.steps == 0


def test_histogram_loss_first_order_descent(tmp_path, cuda_device):
    """End-to-end directional derivative: an SGD step on the generator sized for a small
    predicted decrease of the histogram loss must deliver about that decrease.  Exercises
    G forward -> relu -> fused hist + Hellinger -> hist backward -> conv dgrad/wgrad."""
    from histogan_b200 import RGBuvHistBlock, hellinger_loss
    from histogan_b200.trainer import HistoGAN, image_noise, styles_def_to_tensor
    torch.manual_seed(1)
    gan = HistoGAN(image_size=32, network_capacity=16)
    blk = RGBuvHistBlock(insz=150)
    target = torch.rand(4, 3, 64, 64, device="cuda")
    target = target / target.sum(dim=(1, 2, 3), keepdim=True)
    z = torch.randn(4, 512, device="cuda")
    nz = image_noise(4, 32, fast=True)
    with torch.no_grad():
        w = styles_def_to_tensor([(gan.S(z), 2)])
        hw = gan.H(target).unsqueeze(1).repeat(1, 2, 1)

    def loss_fn():
        return hellinger_loss(target, blk(torch.relu(gan.G(w, hw, nz))), 2.0)

    params = [p for p in gan.G.parameters()]
    loss0 = loss_fn()
    grads = torch.autograd.grad(loss0, params, allow_unused=True)
    pairs = [(p, g) for p, g in zip(params, grads) if g is not None]
    gnorm2 = sum((g.double() ** 2).sum().item() for _, g in pairs)
    base = [p.detach().clone() for p, _ in pairs]
    ratios = {}
    for frac in (1e-2, 3e-3, 1e-3, 5e-4):
        eta = frac * loss0.item() / gnorm2
        with torch.no_grad():
            for (p, g), b in zip(pairs, base):
                p.copy_(b - eta * g)
            loss1 = loss_fn()
        ratios[frac] = (loss0.item() - loss1.item()) / (frac * loss0.item())
    print("loss0", loss0.item(), "actual/predicted decrease by step size:", ratios)
    # descent at every step size; first-order agreement once the step is small compared with
    # the kernel width (sigma = 0.02 in log-chroma makes the loss strongly curved) yet large
    # compared with the TF32 quantum of the packed weights (below it most weights do not move)
    assert all(r > 0 for r in ratios.values()), ratios
    assert 0.5 < sorted(ratios[f] for f in (3e-3, 1e-3, 5e-4))[1] < 1.5, ratios


def test_fused_diffgrad_matches_foreach(cuda_device):
    """hg_diffgrad_step (one fused multi-tensor kernel) == the torch._foreach restatement."""
    from histogan_b200.optim import DiffGrad
    torch.manual_seed(0)
    shapes = [(300, 7), (5,), (64, 32, 3, 3), (1,), (100000,)] * 12      # > 48 tensors: two launches
    pa = [torch.nn.Parameter(torch.randn(s, device="cuda")) for s in shapes]
    pb = [torch.nn.Parameter(p.detach().cpu().clone()) for p in pa]         # CPU -> _foreach path
    oa, ob = DiffGrad(pa, lr=2e-4, betas=(0.5, 0.9)), DiffGrad(pb, lr=2e-4, betas=(0.5, 0.9))
    for _ in range(3):
        for a, b in zip(pa, pb):
            g = torch.randn(a.shape)
            a.grad, b.grad = g.cuda(), g.clone()
        oa.step(); ob.step()
    for a, b in zip(pa, pb):
        assert torch.allclose(a.detach().cpu(), b.detach(), rtol=1e-5, atol=1e-7)
    # in-place update through raw pointers must still bump the version counter (the packed
    # weight caches and autograd's saved-tensor checks depend on it)
    v0 = pa[0]._version
    for a in pa:
        a.grad = torch.randn_like(a)
    oa.step()
    assert pa[0]._version > v0


def test_optimizer_step_invalidates_packed_weights(cuda_device):
    from histogan_b200 import ops
    from histogan_b200.optim import DiffGrad
    torch.manual_seed(0)
    w = torch.nn.Parameter(torch.randn(32, 32, 3, 3, device="cuda") / 17)
    x = torch.randn(2, 32, 8, 8, device="cuda")
    y0 = ops.conv2d(x, w, None, 1, 1).detach().clone()
    w.grad = torch.randn_like(w)
    DiffGrad([w], lr=1e-2).step()
    y1 = ops.conv2d(x, w, None, 1, 1).detach()
    ref = torch.nn.functional.conv2d(x, w.detach(), padding=1)
    assert (y1 - y0).abs().max() > 1e-3                     # new weights are used ...
    assert ((y1 - ref).norm() / ref.norm()).item() < 2e-3   # ... and they are the current ones


def test_cuda_graph_training_path(tmp_path, cuda_device):
    """cuda_graphs=True: D phase (with / without gradient penalty) and G phase replayed as
    CUDA graphs; PL steps stay eager.  Must train like the eager path: finite losses,
    moving weights, and the same loss scale as an eager trainer fed the same data."""
    torch.manual_seed(0)
    tg = _trainer(tmp_path / "g", cuda_graphs=True, fast_rng=True)
    te = _trainer(tmp_path / "e", fast_rng=True)
    for t in (tg, te):
        t.init_GAN()
    te.GAN.load_state_dict(tg.GAN.state_dict())
    for t in (tg, te):
        t.steps = 2501
        for _ in range(9):                 # 2501..2509: graphs incl. two GP steps
            t.train(alpha=2)
    assert set(tg._graphs) >= {('D', False), ('D', True), ('G', 2.0)}
    for name in ("d_loss", "g_loss", "h_loss", "last_gp_loss"):
        a, b = getattr(tg, name), getattr(te, name)
        assert a == a and abs(a) < 1e7, (name, a)
        print(name, "graphed", a, "eager", b)
    assert abs(tg.h_loss - te.h_loss) < 0.25 * abs(te.h_loss) + 0.05
    # both trainers started from the same weights: after 9 steps the discriminator must have
    # learned comparably (a graph that updates the wrong gradient buffers leaves d_loss at its
    # initial ~110) and the penalty read-out must be a valid non-negative number
    assert 0.2 < tg.d_loss / te.d_loss < 5, (tg.d_loss, te.d_loss)
    assert tg.last_gp_loss >= 0
    # the read-out must be THIS trainer's penalty, not whatever a later graph left in the shared
    # pool: same order of magnitude as the eager trainer's (same data, same initial weights)
    assert 0.05 < tg.last_gp_loss / te.last_gp_loss < 20, (tg.last_gp_loss, te.last_gp_loss)
    moved = sum(not torch.equal(p, q) for p, q in zip(tg.GAN.G.parameters(), te.GAN.G.parameters()))
    assert moved > 10          # different random latents -> different but comparable trajectories
    tg.steps = 2528                        # a path-length step runs eagerly inside a graphed trainer
    tg.train(alpha=2)


def test_graph_replay_equals_eager_phase(tmp_path, cuda_device, monkeypatch):
    """the captured D phase (with gradient penalty) and G phase reproduce the eager phases:
    same losses and same parameter gradients when fed the same latents / noise."""
    from histogan_b200 import trainer as T
    torch.manual_seed(0)
    t = _trainer(tmp_path, cuda_graphs=True, fast_rng=True)
    t.init_GAN()
    t.GAN.train()
    B, S_, L = 4, 32, t.GAN.G.num_layers - 2
    t._static = {
        'images': torch.rand(B, 3, S_, S_, device='cuda'),
        'hists': torch.rand(B, 3, 64, 64, device='cuda'),
        'mask': (torch.arange(L, device='cuda') < 1).float(),
        'mask_host': torch.ones(L).pin_memory(),
    }
    t._static['hists'] /= t._static['hists'].sum(dim=(1, 2, 3), keepdim=True)
    fixed = {"randn": [torch.randn(B, 512, device='cuda') for _ in range(2)],
             "rand": torch.rand(B, S_, S_, 1, device='cuda')}
    calls = {"n": 0}

    def fake_randn(*a, **k):
        calls["n"] += 1
        return fixed["randn"][(calls["n"] - 1) % 2]

    monkeypatch.setattr(T.torch, "randn", fake_randn)
    monkeypatch.setattr(T.torch, "rand", lambda *a, **k: fixed["rand"])

    def grads(params):
        return [p.grad.detach().clone() for p in params if p.grad is not None]

    for key, fn, params in ((('D', True), lambda: t._phase_d(True), list(t.GAN.D.parameters())),
                            (('D', False), lambda: t._phase_d(False), list(t.GAN.D.parameters())),
                            (('G', 2.0), lambda: t._phase_g(2.0), list(t.GAN.G.parameters()))):
        out_e = [o.clone() for o in fn() if o is not None]
        g_e = grads(params)
        out_g = [o.clone() for o in t._graphed(key, fn, params) if o is not None]     # capture + replay
        g_g = grads(params)
        held = [p.grad for p in params]
        out_g2 = [o.clone() for o in t._graphed(key, fn, params) if o is not None]    # second replay
        # every replay re-attaches the graph's own gradient tensors to the parameters
        assert all(p.grad is h for p, h in zip(params, held))
        for a, b in zip(out_e, out_g):
            assert abs(a.item() - b.item()) <= 1e-3 * abs(a.item()) + 1e-6, (key, a.item(), b.item())
        for a, b in zip(out_g, out_g2):
            assert abs(a.item() - b.item()) <= 1e-3 * abs(a.item()) + 1e-6, (key, "replay drift")
        assert len(g_e) == len(g_g) > 0
        worst = max(((a - b).norm() / b.norm().clamp_min(1e-20)).item() for a, b in zip(g_g, g_e))
        print(key, "losses", [o.item() for o in out_e], "max grad rel diff graph vs eager", worst)
        assert worst < 2e-2, (key, worst)      # fp32 atomics reorder sums; TF32 rounding flips
    # the variants own different buffers: replaying an older graph must hand ITS gradients over
    d_params = list(t.GAN.D.parameters())
    t._graphed(('D', True), None, d_params)
    with_gp = [p.grad for p in d_params]
    t._graphed(('D', False), None, d_params)
    assert all(a is not b for a, b in zip(with_gp, [p.grad for p in d_params]))
    t._graphed(('D', True), None, d_params)
    assert all(a is b for a, b in zip(with_gp, [p.grad for p in d_params]))
