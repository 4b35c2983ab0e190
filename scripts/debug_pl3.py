"""GPU debug: path-length gradient split by pass, product path vs the torch-forward variant"""
import os, sys, tempfile, pathlib, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
import histogan_b200.fused as fz
from histogan_b200 import gan
from histogan_b200.trainer import styles_def_to_tensor, EPS
from oracle import make_golden_step as mgs, train_oracle as to
from tests.test_trainer_gpu import _golden_trainer

real_fwd = fz.grouped_linear


def torch_fwd(xs, ws, bs, flags=0, slope=0.2, eps=1e-8):
    out = []
    for x, w, b in zip(xs, ws, bs):
        xe = x * x if flags & fz.LIN_SQUARE_INPUT else x
        v = xe @ w.t()
        if b is not None: v = v + b
        if flags & fz.LIN_RSQRT_EPS: v = torch.rsqrt(v + eps)
        if flags & fz.LIN_LRELU: v = F.leaky_relu(v, slope)
        if flags & fz.LIN_ADD_ONE: v = v + 1
        out.append(v.contiguous())
    return out


with tempfile.TemporaryDirectory() as tmp:
    t = _golden_trainer(pathlib.Path(tmp))
    GAN = t.GAN
    GAN.train()
    images, hists = mgs.step_inputs(32)
    L = int(math.log2(mgs.IMAGE_SIZE) - 1) - 2
    mgs.seed_step(32)
    dr = to.draw_step_inputs(mgs.BATCH, L, 512, mgs.IMAGE_SIZE, path_penalty=True)
    hist_b = hists[1].cuda()
    params = [p for p in GAN.G.parameters()]
    names = [k for k, _ in GAN.G.named_parameters()]

    def grads(which):
        h_w = GAN.H(hist_b).unsqueeze(1)
        h_w = torch.cat((h_w, h_w), dim=1)
        w_styles = styles_def_to_tensor([(GAN.S(z.cuda()), n) for z, n in dr["g_style"]])
        nz = dr["g_noise"].cuda()
        fake = GAN.G(w_styles, h_w, nz)
        std = 0.1 / (w_styles.std(dim=0, keepdim=True) + EPS)
        pl_images = GAN.G(w_styles + dr["pl_noise"].cuda() / (std + EPS), h_w, nz)
        a = fake.detach() if which == "p2" else fake
        b = pl_images.detach() if which == "p1" else pl_images
        loss = (((b - a) ** 2).mean(dim=(1, 2, 3)) ** 2).mean()
        gs = torch.autograd.grad(loss, params, allow_unused=True)
        return [g.detach().clone() if g is not None else None for g in gs], fake.detach().clone(), pl_images.detach().clone()

    res = {}
    for variant in ("product", "torchfwd"):
        fz.grouped_linear = real_fwd if variant == "product" else torch_fwd
        for which in ("full", "p1", "p2"):
            res[(variant, which)] = grads(which)
    for which in ("full", "p1", "p2"):
        ga, fa, pa = res[("product", which)]
        gb, fb, pb = res[("torchfwd", which)]
        print(f"[{which}] forward: fake diff {((fa - fb).norm() / fb.norm()).item():.2e}  pl diff {((pa - pb).norm() / pb.norm()).item():.2e}")
        rows = []
        for n, a, b in zip(names, ga, gb):
            if a is None or b is None:
                rows.append((n, "None", a is None, b is None)); continue
            cos = (a.flatten() @ b.flatten() / (a.norm() * b.norm()).clamp_min(1e-30)).item()
            rows.append((n, round(cos, 4), round((a.norm() / b.norm().clamp_min(1e-30)).item(), 3)))
        bad = [r for r in rows if r[1] == "None" or r[1] < 0.99]
        print(f"   {len(bad)}/{len(rows)} tensors differ (cos < 0.99): {bad[:10]}", flush=True)
    # consistency: full == p1 + p2 ?
    for variant in ("product", "torchfwd"):
        gf, g1, g2 = res[(variant, "full")][0], res[(variant, "p1")][0], res[(variant, "p2")][0]
        worst = max(((f - (a + b)).norm() / f.norm().clamp_min(1e-30)).item() for f, a, b in zip(gf, g1, g2) if f is not None)
        print(f"[{variant}] |full - (p1 + p2)| / |full| worst over tensors: {worst:.3e}")
