"""GPU debug: G-phase gradients of the path-length step per LOSS TERM, product path vs demod-by-torch"""
import os, sys, tempfile, pathlib, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
import histogan_b200.fused as fz
from histogan_b200.hist import hellinger_loss
from histogan_b200.trainer import styles_def_to_tensor, set_requires_grad, EPS
from oracle import make_golden_step as mgs, train_oracle as to
from tests.test_trainer_gpu import _golden_trainer

real = fz.grouped_linear


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


def demod_torch(xs, ws, bs, flags=0, slope=0.2, eps=1e-8):
    return (torch_fwd if flags & fz.LIN_RSQRT_EPS else real)(xs, ws, bs, flags, slope, eps)


with tempfile.TemporaryDirectory() as tmp:
    t = _golden_trainer(pathlib.Path(tmp))
    GAN = t.GAN
    GAN.train()
    images, hists = mgs.step_inputs(32)
    L = int(math.log2(mgs.IMAGE_SIZE) - 1) - 2
    mgs.seed_step(32)
    dr = to.draw_step_inputs(mgs.BATCH, L, 512, mgs.IMAGE_SIZE, path_penalty=True)
    hist_b = hists[1].cuda()
    named = [(k, p) for k, p in GAN.G.named_parameters()]
    params = [p for _, p in named]
    set_requires_grad(GAN.D, False)

    def terms(with_pl_pass):
        h_w = GAN.H(hist_b).unsqueeze(1)
        h_w = torch.cat((h_w, h_w), dim=1)
        w_styles = styles_def_to_tensor([(GAN.S(z.cuda()), n) for z, n in dr["g_style"]])
        nz = dr["g_noise"].cuda()
        fake = GAN.G(w_styles, h_w, nz)
        fake_out, _ = GAN.D(fake)
        out = {"D": fake_out.mean(), "hist": hellinger_loss(hist_b, t.histBlock(F.relu(fake)), 2.0)}
        if with_pl_pass:
            std = 0.1 / (w_styles.std(dim=0, keepdim=True) + EPS)
            pl_images = GAN.G(w_styles + dr["pl_noise"].cuda() / (std + EPS), h_w, nz)
            out["PL"] = (((pl_images - fake) ** 2).mean(dim=(1, 2, 3)) ** 2).mean()
        return out

    res = {}
    for variant, impl in (("product", real), ("demod-torch", demod_torch)):
        fz.grouped_linear = impl
        for with_pl in (False, True):
            tl = terms(with_pl)
            keys = list(tl)
            for i, k in enumerate(keys):
                gs = torch.autograd.grad(tl[k], params, retain_graph=True, allow_unused=True)
                res[(variant, with_pl, k)] = [g.detach().clone() if g is not None else None for g in gs]
            gs = torch.autograd.grad(sum(tl.values()), params, allow_unused=True)
            res[(variant, with_pl, "sum")] = [g.detach().clone() if g is not None else None for g in gs]
    for with_pl in (False, True):
        for k in ("D", "hist", "PL", "sum"):
            if (("product", with_pl, k)) not in res:
                continue
            ga, gb = res[("product", with_pl, k)], res[("demod-torch", with_pl, k)]
            rows = []
            for (n, _), a, b in zip(named, ga, gb):
                if a is None or b is None:
                    continue
                cos = (a.flatten() @ b.flatten() / (a.norm() * b.norm()).clamp_min(1e-30)).item()
                rows.append((n, round(cos, 4), round((a.norm() / b.norm().clamp_min(1e-30)).item(), 3)))
            bad = [r for r in rows if r[1] < 0.99]
            print(f"[second G pass present: {with_pl}] term {k}: {len(bad)}/{len(rows)} tensors differ between product and "
                  f"demod-by-torch: {bad[:6]}", flush=True)
    # linearity inside each variant
    for variant in ("product", "demod-torch"):
        s = res[(variant, True, "sum")]
        parts = [res[(variant, True, k)] for k in ("D", "hist", "PL")]
        worst = max(((f - sum(p[i] for p in parts)).norm() / f.norm().clamp_min(1e-30)).item() for i, f in enumerate(s) if f is not None)
        print(f"[{variant}] |sum - (D + hist + PL)| / |sum| worst: {worst:.3e}")
