"""Shared helpers for the train-step parity tests: the golden of ONE reference
``Trainer.train`` call (tests/golden/train_step_64.npz, made by the unmodified reference
through oracle/make_golden_step.py) and comparisons of gradient sets against its
fingerprints (float64 norm + up to 256 strided entries per parameter tensor)."""
import json
import os

import numpy as np
import torch

from oracle import make_golden_step as mgs
from tests import parity


def load_golden():
    z = np.load(os.path.join(parity.GOLDEN_DIR, "train_step_64.npz"))
    g = {k: z[k] for k in z.files}
    g["names_d"] = json.loads(str(g["names_d"]))
    g["names_g"] = json.loads(str(g["names_g"]))
    for c in mgs.CASES:
        for suffix, _ in mgs.ALPHAS:
            g[f"c{c}{suffix}_scalars"] = json.loads(str(g[f"c{c}{suffix}_scalars"]))
    return g


def key(case, alpha):
    """golden key prefix of (trainer.steps, histogram-loss weight)"""
    return f"c{case}" + {2.0: "", 0.0: "n"}[float(alpha)]


def compare_grads(grads, names, norms, samples):
    """grads: {name: tensor} or list in `names` order.  Returns a table
    {name: (norm_rel_err, cosine_on_samples, max_abs_err_on_samples / max_abs_ref)}."""
    table, off = {}, 0
    for i, nm in enumerate(names):
        g = grads[nm] if isinstance(grads, dict) else grads[i]
        assert g is not None, f"no gradient for {nm}"
        n, s = mgs.fingerprint(g)
        ref = samples[off:off + s.size].astype(np.float64)
        off += s.size
        s = s.astype(np.float64)
        denom = np.linalg.norm(s) * np.linalg.norm(ref)
        cos = float(s @ ref / denom) if denom > 0 else (1.0 if not s.any() and not ref.any() else 0.0)
        scale = max(np.abs(ref).max(), 1e-30)
        table[nm] = (abs(n - norms[i]) / max(norms[i], 1e-30), cos, float(np.abs(s - ref).max() / scale))
    assert off == samples.size
    return table


def worst(table):
    wn = max(table.items(), key=lambda kv: kv[1][0])
    wc = min(table.items(), key=lambda kv: kv[1][1])
    wa = max(table.items(), key=lambda kv: kv[1][2])
    return {"norm_rel": (wn[0], wn[1][0]), "cosine": (wc[0], wc[1][1]), "max_abs_rel": (wa[0], wa[1][2])}


def rel(a, b):
    return abs(float(a) - float(b)) / max(abs(float(b)), 1e-30)


def emulated_step_tables(case, golden=None, alpha=None):
    """TF32 noise floor of ONE train step: the histogan_b200 modules (same algorithm as the CUDA path:
    activation-side modulation, shared weights) evaluated on the CPU with torch stand-ins for the
    conv primitives and TF32-rounded operands (tests/emulation.py), the histogram block replaced by
    the CPU oracle; same weights / inputs / random draws as the reference golden.  Returns
    (scalars, D-gradient table, G-gradient table) against that golden."""
    import math
    import torch.nn.functional as F
    from histogan_b200 import gan
    from histogan_b200.trainer import gradient_penalty, styles_def_to_tensor, EPS
    from oracle import hist_oracle as ho
    from oracle import train_oracle as to
    from tests.emulation import emulated_conv
    g = golden or load_golden()
    alpha = mgs.ALPHA if alpha is None else alpha
    ck = key(case, alpha)
    S = mgs.IMAGE_SIZE
    with torch.device("cpu"):
        mods = {"G": gan.Generator(S, 512, mgs.CAPACITY), "D": gan.Discriminator(S, mgs.CAPACITY),
                "S": gan.StyleVectorizer(512, 8), "H": gan.HistVectorizer(64, 512, 8)}

    class _Box:
        pass
    box = _Box()
    for k, m in mods.items():
        setattr(box, k, m)
    sd = mgs.seeded_gan_state(box)
    for k, m in mods.items():
        m.load_state_dict({n[len(k) + 1:]: v for n, v in sd.items() if n.startswith(k + ".")})
    images, hists = mgs.step_inputs(case)
    L = int(math.log2(S) - 1) - 2
    mgs.seed_step(case)
    dr = to.draw_step_inputs(mgs.BATCH, L, 512, S, path_penalty=case % 32 == 0)
    Gm, Dm, Sm, Hm = mods["G"], mods["D"], mods["S"], mods["H"]

    def w_of(style):
        return styles_def_to_tensor([(Sm(z), n) for z, n in style])

    def hw_of(h):
        v = Hm(h).unsqueeze(1)
        return torch.cat((v, v), dim=1)

    with emulated_conv(round_operands=True):
        with torch.no_grad():
            fake = Gm(w_of(dr["d_style"]), hw_of(hists[0]), dr["d_noise"])
        img = images.clone().requires_grad_(True)
        fo, _ = Dm(fake)
        ro, _ = Dm(img)
        div = (F.relu(1 + ro) + F.relu(1 - fo)).mean()
        loss, gp = div, None
        if case % 4 == 0:
            gp = gradient_penalty(img, ro)
            loss = loss + gp
        dgr = torch.autograd.grad(loss, list(Dm.parameters()))
        w_styles, hw = w_of(dr["g_style"]), hw_of(hists[1])
        fake = Gm(w_styles, hw, dr["g_noise"])
        fo, _ = Dm(fake)
        hl = ho.hellinger_loss(hists[1], ho.rgb_uv_hist(F.relu(fake), insz=150), alpha)
        gl = fo.mean()
        gen = gl + hl
        pl = None
        if case % 32 == 0:
            std = 0.1 / (w_styles.std(dim=0, keepdim=True) + EPS)
            pli = Gm(w_styles + dr["pl_noise"] / (std + EPS), hw, dr["g_noise"])
            pll = ((pli - fake) ** 2).mean(dim=(1, 2, 3))
            pl = pll.detach().mean().item()
            gen = gen + (pll ** 2).mean()
        gparams = list(Gm.parameters()) + list(Sm.parameters()) + list(Hm.parameters())
        ggr = torch.autograd.grad(gen, gparams)
    ref = g[f"{ck}_scalars"]
    scal = {"d_loss": rel(div, ref["d_loss"]), "g_loss_abs_over_dscale": abs(gl.item() - ref["g_loss"]) / abs(ref["d_loss"]),
            "h_loss": rel(hl, ref["h_loss"]) if ref["h_loss"] else abs(float(hl))}
    if gp is not None:
        scal["gp"] = rel(gp, ref["gp"])
    if pl is not None:
        scal["pl_mean"] = rel(0.01 * pl, ref["pl_mean"])
    td = compare_grads(list(dgr), g["names_d"], g[f"{ck}_d_norms"], g[f"{ck}_d_samples"])
    tg = compare_grads(list(ggr), g["names_g"], g[f"{ck}_g_norms"], g[f"{ck}_g_samples"])
    return scal, td, tg


def rel_error(entry):
    """|g - ref| / |ref| (on the stored entries) from a table entry (norm_rel, cosine, ...): with
    r = |g|/|ref| = 1 +- norm_rel,  e^2 = 1 + r^2 - 2 r cos.  One number for "how far", so that a tensor
    whose floor is mostly a direction error is not failed for a norm error of the same size."""
    import math
    nr, cos = float(entry[0]), float(entry[1])
    return max(math.sqrt(max(1 + (1 + nr) ** 2 - 2 * (1 + nr) * cos, 0.0)),
               math.sqrt(max(1 + (1 - nr) ** 2 - 2 * (1 - nr) * cos, 0.0)))


def within_floor(tab, tab_emu, margin=3e-2):
    """{tensor: (measured, floor)} for every tensor whose relative error vs the golden exceeds 2x the
    relative error of the TF32 emulation of the same algorithm (+ margin)"""
    return {k: (v, tab_emu[k]) for k, v in tab.items()
            if rel_error(v) > 2 * rel_error(tab_emu[k]) + margin}
