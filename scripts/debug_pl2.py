"""GPU debug: are the demodulation factors / modulations of the two generator passes of a path-length
step intact when the backward runs?  (records them at forward time, compares after backward)"""
import os, sys, tempfile, pathlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import histogan_b200.fused as fz
from histogan_b200 import gan
from oracle import make_golden_step as mgs
from tests import step_checks as sc
from tests.test_trainer_gpu import _golden_trainer

rec = []
real_demod, real_style = fz.demod_all, fz.style_mods


def demod_rec(mods, wsqs, eps=1e-8):
    out = real_demod(mods, wsqs, eps)
    ref = [torch.rsqrt(m.detach().double().pow(2) @ w.double().t() + eps).float() for m, w in zip(mods, wsqs)]
    err = max(((o - r).abs() / r.abs()).max().item() for o, r in zip(out, ref))
    rec.append(("d", out, [o.clone() for o in out], err))
    return out


def style_rec(per_block, linears):
    out = real_style(per_block, linears)
    ref = [torch.nn.functional.linear(per_block[b].double(), l.weight.double(), l.bias.double()).float() + 1
           for b, l in linears]
    err = max((o - r).abs().max().item() for o, r in zip(out, ref))
    rec.append(("mod", out, [o.detach().clone() for o in out], err))
    return out


fz.demod_all, fz.style_mods = demod_rec, style_rec
with tempfile.TemporaryDirectory() as tmp:
    t = _golden_trainer(pathlib.Path(tmp))
    images, hists = mgs.step_inputs(32)
    t.loader = iter([{"images": images, "histograms": hists[0]}, {"images": images, "histograms": hists[1]}])
    t.steps, t.pl_mean = 32, 0
    mgs.seed_step(32)
    t.train(alpha=mgs.ALPHA)
    torch.cuda.synchronize()
    for i, (kind, live, snap, err) in enumerate(rec):
        changed = max((a.detach() - b).abs().max().item() for a, b in zip(live, snap))
        ptrs = [a.data_ptr() for a in live]
        print(f"call {i} {kind}: forward err vs torch {err:.2e}; changed after backward by {changed:.3e}; "
              f"{len(set(ptrs))}/{len(ptrs)} distinct buffers; shapes {[tuple(a.shape) for a in live][:4]}", flush=True)
    # do buffers of different calls overlap?
    spans = []
    for i, (kind, live, _, _) in enumerate(rec):
        for a in live:
            spans.append((a.data_ptr(), a.data_ptr() + a.numel() * 4, i, kind))
    spans.sort()
    over = [(s1, s2) for s1, s2 in zip(spans, spans[1:]) if s2[0] < s1[1]]
    print("overlapping buffers:", over[:5])
