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
        g[f"c{c}_scalars"] = json.loads(str(g[f"c{c}_scalars"]))
    return g


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
