"""CPU, build container only: the oracle against the live reference module over
more configurations than the committed goldens, including real JPEGs shipped
with the reference (both resize branches).  Skipped where /root/reference is
absent (the GPU box)."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import hist_oracle as ho
from oracle import ref_shim
from tests import parity

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="reference not mounted")

CONFIGS = [
    dict(), dict(method="RBF"), dict(method="thresholding"), dict(resizing="sampling", insz=32),
    dict(insz=40), dict(green_only=True), dict(intensity_scale=False),
    dict(hist_boundary=[4, -2]), dict(h=16, sigma=0.1), dict(h=100, insz=32),
]


@pytest.mark.parametrize("kw", CONFIGS, ids=[str(c) for c in CONFIGS])
def test_forward_matches_reference(kw):
    mod = ref_shim.ref_hist_module()
    x = ho.synth_signed(2, 56, seed=5, C=4)
    ref = mod.RGBuvHistBlock(device="cpu", **{k: (list(v) if isinstance(v, list) else v)
                                              for k, v in kw.items()})(x)
    mine = ho.rgb_uv_hist(x, **kw)
    assert mine.shape == ref.shape
    assert parity.rel_err(mine, ref).max().item() < 2e-6


def test_backward_matches_reference():
    mod = ref_shim.ref_hist_module()
    x = ho.synth_signed(2, 40, seed=7)
    t = ho.synth_random_target(2)
    xr = x.clone().requires_grad_(True)
    h = mod.RGBuvHistBlock(device="cpu", insz=32)(torch.relu(xr))
    loss = 2 * (1 / np.sqrt(2.0)) * torch.sqrt(torch.sum(torch.pow(torch.sqrt(t) - torch.sqrt(h), 2))) / 2
    loss.backward()
    _, l2, gx = ho.hist_loss_and_grad(x, t, 2.0, insz=32)
    assert abs(float(loss) - float(l2)) < 1e-6 * abs(float(loss))
    parity.assert_grad(gx, xr.grad, "oracle autograd vs reference autograd")


def test_real_images():
    from PIL import Image
    mod = ref_shim.ref_hist_module()
    files = sorted(glob.glob(os.path.join(ref_shim.REF_ROOT, "target_images", "*.jpg")))[:3]
    assert files
    for f in files:
        img = torch.from_numpy(np.asarray(Image.open(f).convert("RGB"), dtype=np.float32) / 255.0)
        x = img.permute(2, 0, 1).unsqueeze(0).contiguous()
        for kw in (dict(insz=150), dict(insz=250, resizing="sampling")):
            ref = mod.RGBuvHistBlock(device="cpu", **kw)(x)
            mine = ho.rgb_uv_hist(x, **kw)
            assert parity.rel_err(mine, ref).max().item() < 2e-6


def test_data_asset_invariants():
    """histogram_data/histograms.npy: positive, sums to 1."""
    a = np.load(os.path.join(ref_shim.REF_ROOT, "histogram_data", "histograms.npy"))
    assert a.shape[1:] == (1, 3, 64, 64) and (a > 0).all()
    assert np.allclose(a.reshape(a.shape[0], -1).sum(1), 1.0, atol=1e-4)
