"""GPU: the target-histogram producer of Trainer.set_data_src (SURVEY 8f-2,
histoGAN/histoGAN.py:263-266,292-302) on the CUDA histogram block: batched per image size,
equal to the CPU oracle's histogram of every source image."""
import numpy as np
import pytest
import torch

from oracle import hist_oracle as ho
from tests import parity

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("resizing", ["sampling", "interpolation"])
def test_folder_target_histograms_on_device(resizing, tmp_path, cuda_device):
    from PIL import Image
    from histogan_b200 import RGBuvHistBlock, data, _lib
    rng = np.random.default_rng(5)
    sizes = [(200, 180)] * 5 + [(96, 96)] * 3           # (w, h): two groups, the first needs resizing
    for i, (w, h) in enumerate(sizes):
        img = (rng.random((h, w, 3)) ** 2 * 255).astype(np.uint8)
        Image.fromarray(img).save(tmp_path / f"im{i}.png")

    class Tr:
        histBlock = RGBuvHistBlock(insz=150, h=64, resizing=resizing, device="cuda")
        batch_size, image_size = 4, 64

    train, evaluate = data.make_loaders(Tr(), str(tmp_path))
    lib = _lib.load()
    n0 = lib.hg_launch_count()
    e = next(evaluate)                                   # 4 single-image histograms
    assert e["histograms"].is_cuda and e["histograms"].shape == (4, 3, 64, 64)
    singles = {p: ho.rgb_uv_hist(train._load(p).unsqueeze(0), h=64, insz=150, resizing=resizing)[0]
               for p in train.paths}
    for hk in e["histograms"]:
        best = min(parity.fro_rel(hk, s) for s in singles.values())
        assert best < 5e-6, best
    b = next(train)
    assert b["images"].shape == (4, 3, 64, 64) and b["histograms"].shape == (4, 3, 64, 64)
    s = b["histograms"].sum(dim=(1, 2, 3))
    assert torch.allclose(s, torch.ones_like(s), atol=1e-4)
    # every mixed target lies in the convex hull of two single-image histograms: positive, and
    # reproducible from the loader's own draws
    assert (b["histograms"] > 0).all()
    assert lib.hg_launch_count() > n0
