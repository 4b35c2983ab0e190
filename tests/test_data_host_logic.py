"""CPU: the image-folder data source of Trainer.set_data_src with the histogram
block replaced by the CPU oracle -- decoding, resizing / cropping, the convex mix of two target
histograms and the evaluation loader."""
import numpy as np
import torch

from histogan_b200 import data
from oracle import hist_oracle as ho


class _OracleBlock:
    device = "cpu"

    def __call__(self, x):
        return ho.rgb_uv_hist(x, h=64, insz=150, resizing="sampling")


class _FakeTrainer:
    histBlock = _OracleBlock()
    batch_size = 3
    image_size = 32


def test_folder_batches(tmp_path):
    from PIL import Image
    rng = np.random.default_rng(0)
    for i, (w, h) in enumerate(((48, 40), (70, 33), (36, 36))):
        Image.fromarray(rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)).save(tmp_path / f"im{i}.png")
    train, evaluate = data.make_loaders(_FakeTrainer(), str(tmp_path))
    b = next(train)
    assert b["images"].shape == (3, 3, 32, 32) and b["images"].dtype == torch.float32
    assert 0.0 <= float(b["images"].min()) and float(b["images"].max()) <= 1.0
    assert b["histograms"].shape == (3, 3, 64, 64)
    s = b["histograms"].sum(dim=(1, 2, 3))
    assert torch.allclose(s, torch.ones(3), atol=1e-4)            # mixes of normalised histograms
    e = next(evaluate)
    assert set(e) == {"histograms"} and e["histograms"].shape == (4, 3, 64, 64)
    # an evaluation histogram is exactly the block's histogram of one of the source images
    singles = [ho.rgb_uv_hist(train._load(p).unsqueeze(0), h=64, insz=150, resizing="sampling")[0]
               for p in train.paths]
    assert all(any(torch.allclose(h, s1, atol=1e-7) for s1 in singles) for h in e["histograms"])
