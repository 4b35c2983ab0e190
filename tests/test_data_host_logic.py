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


def _make_images(tmp_path, sizes):
    from PIL import Image
    rng = np.random.default_rng(1)
    for i, (w, h) in enumerate(sizes):
        Image.fromarray(rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)).save(tmp_path / f"im{i:02d}.png")


def test_epoch_without_replacement_and_batched_calls(tmp_path):
    """reference loader semantics (shuffle=True, drop_last=True, histoGAN.py:840-845): every image
    once per epoch; the 2 x batch histogram sources are evaluated with ONE block call per
    distinct image size."""
    _make_images(tmp_path, [(40, 40)] * 4 + [(52, 36)] * 2)
    calls = []

    class Blk(_OracleBlock):
        def __call__(self, x):
            calls.append(tuple(x.shape))
            return super().__call__(x)

    class Tr(_FakeTrainer):
        histBlock = Blk()

    train, _ = data.make_loaders(Tr(), str(tmp_path))
    seen = []
    orig = train._next_indices
    train._next_indices = lambda: (seen.append(orig()) or seen[-1])
    next(train); n_calls = len(calls); next(train)
    assert sorted(int(i) for b in seen for i in b) == list(range(6))        # one epoch = every image once
    assert n_calls <= 2 and all(c[0] >= 1 for c in calls)                   # <= one call per distinct size
    assert sum(c[0] for c in calls[:n_calls]) == 6                          # 2 sources x batch of 3


def test_ranks_see_disjoint_shards(tmp_path, monkeypatch):
    """every rank used to draw the same indices; now the epoch permutation is shared
    and sharded, the histogram sources are per-rank."""
    _make_images(tmp_path, [(32, 32)] * 12)
    loaders = []
    for r in (0, 1):
        monkeypatch.setattr(data, "_rank_world", lambda r=r: (r, 2))
        loaders.append(data.make_loaders(_FakeTrainer(), str(tmp_path))[0])
    a = [int(i) for _ in range(2) for i in loaders[0]._next_indices()]
    b = [int(i) for _ in range(2) for i in loaders[1]._next_indices()]
    assert not set(a) & set(b) and sorted(a + b) == list(range(12))
    assert not np.array_equal(loaders[0].rng.integers(0, 12, 8), loaders[1].rng.integers(0, 12, 8))
