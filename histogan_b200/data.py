"""Minimal image-folder data source for Trainer.set_data_src
(histoGAN/histoGAN.py:253-307,827-851).

The reference computes each sample's target histogram in DataLoader worker
processes with a CPU RGBuvHistBlock on two random images (:296-302), 0.06-0.4 s per
sample.  Here the images are decoded on the host and their target histograms come
from the CUDA histogram block on the device: one call per source image
(the images keep their own sizes, as in the reference), ~0.1 ms each.
JPEG decoding / augmentation is host-side I/O and out of scope for the kernels.
"""
from __future__ import annotations

from pathlib import Path

import numpy as np
import torch

EXTS = ['jpg', 'png']


class _FolderBatches:
    def __init__(self, trainer, folder, batch_size, image_size, with_images):
        from PIL import Image  # noqa: F401
        self.paths = [p for ext in EXTS for p in Path(f'{folder}').glob(f'**/*.{ext}')]
        if not self.paths:
            raise RuntimeError(f'no images found under {folder}')
        self.trainer = trainer
        self.batch_size = batch_size
        self.image_size = image_size
        self.with_images = with_images
        self.rng = np.random.default_rng(0)
        # where the histogram block runs (its constructor normalised 'cuda' / int / 'cuda:N')
        self.device = getattr(trainer.histBlock, 'device', 'cuda')

    def _load(self, path, size=None):
        from PIL import Image
        img = Image.open(path).convert('RGB')
        if size is not None:
            w, h = img.size
            s = size / min(w, h)
            img = img.resize((max(size, round(w * s)), max(size, round(h * s))), Image.BILINEAR)
            w, h = img.size
            l, t = (w - size) // 2, (h - size) // 2
            img = img.crop((l, t, l + size, t + size))
        return torch.from_numpy(np.asarray(img, dtype=np.float32) / 255.0).permute(2, 0, 1)

    def __iter__(self):
        return self

    def __next__(self):
        blk = self.trainer.histBlock
        n = len(self.paths)
        out = {}
        if self.with_images:
            idx = self.rng.integers(0, n, size=self.batch_size)
            out['images'] = torch.stack([self._load(self.paths[i], self.image_size) for i in idx])
        hists = []
        with torch.no_grad():
            for _ in range(self.batch_size):
                i1, i2 = self.rng.integers(0, n, size=2)
                h1 = blk(self._load(self.paths[i1]).unsqueeze(0).to(self.device))
                if self.with_images:        # random convex mix of two histograms (:179-181)
                    h2 = blk(self._load(self.paths[i2]).unsqueeze(0).to(self.device))
                    r = float(torch.rand(1))
                    h1 = h1 * r + h2 * (1 - r)
                hists.append(h1.squeeze(0))
        out['histograms'] = torch.stack(hists)
        return out


def make_loaders(trainer, folder):
    return (_FolderBatches(trainer, folder, trainer.batch_size, trainer.image_size, True),
            _FolderBatches(trainer, folder, 4, 150, False))
