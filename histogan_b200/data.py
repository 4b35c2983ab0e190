"""Minimal image-folder data source for Trainer.set_data_src
(histoGAN/histoGAN.py:253-307,827-851).

The reference computes each sample's target histogram in DataLoader worker
processes with a CPU RGBuvHistBlock on two random images (:296-302), 0.06-0.4 s per
sample.  Here the images are decoded on the host and the target histograms of a whole
batch come from the CUDA histogram block on the device: the 2 x batch
source images are grouped by size and every group is ONE batched call of the block
(a dataset of equally sized images = one call per batch).
Iteration mirrors the reference's ``DataLoader(shuffle=True, drop_last=True)``: a fresh
permutation per epoch, without replacement; under torch.distributed the permutation is
shared and rank r takes every world-th index, while the random histogram sources and mix
ratios are drawn from a per-rank generator.
JPEG decoding / augmentation is host-side I/O and out of scope for the kernels.
"""
from __future__ import annotations

from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

EXTS = ['jpg', 'png']


def _rank_world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


class _FolderBatches:
    def __init__(self, trainer, folder, batch_size, image_size, with_images, seed=0):
        from PIL import Image  # noqa: F401
        self.paths = sorted(p for ext in EXTS for p in Path(f'{folder}').glob(f'**/*.{ext}'))
        if not self.paths:
            raise RuntimeError(f'no images found under {folder}')
        self.trainer = trainer
        self.batch_size = batch_size
        self.image_size = image_size
        self.with_images = with_images
        self.rank, self.world = _rank_world()
        self.seed = seed
        self.rng = np.random.default_rng([seed, self.rank])       # per-rank draws
        self.epoch = 0
        self._order = []
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

    def _next_indices(self):
        """the next `batch_size` indices of this rank's share of the epoch permutation"""
        n = len(self.paths)
        while len(self._order) < self.batch_size:
            perm = np.random.default_rng([self.seed, 1 << 20, self.epoch]).permutation(n)   # same on all ranks
            self.epoch += 1
            share = perm[self.rank::self.world]
            if n >= self.batch_size * self.world:          # drop_last, as the reference's loader
                share = share[:len(share) // self.batch_size * self.batch_size]
            self._order = list(self._order) + list(share)
        idx, self._order = self._order[:self.batch_size], self._order[self.batch_size:]
        return idx

    def _histograms(self, sources):
        """histBlock of every source image: one batched call per distinct image size"""
        blk = self.trainer.histBlock
        imgs = [self._load(self.paths[i]) for i in sources]
        out = [None] * len(imgs)
        groups = {}
        for k, im in enumerate(imgs):
            groups.setdefault(tuple(im.shape), []).append(k)
        with torch.no_grad():
            for ks in groups.values():
                h = blk(torch.stack([imgs[k] for k in ks]).to(self.device))
                for k, hk in zip(ks, h):
                    out[k] = hk
        return torch.stack(out)

    def __iter__(self):
        return self

    def __next__(self):
        n = len(self.paths)
        out = {}
        idx = self._next_indices()
        if self.with_images:
            out['images'] = torch.stack([self._load(self.paths[i], self.image_size) for i in idx])
            src = self.rng.integers(0, n, size=(self.batch_size, 2))       # two random images (:294)
            h = self._histograms(src.reshape(-1)).reshape(self.batch_size, 2, *self._hshape())
            r = torch.rand(self.batch_size, 1, 1, 1).to(h.device)          # hist_interpolation (:179-181)
            out['histograms'] = h[:, 0] * r + h[:, 1] * (1 - r)
        else:
            out['histograms'] = self._histograms(idx)                       # test=True branch (:303-307)
        return out

    def _hshape(self):
        blk = self.trainer.histBlock
        h = getattr(blk, 'h', 64)
        return (1 if getattr(blk, 'green_only', False) else 3, h, h)


def make_loaders(trainer, folder):
    return (_FolderBatches(trainer, folder, trainer.batch_size, trainer.image_size, True),
            _FolderBatches(trainer, folder, 4, 150, False))
