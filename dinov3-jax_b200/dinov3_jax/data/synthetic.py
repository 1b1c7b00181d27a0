"""Synthetic image source with the distribution of the reference's own decoder (data/datasets/decoders.py:31-34 returns
min-max-scaled Gaussian noise as a 224x224 uint8 RGB PIL image instead of decoding the file), as a map-style dataset:
`train.dataset_path=synthetic:noise` runs the full host pipeline (DINO augmentation -> collate_data_and_cast -> engine)
without a dataset on disk or the reference's jax-dependent samplers."""
from __future__ import annotations

import numpy as np
import torch


class NoiseImageDataset(torch.utils.data.Dataset):
    def __init__(self, length: int = 1 << 20, size: int = 224, transform=None, target_transform=None, seed: int = 0):
        self.length, self.size, self.transform, self.target_transform, self.seed = length, size, transform, target_transform, seed

    def __len__(self):
        return self.length

    def __getitem__(self, index: int):
        from PIL import Image
        rng = np.random.default_rng(self.seed * 1000003 + index)
        img = rng.standard_normal((self.size, self.size, 3))
        img = (img - img.min()) / (img.max() - img.min())
        image = Image.fromarray((img * 255).astype(np.uint8))
        target = ()
        if self.transform is not None:
            image = self.transform(image)
        if self.target_transform is not None:
            target = self.target_transform(target)
        return image, target


class SeededBatchSampler(torch.utils.data.Sampler):
    """Infinite shuffled index stream, rank-strided, resumable: `advance` skips the samples consumed before a resume
    (the reference passes sampler_advance = start_iter * batch, train/train.py:843)."""

    def __init__(self, n: int, batch_size: int, seed: int, rank: int = 0, world: int = 1, advance: int = 0):
        self.n, self.batch_size, self.seed, self.rank, self.world, self.advance = n, batch_size, seed, rank, world, advance

    def __iter__(self):
        epoch, skip = 0, self.advance
        while True:
            g = torch.Generator().manual_seed(self.seed + epoch)
            perm = torch.randperm(self.n, generator=g)[self.rank::self.world].tolist()
            usable = len(perm) - len(perm) % self.batch_size
            if skip >= usable:
                skip -= usable
            else:
                for i in range(skip, usable, self.batch_size):
                    yield perm[i:i + self.batch_size]
                skip = 0
            epoch += 1
