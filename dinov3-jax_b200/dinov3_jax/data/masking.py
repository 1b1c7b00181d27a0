"""iBOT block-mask generator with the reference's interface (dinov3_jax/data/masking.py:14-100).

Own implementation; it consumes the `random` / `numpy.random` streams in exactly the reference's order, so equal
seeds give bit-identical masks (tests/test_golden_reference.py checks this against vectors produced by the reference
file).  The draw order that has to be preserved, per rectangle attempt: area ~ U(min, budget) and log-aspect ~ U(lo, hi)
from `random.uniform`, then the top and the left corner from `random.randint` — only when the rectangle fits strictly
inside the grid; a rectangle is accepted when it adds between 1 and `budget` new cells; the remainder up to the
requested count is filled with one `numpy.random.choice` over the still-free cells (flattened, row-major).
"""
from __future__ import annotations

import math
import random

import numpy as np


class MaskingGenerator:
    def __init__(self, input_size, num_masking_patches=None, min_num_patches=4, max_num_patches=None,
                 min_aspect=0.3, max_aspect=None):
        hw = input_size if isinstance(input_size, tuple) else (input_size, input_size)
        self.height, self.width = hw
        self.num_masking_patches = num_masking_patches
        self.min_num_patches = min_num_patches
        self.max_num_patches = max_num_patches if max_num_patches is not None else num_masking_patches
        hi = max_aspect or 1 / min_aspect
        self.log_aspect_ratio = (math.log(min_aspect), math.log(hi))

    def __repr__(self):
        lo, hi = self.log_aspect_ratio
        return f"MaskingGenerator(grid={self.height}x{self.width}, block={self.min_num_patches}..{self.max_num_patches}, log_aspect=[{lo:.3f}, {hi:.3f}])"

    def get_shape(self):
        return self.height, self.width

    def _place_block(self, grid: np.ndarray, budget) -> int:
        """Up to ten attempts at dropping one rectangle; returns how many cells it newly covered (0 = none placed)."""
        H, W = self.height, self.width
        for _ in range(10):
            area = random.uniform(self.min_num_patches, budget)
            ratio = math.exp(random.uniform(*self.log_aspect_ratio))
            h, w = int(round(math.sqrt(area * ratio))), int(round(math.sqrt(area / ratio)))
            if not (w < W and h < H):
                continue
            top, left = random.randint(0, H - h), random.randint(0, W - w)
            window = grid[top:top + h, left:left + w]
            fresh = h * w - int(window.sum())
            if 0 < fresh <= budget:
                window[...] = True
                return fresh
        return 0

    def complete_mask_randomly(self, mask: np.ndarray, num_masking_patches: int) -> np.ndarray:
        flat = mask.reshape(-1).copy()
        free = np.flatnonzero(~flat)
        flat[np.random.choice(free, size=num_masking_patches - int(flat.sum()), replace=False)] = True
        return flat.reshape(mask.shape)

    def __call__(self, num_masking_patches: int = 0) -> np.ndarray:
        grid = np.zeros((self.height, self.width), dtype=bool)
        covered = 0
        while covered < num_masking_patches:
            budget = min(num_masking_patches - covered, self.max_num_patches)
            got = self._place_block(grid, budget)
            if got == 0:
                break
            covered += got
        return self.complete_mask_randomly(grid, num_masking_patches)
