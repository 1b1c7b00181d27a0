"""Oracle: iBOT block-mask generation + collate (mask part) restated, and synthetic crops — test infrastructure.

`MaskGen` follows data/masking.py:14-100 statement by statement (same `random` / `numpy.random` call order, so the
same seeds give bit-identical masks — checked against the reference file itself in tests/golden/make_golden.py).
`collate_masks` follows data/collate.py:41-70.
"""
from __future__ import annotations

import math
import random

import numpy as np
import torch

from .arch import ModelCfg


class MaskGen:
    def __init__(self, input_size, num_masking_patches=None, min_num_patches=4, max_num_patches=None,
                 min_aspect=0.3, max_aspect=None):                                   # data/masking.py:15-33
        if not isinstance(input_size, tuple):
            input_size = (input_size,) * 2
        self.height, self.width = input_size
        self.num_masking_patches = num_masking_patches
        self.min_num_patches = min_num_patches
        self.max_num_patches = num_masking_patches if max_num_patches is None else max_num_patches
        max_aspect = max_aspect or 1 / min_aspect
        self.log_aspect_ratio = (math.log(min_aspect), math.log(max_aspect))

    def _mask(self, mask, max_mask_patches):                                           # data/masking.py:50-74
        delta = 0
        for _ in range(10):
            target_area = random.uniform(self.min_num_patches, max_mask_patches)
            aspect_ratio = math.exp(random.uniform(*self.log_aspect_ratio))
            h = int(round(math.sqrt(target_area * aspect_ratio)))
            w = int(round(math.sqrt(target_area / aspect_ratio)))
            if w < self.width and h < self.height:
                top = random.randint(0, self.height - h)
                left = random.randint(0, self.width - w)
                num_masked = mask[top:top + h, left:left + w].sum()
                if 0 < h * w - num_masked <= max_mask_patches:
                    for i in range(top, top + h):
                        for j in range(left, left + w):
                            if mask[i, j] == 0:
                                mask[i, j] = 1
                                delta += 1
                if delta > 0:
                    break
        return delta

    def __call__(self, num_masking_patches=0):                                         # data/masking.py:76-89
        mask = np.zeros(shape=(self.height, self.width), dtype=bool)
        mask_count = 0
        while mask_count < num_masking_patches:
            max_mask_patches = num_masking_patches - mask_count
            max_mask_patches = min(max_mask_patches, self.max_num_patches)
            delta = self._mask(mask, max_mask_patches)
            if delta == 0:
                break
            mask_count += delta
        return self.complete_mask_randomly(mask, num_masking_patches)

    def complete_mask_randomly(self, mask, num_masking_patches):                       # data/masking.py:92-100
        shape = mask.shape
        m2 = mask.flatten()
        to_add = np.random.choice(np.where(~m2)[0], size=num_masking_patches - m2.sum(), replace=False)
        m2[to_add] = True
        return m2.reshape(shape)


def make_mask_generator(cfg: ModelCfg) -> MaskGen:
    """train/train.py:778-784 (note the operator precedence of `.5 * img // patch * img // patch`)."""
    g = cfg.global_size // cfg.patch
    return MaskGen(input_size=(g, g), max_num_patches=0.5 * cfg.global_size // cfg.patch * cfg.global_size // cfg.patch)


def collate_masks(n_crops_total: int, n_tokens: int, mask_ratio_tuple, mask_probability: float, mask_generator):
    """data/collate.py:41-70: returns collated_masks [n,P] bool, mask_indices_list int64 [M], masks_weight f32 [M],
    n_masked_patches int64 [1], upperbound int."""
    B = n_crops_total
    N = n_tokens
    n_samples_masked = int(B * mask_probability)
    probs = torch.linspace(*mask_ratio_tuple, n_samples_masked + 1)
    upperbound = 0
    masks_list = []
    for i in range(0, n_samples_masked):
        prob_max = probs[i + 1]
        masks_list.append(torch.BoolTensor(mask_generator(int(N * prob_max))))
        upperbound += int(N * prob_max)
    for _ in range(n_samples_masked, B):
        masks_list.append(torch.BoolTensor(mask_generator(0)))
    random.shuffle(masks_list)
    collated_masks = torch.stack(masks_list).flatten(1)
    mask_indices_list = collated_masks.flatten().nonzero().flatten()
    masks_weight = (1 / collated_masks.sum(-1).clamp(min=1.0)).unsqueeze(-1).expand_as(collated_masks)[collated_masks]
    return {
        "collated_masks": collated_masks,
        "mask_indices_list": mask_indices_list,
        "masks_weight": masks_weight,
        "upperbound": upperbound,
        "n_masked_patches": torch.full((1,), fill_value=mask_indices_list.shape[0], dtype=torch.long),
    }


def synthetic_batch(cfg: ModelCfg, B: int, seed: int = 0, dtype=torch.bfloat16) -> dict:
    """A batch with the reference's dict contract (data/collate.py:72-93): crop-major NHWC crops cast to
    `compute_precision.param_dtype` (bf16), masks from the reference generator under random.seed/np.random.seed."""
    random.seed(seed)
    np.random.seed(seed)
    gen = torch.Generator().manual_seed(seed)
    g = torch.randn((cfg.n_global * B, cfg.global_size, cfg.global_size, 3), generator=gen).to(dtype)
    l = torch.randn((cfg.n_local * B, cfg.local_size, cfg.local_size, 3), generator=gen).to(dtype)
    out = {"collated_global_crops": g, "collated_local_crops": l}
    out.update(collate_masks(cfg.n_global * B, cfg.n_patches_global, cfg.mask_ratio, cfg.mask_probability,
                             make_mask_generator(cfg)))
    out["global_batch_size"] = B
    return out
