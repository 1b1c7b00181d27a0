"""On-GPU DINO multi-crop augmentation and batch assembly: the step before the training hot path.

`GpuDataAugmentationDINO` has the constructor of the reference's `DataAugmentationDINO`
(dinov3_jax/data/augmentations.py:23-56) but works on a whole BATCH of decoded uint8 images that already sit in HBM
([B, H, W, 3]) and returns the collated crop tensors of `collate_data_and_cast` (data/collate.py:72-93: crop-major NHWC
in `param_dtype`) directly — no per-sample PIL objects, no host loop over pixels.  Only the random PARAMETERS are drawn
on the host (a few dozen scalars per image, numpy), following torchvision's sampling rules:

  RandomResizedCrop.get_params   10 attempts of area ~ U(scale) * H * W, log-ratio ~ U(log 3/4, log 4/3), integer box,
                                 centre-crop fallback; interpolation bicubic with antialias (what PIL does when it shrinks)
  RandomHorizontalFlip           p = 0.5 (0 when horizontal_flips is false)
  ColorJitter(0.4, 0.4, 0.2, 0.1) applied with p = 0.8, ops in a random order; RandomGrayscale p = 0.2
  GaussianBlur(kernel 9, sigma ~ U(0.1, 2))  — the reference wraps it as RandomApply(p = 1 - p_arg)
                                 (data/transforms.py:30-33), so the blur probabilities AS CODED are 0.0 / 0.9 / 0.5 for
                                 global crop 1 / global crop 2 / local crops; that is what is followed here
  RandomSolarize(threshold 128)  p = 0.2, second global crop only
  ToTensor + Normalize(mean, std)

The kernels (csrc/augment.cu) are deterministic functions of those parameters and are tested against torchvision's
float implementations.  `GpuBatchPipeline` adds the iBOT block masks (host `MaskingGenerator`, as in the reference's
collate) and yields the batch dict `Engine.set_batch` consumes.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch

from .. import _native as N
from .collate import collate_masks
from .masking import MaskingGenerator

CROP_DTYPE = np.dtype([("img", "<i4"), ("x0", "<i4"), ("y0", "<i4"), ("w", "<i4"), ("h", "<i4"), ("flip", "<i4"),
                       ("order", "<i4", (4,)), ("fb", "<f4"), ("fc", "<f4"), ("fs", "<f4"), ("fh", "<f4"),
                       ("gray", "<i4"), ("solarize", "<i4")])
assert CROP_DTYPE.itemsize == 64

IMAGENET_DEFAULT_MEAN = (0.485, 0.456, 0.406)
IMAGENET_DEFAULT_STD = (0.229, 0.224, 0.225)


def _resized_crop_params(rng, H, W, scale, ratio=(3.0 / 4.0, 4.0 / 3.0)):
    """torchvision RandomResizedCrop.get_params -> (y0, x0, h, w)."""
    area = H * W
    log_ratio = (math.log(ratio[0]), math.log(ratio[1]))
    for _ in range(10):
        target = area * rng.uniform(scale[0], scale[1])
        ar = math.exp(rng.uniform(log_ratio[0], log_ratio[1]))
        w = int(round(math.sqrt(target * ar)))
        h = int(round(math.sqrt(target / ar)))
        if 0 < w <= W and 0 < h <= H:
            return int(rng.integers(0, H - h + 1)), int(rng.integers(0, W - w + 1)), h, w
    in_ratio = W / H
    if in_ratio < ratio[0]:
        w, h = W, int(round(W / ratio[0]))
    elif in_ratio > ratio[1]:
        h, w = H, int(round(H * ratio[1]))
    else:
        w, h = W, H
    return (H - h) // 2, (W - w) // 2, h, w


class GpuDataAugmentationDINO:
    def __init__(self, global_crops_scale, local_crops_scale, local_crops_number, global_crops_size=224,
                 local_crops_size=96, gram_teacher_crops_size=None, gram_teacher_no_distortions=False,
                 teacher_no_color_jitter=False, local_crops_subset_of_global_crops=False, patch_size=16,
                 share_color_jitter=False, horizontal_flips=True, mean=IMAGENET_DEFAULT_MEAN, std=IMAGENET_DEFAULT_STD,
                 seed: int = 0, out_dtype=torch.bfloat16):
        if gram_teacher_crops_size is not None or local_crops_subset_of_global_crops or share_color_jitter or teacher_no_color_jitter:
            raise NotImplementedError("gram-teacher crops / local-subset crops / shared or teacher-free colour jitter are not "
                                      "on the GPU augmentation path (the reference defaults are)")
        assert out_dtype == torch.bfloat16, "the kernels emit bf16 (compute_precision.param_dtype: bf16)"
        self.global_scale, self.local_scale = tuple(global_crops_scale), tuple(local_crops_scale)
        self.n_local, self.gs, self.ls = int(local_crops_number), int(global_crops_size), int(local_crops_size)
        self.flip_p = 0.5 if horizontal_flips else 0.0
        self.mean = (C.c_float * 3)(*[float(v) for v in mean])
        self.std = (C.c_float * 3)(*[float(v) for v in std])
        self.rng = np.random.default_rng(seed)
        self._scratch = {}

    # ---- host: random parameters ----------------------------------------------------------------------------------
    def sample(self, B: int, H: int, W: int):
        """Crop records (crop-major: crop index outer, image inner, like the collate's stacking) and blur sigmas for the
        2 global and n_local local crop sets."""
        rng = self.rng
        g = np.zeros(2 * B, dtype=CROP_DTYPE)
        l = np.zeros(self.n_local * B, dtype=CROP_DTYPE)
        gb = np.zeros(2 * B, dtype=np.float32)
        lb = np.zeros(self.n_local * B, dtype=np.float32)

        def fill(rec, i, img, scale, blur_apply_p, solarize_p, sig):
            y0, x0, h, w = _resized_crop_params(rng, H, W, scale)
            rec["img"][i], rec["x0"][i], rec["y0"][i], rec["w"][i], rec["h"][i] = img, x0, y0, w, h
            rec["flip"][i] = int(rng.random() < self.flip_p)
            if rng.random() < 0.8:                              # RandomApply([ColorJitter], p=0.8)
                rec["order"][i] = rng.permutation(4)
                rec["fb"][i], rec["fc"][i] = rng.uniform(0.6, 1.4), rng.uniform(0.6, 1.4)
                rec["fs"][i], rec["fh"][i] = rng.uniform(0.8, 1.2), rng.uniform(-0.1, 0.1)
            else:
                rec["order"][i] = -1
            rec["gray"][i] = int(rng.random() < 0.2)            # RandomGrayscale(p=0.2)
            sig[i] = rng.uniform(0.1, 2.0) if rng.random() < blur_apply_p else 0.0
            rec["solarize"][i] = int(rng.random() < solarize_p)

        for b in range(B):
            # reference GaussianBlur(p=...) applies the blur with probability 1 - p (data/transforms.py:30-33)
            fill(g, 0 * B + b, b, self.global_scale, 1.0 - 1.0, 0.0, gb)          # global_transfo1: GaussianBlur(p=1.0)
            fill(g, 1 * B + b, b, self.global_scale, 1.0 - 0.1, 0.2, gb)          # global_transfo2: GaussianBlur(p=0.1), Solarize(0.2)
            for c in range(self.n_local):
                fill(l, c * B + b, b, self.local_scale, 1.0 - 0.5, 0.0, lb)       # local_transfo: GaussianBlur(p=0.5)
        return (g, gb), (l, lb)

    # ---- device: kernels ----------------------------------------------------------------------------------------------
    def _buf(self, key, shape, dtype, device):
        t = self._scratch.get(key)
        if t is None or t.shape != tuple(shape) or t.device != device:
            t = torch.empty(shape, dtype=dtype, device=device)
            self._scratch[key] = t
        return t

    def apply(self, images_u8: torch.Tensor, crops: np.ndarray, sigmas: np.ndarray, S: int) -> torch.Tensor:
        """images_u8 [B,H,W,3] uint8 on the GPU; returns [n_crops, S, S, 3] bf16 (normalised)."""
        assert images_u8.dtype == torch.uint8 and images_u8.is_cuda and images_u8.is_contiguous() and images_u8.shape[-1] == 3
        lib = N.init()
        dev = images_u8.device
        B, H, W, _ = images_u8.shape
        n = crops.shape[0]
        d_crops = torch.from_numpy(crops.view(np.uint8).reshape(-1).copy()).to(dev, non_blocking=True)
        d_sig = torch.from_numpy(sigmas.astype(np.float32)).to(dev, non_blocking=True)
        x = self._buf(("x", S), (n, S, S, 3), torch.float32, dev)
        t1 = self._buf(("t1", S), (n, S, S, 3), torch.float32, dev)
        t2 = self._buf(("t2", S), (n, S, S, 3), torch.float32, dev)
        gsum = torch.zeros(n, dtype=torch.float32, device=dev)
        out = torch.empty(n, S, S, 3, dtype=torch.bfloat16, device=dev)
        s = N.stream_ptr()
        N.check(lib.d3_aug_resized_crop(N.ptr(images_u8), B, H, W, N.ptr(d_crops), n, N.ptr(x), S, s), "d3_aug_resized_crop")
        N.check(lib.d3_aug_color(N.ptr(x), N.ptr(d_crops), n, S, N.ptr(gsum), s), "d3_aug_color")
        N.check(lib.d3_aug_blur(N.ptr(x), N.ptr(t1), N.ptr(t2), N.ptr(d_sig), n, S, s), "d3_aug_blur")
        N.check(lib.d3_aug_finish(N.ptr(t2), N.ptr(out), N.ptr(d_crops), n, S, self.mean, self.std, s), "d3_aug_finish")
        return out

    def __call__(self, images_u8: torch.Tensor) -> dict:
        B, H, W, _ = images_u8.shape
        (g, gb), (l, lb) = self.sample(B, H, W)
        return {"collated_global_crops": self.apply(images_u8, g, gb, self.gs),
                "collated_local_crops": self.apply(images_u8, l, lb, self.ls)}


class GpuBatchPipeline:
    """uint8 image batch on the GPU -> the batch dict of data/collate.py:72-93 (crops from the kernels above, iBOT block
    masks from the reference's host generator as in its collate).  `config` is the reference-shaped config."""

    def __init__(self, config, seed: int = 0):
        c = config.crops
        self.aug = GpuDataAugmentationDINO(c.global_crops_scale, c.local_crops_scale, c.local_crops_number,
                                           global_crops_size=c.global_crops_size, local_crops_size=c.local_crops_size,
                                           horizontal_flips=c.get("horizontal_flips", True),
                                           mean=c.get("rgb_mean", IMAGENET_DEFAULT_MEAN), std=c.get("rgb_std", IMAGENET_DEFAULT_STD),
                                           patch_size=config.student.patch_size, seed=seed)
        grid = c.global_crops_size // config.student.patch_size
        self.n_tokens = grid * grid
        self.mask_generator = MaskingGenerator(input_size=(grid, grid),
                                               max_num_patches=0.5 * c.global_crops_size // config.student.patch_size
                                               * c.global_crops_size // config.student.patch_size)
        self.mask_ratio = tuple(config.ibot.mask_ratio_min_max)
        self.mask_probability = config.ibot.mask_sample_probability
        self.circular = bool(config.ibot.get("mask_random_circular_shift", False))

    def __call__(self, images_u8: torch.Tensor) -> dict:
        out = self.aug(images_u8)
        n_global = out["collated_global_crops"].shape[0]
        out.update(collate_masks(n_global, self.n_tokens, self.mask_ratio, self.mask_probability, self.mask_generator,
                                 self.circular))
        out["global_batch_size"] = images_u8.shape[0]
        return out
