"""-m gpu: the on-GPU DINO augmentation kernels (csrc/augment.cu) against torch / torchvision's float
implementations of the same transforms with the same parameters, and the batch pipeline -> engine hand-off."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _records(n):
    from dinov3_jax.data.gpu_augment import CROP_DTYPE
    r = np.zeros(n, dtype=CROP_DTYPE)
    r["order"] = -1
    return r


def _aug():
    from dinov3_jax.data.gpu_augment import GpuDataAugmentationDINO
    return GpuDataAugmentationDINO((0.32, 1.0), (0.05, 0.32), 8, global_crops_size=64, local_crops_size=32, seed=0)


def _run(stage, x, rec, sig=None, S=None):
    """Drive one kernel stage on fp32 [n,S,S,3] input (or uint8 images for the crop stage)."""
    from dinov3_jax import _native as N
    lib = N.init()
    dev = x.device
    d_rec = torch.from_numpy(rec.view(np.uint8).reshape(-1).copy()).to(dev)
    s = N.stream_ptr()
    n = rec.shape[0]
    if stage == "crop":
        B, H, W, _ = x.shape
        out = torch.empty(n, S, S, 3, device=dev)
        N.check(lib.d3_aug_resized_crop(N.ptr(x), B, H, W, N.ptr(d_rec), n, N.ptr(out), S, s), "crop")
        return out
    S = x.shape[1]
    if stage == "color":
        y = x.clone(); g = torch.zeros(n, device=dev)
        N.check(lib.d3_aug_color(N.ptr(y), N.ptr(d_rec), n, S, N.ptr(g), s), "color")
        return y
    if stage == "blur":
        t, y = torch.empty_like(x), torch.empty_like(x)
        N.check(lib.d3_aug_blur(N.ptr(x), N.ptr(t), N.ptr(y), N.ptr(torch.from_numpy(sig).to(dev)), n, S, s), "blur")
        return y
    if stage == "finish":
        import ctypes as C
        out = torch.empty(n, S, S, 3, device=dev, dtype=torch.bfloat16)
        mean, std = (C.c_float * 3)(0.485, 0.456, 0.406), (C.c_float * 3)(0.229, 0.224, 0.225)
        N.check(lib.d3_aug_finish(N.ptr(x), N.ptr(out), N.ptr(d_rec), n, S, mean, std, s), "finish")
        return out


def test_resized_crop_matches_antialiased_bicubic():
    g = torch.Generator().manual_seed(0)
    imgs = torch.randint(0, 256, (3, 150, 210, 3), generator=g, dtype=torch.uint8)
    boxes = [(0, 10, 20, 120, 100, 0), (1, 0, 0, 210, 150, 1), (2, 50, 40, 30, 24, 0), (0, 5, 5, 64, 64, 1), (1, 100, 30, 97, 119, 0)]
    rec = _records(len(boxes))
    for i, (img, x0, y0, w, h, flip) in enumerate(boxes):
        rec["img"][i], rec["x0"][i], rec["y0"][i], rec["w"][i], rec["h"][i], rec["flip"][i] = img, x0, y0, w, h, flip
    for S in (64, 32):
        got = _run("crop", imgs.cuda(), rec, S=S).cpu()
        for i, (img, x0, y0, w, h, flip) in enumerate(boxes):
            crop = imgs[img, y0:y0 + h, x0:x0 + w].permute(2, 0, 1).float()[None] / 255.0
            ref = torch.nn.functional.interpolate(crop, size=(S, S), mode="bicubic", antialias=True, align_corners=False)[0]
            ref = ref.clamp(0, 1).permute(1, 2, 0)
            if flip:
                ref = ref.flip(1)
            assert float((got[i] - ref).abs().max()) < 2e-3, (S, i, float((got[i] - ref).abs().max()))


def test_color_jitter_and_grayscale_match_torchvision():
    import itertools
    from torchvision.transforms.v2 import functional as F
    g = torch.Generator().manual_seed(1)
    S = 32
    orders = list(itertools.permutations(range(4)))[::3] + [(-1, -1, -1, -1)]
    n = len(orders)
    x = torch.rand(n, S, S, 3, generator=g)
    rec = _records(n)
    rng = np.random.default_rng(0)
    for i, o in enumerate(orders):
        rec["order"][i] = o
        rec["fb"][i], rec["fc"][i], rec["fs"][i], rec["fh"][i] = rng.uniform(0.6, 1.4), rng.uniform(0.6, 1.4), rng.uniform(0.8, 1.2), rng.uniform(-0.1, 0.1)
        rec["gray"][i] = i % 3 == 0
    got = _run("color", x.cuda(), rec).cpu()
    for i, o in enumerate(orders):
        img = x[i].permute(2, 0, 1)
        if o[0] >= 0:
            for op in o:
                img = (F.adjust_brightness(img, float(rec["fb"][i])) if op == 0 else F.adjust_contrast(img, float(rec["fc"][i])) if op == 1
                       else F.adjust_saturation(img, float(rec["fs"][i])) if op == 2 else F.adjust_hue(img, float(rec["fh"][i])))
        if rec["gray"][i]:
            img = F.rgb_to_grayscale(img, num_output_channels=3)
        assert float((got[i] - img.permute(1, 2, 0)).abs().max()) < 2e-5, (i, o, float((got[i] - img.permute(1, 2, 0)).abs().max()))


def test_gaussian_blur_solarize_normalize_match_torchvision():
    from torchvision.transforms.v2 import functional as F
    g = torch.Generator().manual_seed(2)
    S, n = 48, 4
    x = torch.rand(n, S, S, 3, generator=g)
    sig = np.array([0.0, 0.1, 1.0, 2.0], dtype=np.float32)
    rec = _records(n)
    got = _run("blur", x.cuda(), rec, sig=sig).cpu()
    for i in range(n):
        img = x[i].permute(2, 0, 1)
        ref = img if sig[i] <= 0 else F.gaussian_blur(img, kernel_size=[9, 9], sigma=[float(sig[i])] * 2)
        assert float((got[i] - ref.permute(1, 2, 0)).abs().max()) < 2e-6, i
    rec["solarize"] = [0, 1, 0, 1]
    out = _run("finish", x.cuda(), rec).float().cpu()
    mean, std = torch.tensor([0.485, 0.456, 0.406]), torch.tensor([0.229, 0.224, 0.225])
    for i in range(n):
        img = x[i]
        if rec["solarize"][i]:
            img = torch.where(img >= 128 / 255, 1 - img, img)
        ref = ((img - mean) / std).to(torch.bfloat16).float()
        assert torch.equal(out[i], ref) or float((out[i] - ref).abs().max()) < 2e-2     # one bf16 ulp where fp32 rounding differs


def test_gpu_batch_pipeline_feeds_the_engine():
    """uint8 images in HBM -> GpuBatchPipeline -> Engine.train_step: the collate contract (keys, layouts, dtypes, crop-major
    order) and a finite loss."""
    from dinov3_jax.configs import DinoV3SetupArgs, setup_config
    from dinov3_jax.data.gpu_augment import GpuBatchPipeline
    from dinov3_jax.engine.synth import init_reference_like
    from dinov3_jax.train.ssl_meta_arch import SSLMetaArch
    opts = ["train.batch_size_per_gpu=4", "student.arch=vit_small", "crops.global_crops_size=64", "crops.local_crops_size=32",
            "dino.head_n_prototypes=512", "ibot.head_n_prototypes=512", "dino.head_hidden_dim=256", "ibot.head_hidden_dim=256",
            "dino.head_bottleneck_dim=64", "ibot.head_bottleneck_dim=64"]
    config = setup_config(DinoV3SetupArgs(opts=opts))
    pipe = GpuBatchPipeline(config, seed=3)
    g = torch.Generator().manual_seed(0)
    noise = torch.randn(4, 224, 224, 3, generator=g)                     # the reference decoder's image distribution
    imgs = ((noise - noise.amin((1, 2, 3), keepdim=True)) / (noise.amax((1, 2, 3), keepdim=True) - noise.amin((1, 2, 3), keepdim=True)) * 255).to(torch.uint8).cuda()
    batch = pipe(imgs)
    assert batch["collated_global_crops"].shape == (8, 64, 64, 3) and batch["collated_global_crops"].dtype == torch.bfloat16
    assert batch["collated_local_crops"].shape == (32, 32, 32, 3) and batch["collated_masks"].shape == (8, 16)
    gc = batch["collated_global_crops"].float()
    assert torch.isfinite(gc).all() and 0.2 < float(gc.std()) < 3.0
    # same image -> its two global crops sit B apart (crop-major)
    model = SSLMetaArch(config)
    eng = model.build_engine(max_masked=max(int(batch["mask_indices_list"].shape[0]), 1))
    init_reference_like(eng, seed=0)
    eng.train_step(batch, teacher_temp=0.05, lr=1e-3, wd=0.04, last_layer_lr=5e-4, momentum=0.99)
    m = eng.read_metrics()
    assert m["total_loss"] == m["total_loss"] and m["total_loss"] > 0
