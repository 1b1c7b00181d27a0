"""N-GPU FSDP step == 1-GPU step on the concatenated batch (SURVEY §8e correctness gate), product-side and oracle-free:
the sharded engine (rank-local images, all-gather per unit, reduce-scatter of the gradients, Sinkhorn sums all-reduced)
is compared with an unsharded engine that sees every rank's images at once.  KoLeo is rank-local by definition
(loss/koleo_loss.py:16-35 searches neighbours inside the device batch), so its weight is set to 0 for this comparison;
every other term of the step is identical in exact arithmetic, and both runs round at the same bf16 points.

Used by `bench.py --gpus N` (`fsdp_check` field of the JSON line) and by tests/test_real_shapes_gpu.py."""
from __future__ import annotations

import dataclasses

import torch


def _cat_crop_major(parts, n_crops):
    """[rank][n_crops*B, ...] crop-major -> [n_crops*world*B, ...] crop-major over the concatenated images."""
    B = parts[0].shape[0] // n_crops
    return torch.cat([p[c * B:(c + 1) * B] for c in range(n_crops) for p in parts], dim=0)


def fsdp_equals_single_gpu(comm, device, arch: str = "vit_small", B: int = 2, depth: int = 2, seed: int = 0,
                           hyper=None) -> dict:
    """Collective over `comm` (every rank calls it).  Returns {"loss_rel", "grad_rel", "update_rel", "world", "ok"} on
    rank 0 (other ranks get the same dict without the comparison fields filled)."""
    from ..engine import Engine, config_for
    from ..engine.synth import reference_like_params, synthetic_batch
    hyper = hyper or dict(lr=1e-3, wd=0.04, last_layer_lr=5e-4, momentum=0.99, teacher_temp=0.05)
    world, rank = comm.world, comm.rank
    cfg = dataclasses.replace(config_for(arch, global_size=64, local_size=32, n_prototypes=1024, head_hidden=256,
                                         head_bottleneck=64, layerscale=0.1, koleo_loss_weight=0.0), depth=depth)
    params = reference_like_params(cfg, seed)
    gen = torch.Generator().manual_seed(seed + 1)
    for k, v in params.items():      # reference init has zero biases / unit scales: perturb so that every term is exercised
        if k.startswith("student_") and (k.endswith("/bias") or k.endswith("/scale") or k.endswith("mask_token")):
            v.add_(torch.randn(v.shape, generator=gen) * 0.05)
            params["teacher_" + k[len("student_"):]] = v + torch.randn(v.shape, generator=gen) * 0.005
    batches = [synthetic_batch(cfg, B, seed=100 + r) for r in range(world)]
    mm = max(int(b["mask_indices_list"].shape[0]) for b in batches)
    eng = Engine(cfg, B, device=device, max_masked=max(mm, 1), comm=comm)
    eng.params.load_reference_tree(params)
    eng.set_batch(batches[rank])
    eng.forward_backward(hyper["teacher_temp"])
    eng.fsdp.finish_grads()
    g_n = {k: v.float().cpu() for k, v in eng.params.export_reference_tree("grad").items()}      # collective
    eng.optimizer_step(hyper["lr"], hyper["wd"], hyper["last_layer_lr"], hyper["momentum"])
    met = eng.read_metrics()                                                                      # collective (mean)
    p_n = {k: v.float().cpu() for k, v in eng.params.export_reference_tree("param").items()}      # collective
    out = {"world": world, "config": f"{arch} depth {depth}, {B} img/rank, 64^2/32^2 crops, K=1024, koleo weight 0",
           "grad_reduce_scatter": "push" if eng.fsdp.push else "nccl"}
    del eng
    if rank == 0:
        ng, nl = cfg.n_global, cfg.n_local
        masks = _cat_crop_major([b["collated_masks"] for b in batches], ng)
        big = {"collated_global_crops": _cat_crop_major([b["collated_global_crops"] for b in batches], ng),
               "collated_local_crops": _cat_crop_major([b["collated_local_crops"] for b in batches], nl),
               "collated_masks": masks, "mask_indices_list": masks.flatten().nonzero().flatten()}
        one = Engine(cfg, B * world, device=device, max_masked=max(int(big["mask_indices_list"].shape[0]), 1), comm=None)
        one.params.load_reference_tree(params)
        one.set_batch(big)
        one.forward_backward(hyper["teacher_temp"])
        g_1 = {k: v.float().cpu() for k, v in one.params.export_reference_tree("grad").items()}
        one.optimizer_step(hyper["lr"], hyper["wd"], hyper["last_layer_lr"], hyper["momentum"])
        m1 = one.read_metrics()
        p_1 = {k: v.float().cpu() for k, v in one.params.export_reference_tree("param").items()}
        num = sum(float(((g_n[k] - g_1[k]) ** 2).sum()) for k in g_1)
        den = sum(float((g_1[k] ** 2).sum()) for k in g_1)
        unum = sum(float(((p_n[k] - p_1[k]) ** 2).sum()) for k in p_1)
        uden = sum(float(((p_1[k] - params[k].reshape(p_1[k].shape)) ** 2).sum()) for k in p_1)
        out.update(loss_rel=abs(met["total_loss"] - m1["total_loss"]) / abs(m1["total_loss"]),
                   grad_rel=(num / max(den, 1e-30)) ** 0.5, update_rel=(unum / max(uden, 1e-30)) ** 0.5)
        # identical arithmetic up to summation order (fp32 atomics, split sums) on top of bf16 operands; step 1 of Adam
        # is lr * g / (|g| + eps), which turns noise-level gradient differences into full-size update differences for
        # the few elements whose gradient is ~0, hence the looser bound on the parameter update
        out["ok"] = bool(out["loss_rel"] < 1e-4 and out["grad_rel"] < 1e-2 and out["update_rel"] < 0.2)
        del one
    torch.cuda.empty_cache()
    return out
