"""Regenerate tests/golden/*.npz by EXECUTING reference files from /root/reference (build container only).

  * data/masking.py, train/cosine_lr_scheduler.py      — imported as they are (pure numpy / random)
  * data/collate.py mask section                       — re-executed with the reference MaskingGenerator
  * loss/dino_clstoken_loss.py, loss/ibot_patch_loss.py, loss/koleo_loss.py, layers/rope_position_encoding.py,
    layers/attention.py (rope_rotate_half / rope_apply), train/param_groups.py
                                                       — imported UNMODIFIED under oracle.jaxshim (numpy stand-in for
                                                         jax / flax.linen; the real stack is not installable offline)
Usage:  python tests/golden/make_golden.py      (writes next to this file; the .npz files are committed)
"""
from __future__ import annotations

import importlib.util
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/dinov3_jax"
sys.path.insert(0, ROOT)


def load(name, rel):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    assert os.path.isdir(REF), "reference checkout not present: golden vectors can only be regenerated in the build container"
    import torch
    from oracle.jaxshim import install, Arr
    install()
    J = lambda a: np.array(a, copy=True).view(Arr)   # inputs enter the reference code as (stand-in) jax arrays
    out = {}

    # ---- masks (reference generator + the mask section of data/collate.py:41-70)
    masking = load("ref_masking", "data/masking.py")
    for tag, (n, grid, size) in {"b8": (16, 14, 224), "tiny": (8, 4, 64)}.items():
        random.seed(7); np.random.seed(7)
        gen = masking.MaskingGenerator(input_size=(grid, grid), max_num_patches=0.5 * size // 16 * size // 16)
        N = grid * grid
        n_masked = int(n * 0.5)
        probs = torch.linspace(0.1, 0.5, n_masked + 1)
        masks, upper = [], 0
        for i in range(n_masked):
            masks.append(torch.BoolTensor(gen(int(N * probs[i + 1])))); upper += int(N * probs[i + 1])
        for _ in range(n_masked, n):
            masks.append(torch.BoolTensor(gen(0)))
        random.shuffle(masks)
        cm = torch.stack(masks).flatten(1)
        out[f"masks_{tag}"] = cm.numpy()
        out[f"mask_indices_{tag}"] = cm.flatten().nonzero().flatten().numpy()
        out[f"masks_weight_{tag}"] = (1 / cm.sum(-1).clamp(min=1.0)).unsqueeze(-1).expand_as(cm)[cm].numpy()
        out[f"upperbound_{tag}"] = np.array(upper)

    # ---- schedules
    sched = load("ref_sched", "train/cosine_lr_scheduler.py")
    out["sched_lr"] = sched.CosineScheduler(base_value=1e-3, final_value=1e-6, total_iters=500, warmup_iters=50, start_warmup_value=0).gen()
    out["sched_wd"] = sched.CosineScheduler(base_value=0.04, final_value=0.4, total_iters=500).gen()
    out["sched_temp"] = sched.CosineScheduler(base_value=0.07, final_value=0.07, total_iters=120, warmup_iters=120, start_warmup_value=0.04).gen()
    out["sched_freeze"] = sched.CosineScheduler(base_value=1.0, final_value=0.0, total_iters=60, warmup_iters=10, freeze_iters=5).gen()

    # ---- losses / rope under the shim
    rng = np.random.default_rng(0)
    dino = load("ref_dino", "loss/dino_clstoken_loss.py")
    ibot = load("ref_ibot", "loss/ibot_patch_loss.py")
    koleo = load("ref_koleo", "loss/koleo_loss.py")
    rope = load("ref_rope", "layers/rope_position_encoding.py")
    attn = load("ref_attn", "layers/attention.py")
    K, B = 48, 5
    t_logits = rng.standard_normal((2 * B, K)) * 0.3
    s_global = rng.standard_normal((2, B, K))
    s_local = rng.standard_normal((8, B, K))
    dl = dino.DINOLoss(K)
    probs = dl.sinkhorn_knopp_teacher(J(t_logits), teacher_temp=0.05)
    out.update(dino_t_logits=t_logits, dino_s_global=s_global, dino_s_local=s_local, dino_probs=np.asarray(probs),
               dino_loss_local=np.asarray(dl(J(s_local), J(probs).reshape(2, B, K))),
               dino_loss_global=np.asarray(dl(J(s_global), J(probs).reshape(2, B, K), ignore_diagonal=True)))
    M = 11
    p_logits = rng.standard_normal((M, K)) * 0.3
    s_patch = rng.standard_normal((M, K))
    il = ibot.iBOTPatchLoss(K)
    p_probs = il.sinkhorn_knopp_teacher(J(p_logits), teacher_temp=0.05, n_masked_patches_tensor=J(np.array([M])))
    masks_flat = np.zeros((2 * B, 16), dtype=bool)
    out.update(ibot_t_logits=p_logits, ibot_s=s_patch, ibot_probs=np.asarray(p_probs),
               ibot_loss=np.asarray(il.forward_masked(J(s_patch), J(p_probs), student_masks_flat=J(masks_flat),
                                                      n_masked_patches=M, masks_weight=np.ones(M))))
    x = rng.standard_normal((7, 32))
    out.update(koleo_x=x, koleo_loss=np.asarray(koleo.KoLeoLoss()(J(x))))
    for (H, W) in ((14, 14), (6, 6), (3, 5)):
        r = rope.RopePositionEmbedding(embed_dim=384, num_heads=6)
        sin, cos = r(H=H, W=W)
        out[f"rope_sin_{H}x{W}"], out[f"rope_cos_{H}x{W}"] = np.asarray(sin), np.asarray(cos)
    xr = rng.standard_normal((2, 3, 9, 64))
    sin, cos = rope.RopePositionEmbedding(embed_dim=128, num_heads=2)(H=3, W=3)
    out.update(rope_x=xr, rope_y=np.asarray(attn.rope_apply(J(xr), J(sin), J(cos))))

    # ---- param groups (layer-wise decay etc.)
    pg = load("ref_pg", "train/param_groups.py")
    depth = 4
    tree = {"patch_embed": {"proj": {"kernel": 0, "bias": 0}}, "cls_token": 0, "mask_token": 0, "norm": {"scale": 0, "bias": 0}}
    for i in range(depth):
        tree[f"blocks_{i}"] = {"norm1": {"scale": 0, "bias": 0}, "attn": {"qkv": {"kernel": 0, "bias": 0}, "proj": {"kernel": 0, "bias": 0}},
                               "ls1": {"gamma": 0}, "norm2": {"scale": 0, "bias": 0},
                               "mlp": {"Dense_0": {"kernel": 0, "bias": 0}, "Dense_1": {"kernel": 0, "bias": 0}}, "ls2": {"gamma": 0}}
    head = {"mlp": {f"layers_{i}": {"kernel": 0, "bias": 0} for i in (0, 2, 4)}, "last_layer": {"kernel": 0}}
    names, vals = [], []
    from flax.traverse_util import flatten_dict
    for root, t in (("student_backbone", tree), ("student_dino_head", head), ("student_ibot_head", head)):
        g = pg.get_params_groups_with_decay_fsdp(t, lr_decay_rate=0.9, patch_embed_lr_mult=0.2, dino_head_wd_multiplier=1.0, root_name=root)
        for k, v in flatten_dict(g, sep="/").items():
            names.append(f"{root}/{k}"); vals.append((v.lr_multiplier, v.wd_multiplier, float(v.is_last_layer)))
    out["pg_names"] = np.array(names)
    out["pg_values"] = np.array(vals, dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "reference_vectors.npz"), **out)
    print("wrote", os.path.join(HERE, "reference_vectors.npz"), len(out), "arrays")


if __name__ == "__main__":
    main()
