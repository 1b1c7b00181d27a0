"""Oracle: SSL meta-arch forward, loss assembly, gradient, clip, AdamW, teacher EMA, schedules — test infrastructure."""
from __future__ import annotations

import numpy as np
import torch

from .arch import ModelCfg
from .losses import dino_loss, gram_loss, ibot_loss_masked, koleo_loss, sinkhorn_knopp
from .model import Emu, backbone_forward, head_forward, sub

STUDENT_MODULES = ("student_backbone", "student_dino_head", "student_ibot_head")


def ssl_forward(params: dict, batch: dict, teacher_temp: float, cfg: ModelCfg, emu: Emu = Emu(False),
                world: int = 1, allreduce=None, dtype=torch.float32, return_aux: bool = False, centers=None,
                gram: dict | None = None):
    """train/ssl_meta_arch.py:289-363 (__call__), :366-402 (teacher), :406-460 (student), :463-557 (losses).

    Single-device semantics: `batch` holds this rank's B images, crop-major.

    `gram`: {"weight", "ema_teacher", "normalized", "img_level", "remove_neg", "remove_only_teacher_neg"} adds
    the Gram-anchoring term of train/ssl_meta_arch.py:527-541 with loss/gram_loss.py:13-50 (pinned against the
    reference's GramLoss, tests/golden).  PARITY UNPINNED for the call path: the reference's own
    (ssl_meta_arch.py:337-347: `get_gram_teacher_output` is called but not defined, `.reshpae`) does not execute, so
    the features follow upstream DINOv3: student = the global crops' final-norm patch tokens, teacher = the same
    tokens of the gram teacher (`gram_backbone/...` parameters) or, with ema_teacher, of the EMA teacher."""
    n_g, n_l = cfg.n_global, cfg.n_local
    g = batch["collated_global_crops"].to(dtype)      # bf16 crops promoted at the first conv
    l = batch["collated_local_crops"].to(dtype)
    masks = batch["collated_masks"]
    idx = batch["mask_indices_list"]
    B = l.shape[0] // n_l
    K = cfg.n_prototypes

    # ---- teacher (no gradient: train/train.py:501-513 differentiates w.r.t. student params only)
    with torch.no_grad():
        t_out = backbone_forward(sub(params, "teacher_backbone"), [g], [None], cfg, emu)[0]
        t_cls = t_out["x_norm_clstoken"]                                    # [2B, D]
        t_patch = t_out["x_norm_patchtokens"]                               # [2B, P, D]
        t_buf = t_patch.reshape(-1, t_patch.shape[-1])[idx]                 # :377
        t_patch_logits = head_forward(sub(params, "teacher_ibot_head"), t_buf, emu)      # :378
        t_cls_logits = head_forward(sub(params, "teacher_dino_head"), t_cls, emu)        # :380
        if centers is None:
            cls_centered = sinkhorn_knopp(t_cls_logits, teacher_temp, B_total=t_cls_logits.shape[0] * world,
                                          allreduce=allreduce).reshape(n_g, B, K)            # :382-386
            n_masked = batch["n_masked_patches"].sum().to(dtype)
            if allreduce is not None:
                n_masked = allreduce(n_masked)
            patch_centered = sinkhorn_knopp(t_patch_logits, teacher_temp, B_total=n_masked, allreduce=allreduce)  # :388-393
        else:
            # optional softmax-centering path (train.centering != sinkhorn_knopp; disabled by the reference's assert,
            # ssl_meta_arch.py:49): update the center first, then softmax((x - center)/temp) (dino_clstoken_loss.py:24-33)
            from .losses import center_update, softmax_center_teacher
            centers["dino"] = center_update(centers["dino"], t_cls_logits, centers.get("momentum", 0.9))
            centers["ibot"] = center_update(centers["ibot"], t_patch_logits, centers.get("momentum", 0.9))
            cls_centered = softmax_center_teacher(t_cls_logits, centers["dino"], teacher_temp).reshape(n_g, B, K)
            patch_centered = softmax_center_teacher(t_patch_logits, centers["ibot"], teacher_temp)

    # ---- student
    s_g, s_l = backbone_forward(sub(params, "student_backbone"), [g, l], [masks, None], cfg, emu)   # :414-418
    g_cls, g_patch, l_cls = s_g["x_norm_clstoken"], s_g["x_norm_patchtokens"], s_l["x_norm_clstoken"]
    s_buf = g_patch.reshape(-1, g_patch.shape[-1])[idx]                                   # :432
    s_patch_logits = head_forward(sub(params, "student_ibot_head"), s_buf, emu)           # :433
    buf = head_forward(sub(params, "student_dino_head"), torch.cat([g_cls, l_cls], dim=0), emu)   # :435-442
    s_g_logits = buf[: g_cls.shape[0]].reshape(n_g, B, K)
    s_l_logits = buf[g_cls.shape[0]:].reshape(n_l, B, K)

    # ---- losses (:463-525)
    g_terms = n_g * (n_g - 1)                 # dino.global_ignore_diagonal = true (ssl_default_config.yaml:22)
    l_terms = n_g * n_l
    g_scale, l_scale = g_terms / (g_terms + l_terms), l_terms / (g_terms + l_terms)
    L_local = dino_loss(s_l_logits, cls_centered, cfg.student_temp, ignore_diagonal=False)
    L_global = dino_loss(s_g_logits, cls_centered, cfg.student_temp, ignore_diagonal=True)
    g_cls_pre = g_cls.reshape(n_g, B, -1)
    L_koleo = sum(koleo_loss(x) for x in g_cls_pre) / n_g                                  # :513
    L_ibot = ibot_loss_masked(s_patch_logits, patch_centered, cfg.student_temp, n_mask_rows=masks.shape[0])
    loss = (cfg.dino_loss_weight * l_scale * 1.0 * L_local + cfg.dino_loss_weight * g_scale * L_global
            + cfg.koleo_loss_weight * n_g * L_koleo + cfg.ibot_loss_weight * L_ibot)
    metrics = {"dino_local_crops_loss": L_local.detach(), "dino_local_loss_weight": torch.tensor(1.0),
               "dino_global_crops_loss": L_global.detach(), "koleo_loss": L_koleo.detach(),
               "ibot_loss": L_ibot.detach(), "local_batch_size": torch.tensor(float(B))}
    if gram is not None:
        with torch.no_grad():
            if gram.get("ema_teacher", False):
                gt_patch = t_patch
            else:
                # the gram teacher sees its own crops when the loader provides them (crops.gram_teacher_crops_size,
                # data/augmentations.py:197-205); features at another resolution are resized to the student's patch grid
                # (upstream get_gram_teacher_output: F.interpolate, gram.global_teacher_resize_method / _antialias)
                gc = batch.get("collated_gram_teacher_crops", None)
                gin = g if gc is None else gc.to(dtype)
                gt_patch = backbone_forward(sub(params, "gram_backbone"), [gin], [None], cfg, emu)[0]["x_norm_patchtokens"]
                if gt_patch.shape[1] != g_patch.shape[1]:
                    n_, Pg, Dm = gt_patch.shape
                    Hg, Hs = int(round(Pg ** 0.5)), int(round(g_patch.shape[1] ** 0.5))
                    gt_patch = torch.nn.functional.interpolate(
                        gt_patch.reshape(n_, Hg, Hg, Dm).permute(0, 3, 1, 2), size=(Hs, Hs), mode="bicubic", align_corners=False,
                        antialias=bool(gram.get("resize_antialias", False))).permute(0, 2, 3, 1).reshape(n_, Hs * Hs, Dm)
        gs_patch = g_patch
        used = gram.get("tokens_used", "all")                    # train/ssl_meta_arch.py:221-223 (upstream: patches[masks] / [~masks])
        if used != "all":
            assert not gram.get("img_level", False)
            sel = masks if used == "masked" else ~masks
            gs_patch, gt_patch = g_patch[sel], gt_patch[sel]
        L_gram = gram_loss(gs_patch, gt_patch, apply_norm=gram.get("normalized", True), img_level=gram.get("img_level", False),
                           remove_neg=gram.get("remove_neg", False),
                           remove_only_teacher_neg=gram.get("remove_only_teacher_neg", False))
        loss = loss + gram["weight"] * L_gram
        metrics["gram_loss"] = L_gram.detach()
        metrics["gram_loss_weight"] = torch.tensor(float(gram["weight"]))
    if return_aux:
        aux = {"t_cls": t_cls, "t_patch_logits": t_patch_logits, "t_cls_logits": t_cls_logits,
               "cls_centered": cls_centered, "patch_centered": patch_centered, "g_cls": g_cls, "l_cls": l_cls,
               "g_patch": g_patch, "s_patch_logits": s_patch_logits, "s_g_logits": s_g_logits,
               "s_l_logits": s_l_logits, "t_prenorm": t_out["x_prenorm"], "s_g_prenorm": s_g["x_prenorm"],
               "s_l_prenorm": s_l["x_prenorm"]}
        return loss, metrics, aux
    return loss, metrics



def ssl_forward_multi(params: dict, batches: list, teacher_temp: float, cfg: ModelCfg, emu: Emu = Emu(False),
                      dtype=torch.float32):
    """The same step on a 1-D "dp" mesh of len(batches) devices, each holding its own images and rank-local mask
    indices (intended semantics): Sinkhorn row sums / totals are psum'ed over ranks
    (loss/dino_clstoken_loss.py:46,53; loss/ibot_patch_loss.py:84,91,99), i.e. the normalisation runs over the
    concatenation of all ranks' teacher rows; every other term is rank-local (KoLeo included, loss/koleo_loss.py:16-35)
    and the scalar loss is the mean over ranks (train/ssl_meta_arch.py:361).  Returns (loss, [metrics per rank])."""
    world = len(batches)
    n_g, K = cfg.n_global, cfg.n_prototypes
    t_cls_l, t_patch_l = [], []
    with torch.no_grad():
        for b in batches:
            g = b["collated_global_crops"].to(dtype)
            t_out = backbone_forward(sub(params, "teacher_backbone"), [g], [None], cfg, emu)[0]
            t_buf = t_out["x_norm_patchtokens"].reshape(-1, cfg.embed_dim)[b["mask_indices_list"]]
            t_patch_l.append(head_forward(sub(params, "teacher_ibot_head"), t_buf, emu))
            t_cls_l.append(head_forward(sub(params, "teacher_dino_head"), t_out["x_norm_clstoken"], emu))
        cls_all = sinkhorn_knopp(torch.cat(t_cls_l), teacher_temp, B_total=sum(t.shape[0] for t in t_cls_l))
        patch_all = sinkhorn_knopp(torch.cat(t_patch_l), teacher_temp, B_total=float(sum(t.shape[0] for t in t_patch_l)))
    total, mets = 0.0, []
    c0 = p0 = 0
    for r, b in enumerate(batches):
        nc, npatch = t_cls_l[r].shape[0], t_patch_l[r].shape[0]
        B = nc // n_g
        cls_centered = cls_all[c0:c0 + nc].reshape(n_g, B, K)
        patch_centered = patch_all[p0:p0 + npatch]
        c0 += nc; p0 += npatch
        loss, m = _student_losses(params, b, cls_centered, patch_centered, cfg, emu, dtype)
        total = total + loss / world
        mets.append(m)
    return total, mets


def _student_losses(params, batch, cls_centered, patch_centered, cfg, emu, dtype):
    n_g, n_l, K = cfg.n_global, cfg.n_local, cfg.n_prototypes
    g = batch["collated_global_crops"].to(dtype)
    l = batch["collated_local_crops"].to(dtype)
    masks, idx = batch["collated_masks"], batch["mask_indices_list"]
    B = l.shape[0] // n_l
    s_g, s_l = backbone_forward(sub(params, "student_backbone"), [g, l], [masks, None], cfg, emu)
    g_cls, g_patch, l_cls = s_g["x_norm_clstoken"], s_g["x_norm_patchtokens"], s_l["x_norm_clstoken"]
    s_patch_logits = head_forward(sub(params, "student_ibot_head"), g_patch.reshape(-1, g_patch.shape[-1])[idx], emu)
    buf = head_forward(sub(params, "student_dino_head"), torch.cat([g_cls, l_cls], dim=0), emu)
    s_g_logits, s_l_logits = buf[: g_cls.shape[0]].reshape(n_g, B, K), buf[g_cls.shape[0]:].reshape(n_l, B, K)
    g_terms, l_terms = n_g * (n_g - 1), n_g * n_l
    g_scale, l_scale = g_terms / (g_terms + l_terms), l_terms / (g_terms + l_terms)
    L_local = dino_loss(s_l_logits, cls_centered, cfg.student_temp, ignore_diagonal=False)
    L_global = dino_loss(s_g_logits, cls_centered, cfg.student_temp, ignore_diagonal=True)
    L_koleo = sum(koleo_loss(x) for x in g_cls.reshape(n_g, B, -1)) / n_g
    L_ibot = ibot_loss_masked(s_patch_logits, patch_centered, cfg.student_temp, n_mask_rows=masks.shape[0])
    loss = (cfg.dino_loss_weight * l_scale * L_local + cfg.dino_loss_weight * g_scale * L_global
            + cfg.koleo_loss_weight * n_g * L_koleo + cfg.ibot_loss_weight * L_ibot)
    return loss, {"dino_local_crops_loss": L_local.detach(), "dino_global_crops_loss": L_global.detach(),
                  "koleo_loss": L_koleo.detach(), "ibot_loss": L_ibot.detach()}


def grads_multi(params: dict, batches: list, teacher_temp: float, cfg: ModelCfg, dtype=torch.float32):
    """Loss and student gradients of the multi-rank step (what reduce-scatter(mean) must reproduce on every shard)."""
    student = {k: v.detach().to(dtype).requires_grad_(True) for k, v in params.items() if k.startswith("student_")}
    full = {k: v.detach().to(dtype) for k, v in params.items()}
    full.update(student)
    loss, mets = ssl_forward_multi(full, batches, teacher_temp, cfg, dtype=dtype)
    keys = list(student)
    gl = torch.autograd.grad(loss, [student[k] for k in keys], allow_unused=True)
    return loss.detach(), mets, {k: (g if g is not None else torch.zeros_like(student[k])) for k, g in zip(keys, gl)}


# ---------------------------------------------------------------------------------------------------- optimiser
def param_multipliers(names, depth: int, layerwise_decay: float = 0.9, patch_embed_lr_mult: float = 0.2,
                      dino_head_wd_multiplier: float = 1.0) -> dict:
    """train/param_groups.py:56-96 (groups) and :104-134 (layer-wise decay). `names` are full student names
    "student_backbone/blocks_3/attn/qkv/kernel"; returns name -> (lr_mult, wd_mult, is_last_layer)."""
    out = {}
    for full in names:
        mod, name = full.split("/", 1)
        is_backbone = mod.endswith("backbone")
        layer_id = depth + 1 if is_backbone else 1
        n_layers = depth if is_backbone else 0
        if is_backbone:
            if any(t in name for t in ("pos_embed", "patch_embed", "mask_token", "cls_token", "storage_tokens")):
                layer_id = 0
            elif "blocks_" in name:
                layer_id = int(name.split("blocks_")[1].split("/")[0]) + 1
        lr_mult = layerwise_decay ** (n_layers + 1 - layer_id)
        wd_mult = 1.0
        if "dino_head" in name:
            wd_mult = dino_head_wd_multiplier
        is_last = "last_layer" in name
        if name.endswith("bias") or "norm" in name or "gamma" in name:
            wd_mult = 0.0
        if "patch_embed" in name:
            lr_mult *= patch_embed_lr_mult
        out[full] = (lr_mult, wd_mult, is_last)
    return out


def clip_by_module(grads: dict, max_norm: float):
    """train/train.py:516-541: per top-level student submodule, g * min(1, max_norm / (||g|| + 1e-6))."""
    norms = {}
    out = {}
    for mod in STUDENT_MODULES:
        keys = [k for k in grads if k.startswith(mod + "/")]
        norm = torch.sqrt(sum((grads[k].double() ** 2).sum() for k in keys)).to(grads[keys[0]].dtype)
        scale = torch.clamp(max_norm / (norm + 1e-6), max=1.0)
        for k in keys:
            out[k] = grads[k] * scale
        norms[f"{mod}_grad_norm"] = norm
    return out, norms


def adamw_update(p, g, m, v, step: int, lr: float, wd: float, b1=0.9, b2=0.999, eps=1e-8):
    """optax.adamw (0.2.5): scale_by_adam(eps, eps_root=0) -> add_decayed_weights -> scale by -lr.
    `step` is the 1-based count after increment (bias correction uses it)."""
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    mh = m / (1 - b1 ** step)
    vh = v / (1 - b2 ** step)
    upd = mh / (torch.sqrt(vh) + eps) + wd * p
    return p - lr * upd, m, v


def train_step(params: dict, opt_state: dict, batch: dict, cfg: ModelCfg, *, lr: float, wd: float,
               last_layer_lr: float, momentum: float, teacher_temp: float, emu: Emu = Emu(False),
               dtype=torch.float32, mults: dict | None = None):
    """train/train.py:491-565 with the intended semantics: the update IS applied, per-group
    multipliers are honoured, and the teacher follows the student by EMA (train/ssl_meta_arch.py:644-660).
    opt_state = {"step": int, "m": {...}, "v": {...}} over student parameters.
    Returns (new_params, new_opt_state, loss, metrics, grads_unclipped)."""
    student = {k: v.detach().to(dtype).requires_grad_(True) for k, v in params.items() if k.startswith("student_")}
    full = {k: v.detach().to(dtype) for k, v in params.items()}
    full.update(student)
    loss, metrics = ssl_forward(full, batch, teacher_temp, cfg, emu, dtype=dtype)
    keys = list(student.keys())
    gl = torch.autograd.grad(loss, [student[k] for k in keys], allow_unused=True)
    grads = {k: (g if g is not None else torch.zeros_like(student[k])) for k, g in zip(keys, gl)}
    clipped, norms = clip_by_module(grads, cfg.clip_grad) if cfg.clip_grad else (grads, {})
    metrics = dict(metrics)
    metrics.update({k: v.detach() for k, v in norms.items()})
    if mults is None:
        mults = param_multipliers(keys, cfg.depth)
    step = opt_state["step"] + 1
    new_params = {k: v.detach().clone() for k, v in params.items()}
    new_m, new_v = {}, {}
    for k in keys:
        lr_mult, wd_mult, is_last = mults[k]
        base_lr = last_layer_lr if is_last else lr
        p, m, v = adamw_update(student[k].detach(), clipped[k], opt_state["m"][k].to(dtype), opt_state["v"][k].to(dtype),
                               step, lr_mult * base_lr, wd_mult * wd)
        new_params[k] = p
        new_m[k], new_v[k] = m, v
    for k in keys:                                   # EMA: teacher <- m*teacher + (1-m)*student (updated student)
        tk = "teacher_" + k[len("student_"):]
        new_params[tk] = params[tk].to(dtype) * momentum + new_params[k] * (1 - momentum)
    return new_params, {"step": step, "m": new_m, "v": new_v}, loss.detach(), metrics, grads


def init_opt_state(params: dict, dtype=torch.float32) -> dict:
    keys = [k for k in params if k.startswith("student_")]
    return {"step": 0, "m": {k: torch.zeros_like(params[k], dtype=dtype) for k in keys},
            "v": {k: torch.zeros_like(params[k], dtype=dtype) for k in keys}}


# ---------------------------------------------------------------------------------------------------- schedules
def cosine_schedule(base_value, final_value, total_iters, warmup_iters=0, start_warmup_value=0, freeze_iters=0):
    """train/cosine_lr_scheduler.py:14-52 (trunc_extra == 0 branch; the other branch is broken upstream)."""
    freeze = np.zeros((freeze_iters,))
    warm = np.linspace(start_warmup_value, base_value, warmup_iters)
    iters = np.arange(total_iters - warmup_iters - freeze_iters)
    sched = final_value + 0.5 * (base_value - final_value) * (1 + np.cos(np.pi * iters / len(iters)))
    out = np.concatenate([freeze, warm, sched]).astype(np.float64)
    assert len(out) == total_iters
    return out
