"""Oracle: DINO / iBOT / KoLeo losses and Sinkhorn-Knopp teacher normalisation — test infrastructure."""
from __future__ import annotations

import torch


def sinkhorn_knopp(logits, teacher_temp: float, B_total, n_iterations: int = 3, allreduce=None):
    """loss/dino_clstoken_loss.py:35-62 and loss/ibot_patch_loss.py:77-109 (identical but for B).

    logits [Bc, K].  B_total = 2B*world (DINO, :40) or sum over ranks of n_masked (iBOT, :82-86).
    `allreduce` (callable tensor->tensor, sum over ranks) stands for jax.lax.psum (:46,53 / :91,99); None = 1 device.
    The reference has no max-subtraction (:39); exp() of the raw logits/temp is kept here as written.
    """
    ar = allreduce if allreduce is not None else (lambda t: t)
    Q = torch.exp(logits / teacher_temp).T           # [K, Bc]
    K = Q.shape[0]
    Q = Q / ar(Q.sum())
    for _ in range(n_iterations):
        Q = Q / ar(Q.sum(dim=1, keepdim=True))       # rows: over samples (all ranks)
        Q = Q / K
        Q = Q / Q.sum(dim=0, keepdim=True)           # columns: over prototypes
        Q = Q / B_total
    Q = Q * B_total
    return Q.T


def softmax_center_teacher(logits, center, teacher_temp: float):
    """loss/dino_clstoken_loss.py:24-33 (optional centering path, disabled by train.centering=sinkhorn_knopp)."""
    return torch.softmax((logits - center) / teacher_temp, dim=-1)


def center_update(center, logits, momentum: float = 0.9, allreduce_mean=None):
    """loss/dino_clstoken_loss.py:91-95."""
    local = logits.mean(dim=0, keepdim=True)
    g = allreduce_mean(local) if allreduce_mean is not None else local
    return center * momentum + g * (1 - momentum)


def dino_loss(student_logits, teacher_probs, student_temp: float, ignore_diagonal: bool):
    """loss/dino_clstoken_loss.py:66-89.  student_logits [S,B,K], teacher_probs [T,B,K] -> scalar."""
    S, B, _ = student_logits.shape
    T = teacher_probs.shape[0]
    lsm = torch.log_softmax(student_logits / student_temp, dim=-1)
    if ignore_diagonal:
        loss = -torch.einsum("sbk,tbk->st", lsm, teacher_probs)
        loss = loss - torch.diag(torch.diagonal(loss))          # fill_diagonal(0) (:75)
        M = min(S, T)
        return loss.sum() / (B * S * T - B * M)
    loss = -torch.einsum("sbk,tbk->", lsm, teacher_probs)
    return loss / (B * S * T)


def ibot_loss_masked(student_logits, teacher_probs, student_temp: float, n_mask_rows: int):
    """loss/ibot_patch_loss.py:13-14,45-67: -sum_i sum_k t*log_softmax(s/temp) / masks.shape[0]; masks_weight is
    computed by the reference but NOT applied (:66 is commented out) — followed as written."""
    lsm = torch.log_softmax(student_logits / student_temp, dim=-1)
    loss = (teacher_probs * lsm).sum(dim=-1)
    return -loss.sum() / n_mask_rows


def koleo_loss(x, eps: float = 1e-8):
    """loss/koleo_loss.py:16-35: nearest neighbour by dot product of L2-normalised rows (diag = -1)."""
    x = x / (torch.linalg.norm(x, ord=2, dim=-1, keepdim=True) + eps)
    dots = x @ x.T
    dots = dots.clone()
    dots.fill_diagonal_(-1.0)
    idx = torch.argmax(dots, dim=1)
    dist = torch.linalg.norm(x - x[idx], ord=2, dim=-1) + eps      # pairwise_distance (:16-17)
    return -torch.log(dist + eps).mean()


def gram_loss(output_feats, target_feats, apply_norm: bool = True, img_level: bool = True, remove_neg: bool = True,
              remove_only_teacher_neg: bool = False):
    """loss/gram_loss.py:13-50 (not on the round-1 GPU path): MSE between the patch-similarity (Gram)
    matrices of student and gram-teacher features, per image ([B, N, D]) or over the whole batch ([B*N, D])."""
    assert not (remove_neg and remove_only_teacher_neg)   # the reference asserts exactly one (gram_loss.py:20); its YAML
    t, s = target_feats, output_feats                      # default (both false, ssl_default_config.yaml:68-69) = no removal
    if apply_norm:
        t = t / torch.linalg.norm(t, dim=-1, keepdim=True)
        s = s / torch.linalg.norm(s, dim=-1, keepdim=True)
    if not img_level:
        t, s = t.reshape(-1, t.shape[-1]), s.reshape(-1, s.shape[-1])
    t_sim, s_sim = t @ t.transpose(-1, -2), s @ s.transpose(-1, -2)
    if remove_neg:
        t_sim, s_sim = t_sim.clamp_min(0.0), s_sim.clamp_min(0.0)
    elif remove_only_teacher_neg:
        s_sim = torch.where((s_sim < 0) & (t_sim < 0), torch.zeros_like(s_sim), s_sim)
        t_sim = t_sim.clamp_min(0.0)
    return ((s_sim - t_sim) ** 2).mean()
