"""Loss objects with the reference's class / method names (dinov3_jax/loss/*.py), forward values computed by the B200
kernels on CUDA tensors.  (The training engine uses the same kernels in their fused forward+backward form.)"""
from __future__ import annotations

import torch

from .. import ops

f32 = torch.float32


def _sinkhorn(teacher_output: torch.Tensor, teacher_temp: float, btot_local: float, n_iterations: int = 3, comm=None):
    L = teacher_output.to(f32).contiguous()
    R, K = L.shape
    dev = L.device
    mx = torch.full((1,), float("-inf"), device=dev)
    btot = torch.tensor([float(btot_local)], device=dev)
    ops.absmax(L, mx)
    if comm is not None:
        comm.all_reduce_max(mx)
        comm.all_reduce_sum(btot)
    s, a_buf, a = torch.zeros(K, device=dev), torch.empty(R, device=dev), None
    for _ in range(n_iterations):
        s.zero_()
        ops.sinkhorn_colsum(L, mx, teacher_temp, a, s)
        if comm is not None:
            comm.all_reduce_sum(s)
        ops.sinkhorn_rowsum(L, mx, teacher_temp, s, btot, a_buf)
        a = a_buf
    Q = torch.empty(R, K, device=dev)
    ops.sinkhorn_probs(L, mx, teacher_temp, s, a, btot, Q)
    return Q


def _ce(student: torch.Tensor, teacher_probs: torch.Tensor, student_temp: float, t0, t1, w):
    """sum_i w_i * CE(student_i, sum of teacher rows t0_i, t1_i) through d3_ce_fwd_bwd (forward only)."""
    S = student.to(f32).contiguous()
    T = teacher_probs.to(f32).contiguous()
    dev = S.device
    metric = torch.zeros(1, device=dev)
    slot = torch.zeros(S.shape[0], dtype=torch.int32, device=dev)
    ops.ce_fwd_bwd(S, student_temp, T, None, 1.0, None, None, None, t0.to(dev), t1.to(dev), w.to(dev), w.to(dev), slot, metric, None)
    return metric[0]


class DINOLoss:
    """loss/dino_clstoken_loss.py:14-95."""

    def __init__(self, out_dim: int, student_temp: float = 0.1, center_momentum: float = 0.9, comm=None):
        self.out_dim, self.student_temp, self.center_momentum, self.comm = out_dim, student_temp, center_momentum, comm

    def sinkhorn_knopp_teacher(self, teacher_output, teacher_temp, n_iterations=3, init_phase=False):
        world = 1 if (self.comm is None or init_phase) else self.comm.world
        return _sinkhorn(teacher_output, float(teacher_temp), teacher_output.shape[0], n_iterations,
                         None if init_phase else self.comm)

    def softmax_center_teacher(self, teacher_output, teacher_temp, update_centers=True):
        raise NotImplementedError("train.centering=softmax is disabled by the reference (ssl_meta_arch.py:49); only sinkhorn_knopp is on the B200 path")

    def __call__(self, student_logits, teacher_probs, ignore_diagonal=False):
        S, B, K = student_logits.shape
        T = teacher_probs.shape[0]
        i = torch.arange(S * B)
        s_idx, b_idx = i // B, i % B
        if ignore_diagonal:
            assert T == 2 and S == 2, "ignore_diagonal pairs each global crop with the other one (S = T = 2)"
            t0 = ((1 - s_idx) * B + b_idx).to(torch.int32)
            t1 = torch.full_like(t0, -1)
            w = torch.full((S * B,), 1.0 / (B * S * T - B * min(S, T)))
        else:
            assert T == 2, "teacher has the two global crops"
            t0, t1 = b_idx.to(torch.int32), (B + b_idx).to(torch.int32)
            w = torch.full((S * B,), 1.0 / (B * S * T))
        return _ce(student_logits.reshape(S * B, K), teacher_probs.reshape(T * B, K), self.student_temp, t0, t1, w)


class iBOTPatchLoss:
    """loss/ibot_patch_loss.py:17-109."""

    def __init__(self, patch_out_dim: int, student_temp: float = 0.1, center_momentum: float = 0.9, comm=None):
        self.patch_out_dim, self.student_temp, self.comm = patch_out_dim, student_temp, comm

    def sinkhorn_knopp_teacher(self, teacher_output, teacher_temp, n_masked_patches_tensor, n_iterations=3, init_phase=False):
        return _sinkhorn(teacher_output, float(teacher_temp), float(n_masked_patches_tensor.sum()), n_iterations,
                         None if init_phase else self.comm)

    def forward_masked(self, student_patch_tokens_masked, teacher_patch_tokens_masked, student_masks_flat,
                       n_masked_patches=None, masks_weight=None):
        M = student_patch_tokens_masked.shape[0] if n_masked_patches is None else int(n_masked_patches)
        t0 = torch.arange(M, dtype=torch.int32)
        w = torch.full((M,), 1.0 / student_masks_flat.shape[0])      # masks_weight is NOT applied by the reference (:66)
        return _ce(student_patch_tokens_masked[:M], teacher_patch_tokens_masked[:M], self.student_temp, t0,
                   torch.full_like(t0, -1), w)


class KoLeoLoss:
    """loss/koleo_loss.py:20-35."""

    def __call__(self, student_output, eps=1e-8):
        x = student_output.to(f32).contiguous()
        B, D = x.shape
        dev = x.device
        met, dx = torch.zeros(1, device=dev), torch.zeros(B, D, device=dev)
        ops.koleo_fwd_bwd(x, torch.empty(B, D, device=dev), torch.empty(B, device=dev),
                          torch.empty(B, dtype=torch.int32, device=dev), torch.empty(B, device=dev), met, dx, 1.0, 0.0, eps)
        return met[0]


class KoLeoLossDistributed:
    def __init__(self, *a, **k):
        raise NotImplementedError("dino.koleo_loss_distributed is off by default (ssl_default_config.yaml:29) and not on the B200 path yet")
