"""Oracle: ViT backbone + DINO head forward (PyTorch fp32/fp64 on CPU) — test infrastructure.

Parameters are a flat dict  "<module>/<flax path>" -> tensor  with the reference's names and layouts
(SURVEY.md Appendix C): Dense kernels are [in, out], the patch-embed conv kernel is [p, p, 3, D].
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

from .arch import ModelCfg

BF16 = torch.bfloat16


# ---------------------------------------------------------------------------------------------------- rounding
class _RoundSTE(torch.autograd.Function):
    """bf16 round trip in forward and/or backward (straight-through).  Only used when the oracle is asked to
    emulate the engine's bf16 storage points; the default oracle path is pure fp32/fp64."""

    @staticmethod
    def forward(ctx, x, fwd, bwd):
        ctx.bwd = bwd
        return x.to(BF16).to(x.dtype) if fwd else x

    @staticmethod
    def backward(ctx, g):
        return (g.to(BF16).to(g.dtype) if ctx.bwd else g), None, None


class Emu:
    """Rounding policy. Emu(False) = exact reference arithmetic. Emu(True) = engine-like bf16 storage."""

    def __init__(self, on: bool = False):
        self.on = on

    def act(self, x):   # activation stored in bf16 by the engine (forward value and its gradient)
        return _RoundSTE.apply(x, True, True) if self.on else x

    def w(self, x):     # weight matrix cast to bf16 for the tensor cores (gradient stays fp32)
        return _RoundSTE.apply(x, True, False) if self.on else x

    def grad(self, x):  # identity forward, gradient rounded to bf16
        return _RoundSTE.apply(x, False, True) if self.on else x


# ---------------------------------------------------------------------------------------------------- initialisers
def _trunc_normal(shape, std, lo, hi, gen, dtype):
    # jax.nn.initializers.truncated_normal(stddev, lower, upper): bounds in units of sigma
    t = torch.empty(shape, dtype=torch.float64)
    torch.nn.init.trunc_normal_(t, mean=0.0, std=1.0, a=lo, b=hi, generator=gen)
    return (t * std).to(dtype)


def _lecun_normal(shape, fan_in, gen, dtype):
    # flax default kernel_init = variance_scaling(1.0, "fan_in", "truncated_normal")
    std = math.sqrt(1.0 / fan_in) / 0.87962566103423978
    return _trunc_normal(shape, std, -2.0, 2.0, gen, dtype)


def init_backbone(cfg: ModelCfg, gen: torch.Generator, dtype=torch.float32) -> dict:
    """models/vision_transformer.py:86-171 (param creation), layers/*.py initialisers."""
    D, p, Hd = cfg.embed_dim, cfg.patch, cfg.hidden
    P = {}
    P["patch_embed/proj/kernel"] = _lecun_normal((p, p, 3, D), p * p * 3, gen, dtype)  # layers/patch_embed.py:38-42
    P["patch_embed/proj/bias"] = torch.zeros(D, dtype=dtype)
    P["cls_token"] = (torch.randn((1, 1, D), generator=gen, dtype=torch.float64) * 0.02).to(dtype)  # :95-99
    P["mask_token"] = torch.zeros((1, D), dtype=dtype)                                             # :165-169
    if cfg.n_storage:                                                                              # :106-111
        P["storage_tokens"] = (torch.randn((1, cfg.n_storage, D), generator=gen, dtype=torch.float64) * 0.02).to(dtype)
    for i in range(cfg.depth):
        b = f"blocks_{i}/"
        P[b + "norm1/scale"] = torch.ones(D, dtype=dtype)
        P[b + "norm1/bias"] = torch.zeros(D, dtype=dtype)
        P[b + "attn/qkv/kernel"] = _lecun_normal((D, 3 * D), D, gen, dtype)   # layers/attention.py:63
        P[b + "attn/qkv/bias"] = torch.zeros(3 * D, dtype=dtype)
        P[b + "attn/proj/kernel"] = _lecun_normal((D, D), D, gen, dtype)      # layers/attention.py:65
        P[b + "attn/proj/bias"] = torch.zeros(D, dtype=dtype)
        P[b + "ls1/gamma"] = torch.full((D,), cfg.layerscale, dtype=dtype)    # layers/layer_scale.py:12-21
        P[b + "norm2/scale"] = torch.ones(D, dtype=dtype)
        P[b + "norm2/bias"] = torch.zeros(D, dtype=dtype)
        if cfg.ffn_layer == "swiglu":                                         # layers/ffn_layers.py:62-69
            Hs = cfg.swiglu_hidden
            for w_, (i_, o_) in (("w1", (D, Hs)), ("w2", (D, Hs)), ("w3", (Hs, D))):
                P[b + f"mlp/{w_}/kernel"] = _lecun_normal((i_, o_), i_, gen, dtype)
                P[b + f"mlp/{w_}/bias"] = torch.zeros(o_, dtype=dtype)
        else:
            P[b + "mlp/Dense_0/kernel"] = _lecun_normal((D, Hd), D, gen, dtype)   # layers/ffn_layers.py:36-39
            P[b + "mlp/Dense_0/bias"] = torch.zeros(Hd, dtype=dtype)
            P[b + "mlp/Dense_1/kernel"] = _lecun_normal((Hd, D), Hd, gen, dtype)  # layers/ffn_layers.py:43-46
            P[b + "mlp/Dense_1/bias"] = torch.zeros(D, dtype=dtype)
        P[b + "ls2/gamma"] = torch.full((D,), cfg.layerscale, dtype=dtype)
    P["norm/scale"] = torch.ones(D, dtype=dtype)
    P["norm/bias"] = torch.zeros(D, dtype=dtype)
    return P


def init_head(cfg: ModelCfg, gen: torch.Generator, dtype=torch.float32) -> dict:
    """layers/dino_head.py:15-43,65-74: 3-layer MLP + bias-free prototype layer, truncated-normal(0.02)."""
    D, Hh, Bn, K = cfg.embed_dim, cfg.head_hidden, cfg.head_bottleneck, cfg.n_prototypes
    P = {}
    dims = [(D, Hh), (Hh, Hh), (Hh, Bn)]
    for idx, (i, o) in zip((0, 2, 4), dims):
        P[f"mlp/layers_{idx}/kernel"] = _trunc_normal((i, o), 0.02, -1.0, 1.0, gen, dtype)
        P[f"mlp/layers_{idx}/bias"] = torch.zeros(o, dtype=dtype)
    P["last_layer/kernel"] = _trunc_normal((Bn, K), 0.02, -1.0, 1.0, gen, dtype)
    return P


MODULES = ("backbone", "dino_head", "ibot_head")


def hash_uniform(n: int, key: int) -> np.ndarray:
    """n float64 values in [-1, 1), a pure function of (index, key): splitmix64 finaliser in uint64 arithmetic, so the
    same numbers come out on every numpy / platform (fixtures that would be too large to commit are regenerated from
    this instead of being stored)."""
    with np.errstate(over="ignore"):
        x = np.arange(n, dtype=np.uint64) + np.uint64(key % (1 << 32)) * np.uint64(0x9E3779B97F4A7C15)
        x ^= x >> np.uint64(30); x *= np.uint64(0xBF58476D1CE4E5B9)
        x ^= x >> np.uint64(27); x *= np.uint64(0x94D049BB133111EB)
        x ^= x >> np.uint64(31)
    return (x >> np.uint64(40)).astype(np.float64) / float(1 << 23) - 1.0


def formula_params(cfg: ModelCfg, seed: int = 0, dtype=torch.float64) -> dict:
    """Same tree as init_params, every leaf a closed-form function of (name, index, seed) — lecun-scaled kernels,
    non-trivial biases / LN affine / LayerScale / tokens, and a teacher that differs from the student."""
    import zlib
    out = {}
    for name, t in init_params(cfg, 0, dtype=torch.float32).items():
        leaf = name.split("/", 1)[1]
        u = hash_uniform(t.numel(), zlib.crc32(name.encode()) + 7919 * seed).reshape(tuple(t.shape))
        if leaf.endswith("kernel"):
            fan_in = t.numel() // t.shape[-1]
            amp = (3.0 / fan_in) ** 0.5 * (2.0 if "head" in name else 1.0)
            v = amp * u
        elif leaf.endswith("scale"):
            v = 1.0 + 0.1 * u
        elif leaf.endswith("gamma"):
            v = 0.5 + 0.2 * u
        else:                                   # biases, cls_token, mask_token
            v = 0.05 * u
        out[name] = torch.from_numpy(v).to(dtype)
    return out


def formula_images(shape, key: int, dtype=torch.float64):
    """Unit-variance synthetic crops from hash_uniform (see formula_params)."""
    n = int(np.prod(shape))
    return torch.from_numpy(hash_uniform(n, key).reshape(shape) * 3.0 ** 0.5).to(dtype)


def init_params(cfg: ModelCfg, seed: int = 0, dtype=torch.float32, teacher_copy: bool = True,
                perturb: float = 0.0) -> dict:
    """Full parameter dict with the reference's six top-level modules (train/ssl_meta_arch.py:62-64,86-87,130-131).

    teacher_copy=True starts the teacher equal to the student (upstream DINO intent); the reference initialises the
    teacher modules independently.  `perturb` adds N(0, perturb^2) to biases / LN affine / gamma / mask_token so that
    parity fixtures exercise every term (at the reference init several of them are exactly 0 or 1e-5).
    """
    gen = torch.Generator().manual_seed(seed)
    out = {}
    student = {"backbone": init_backbone(cfg, gen, dtype), "dino_head": init_head(cfg, gen, dtype),
               "ibot_head": init_head(cfg, gen, dtype)}
    if perturb > 0:
        for sub in student.values():
            for k, v in sub.items():
                if k.endswith("/bias") or k.endswith("/scale") or k.endswith("/gamma") or k in ("mask_token", "storage_tokens"):
                    v.add_(torch.randn(v.shape, generator=gen, dtype=torch.float64).to(dtype) * perturb)
    for m in MODULES:
        for k, v in student[m].items():
            out[f"student_{m}/{k}"] = v
    if teacher_copy:
        tgen = torch.Generator().manual_seed(seed + 1)
        for m in MODULES:
            for k, v in student[m].items():
                t = v.clone()
                if perturb > 0:  # make the teacher differ from the student so EMA / CE terms are non-trivial
                    t.add_(torch.randn(t.shape, generator=tgen, dtype=torch.float64).to(dtype) * perturb * 0.1)
                out[f"teacher_{m}/{k}"] = t
    else:
        teacher = {"backbone": init_backbone(cfg, gen, dtype), "dino_head": init_head(cfg, gen, dtype),
                   "ibot_head": init_head(cfg, gen, dtype)}
        for m in MODULES:
            for k, v in teacher[m].items():
                out[f"teacher_{m}/{k}"] = v
    return out


def sub(params: dict, prefix: str) -> dict:
    pl = len(prefix) + 1
    return {k[pl:]: v for k, v in params.items() if k.startswith(prefix + "/")}


# ---------------------------------------------------------------------------------------------------- layers
def layer_norm(x, scale, bias, eps):
    """flax nn.LayerNorm(epsilon=1e-6), use_fast_variance: var = E[x^2] - E[x]^2 (models/vision_transformer.py:40)."""
    mean = x.mean(-1, keepdim=True)
    var = ((x * x).mean(-1, keepdim=True) - mean * mean).clamp_min(0.0)
    return (x - mean) * torch.rsqrt(var + eps) * scale + bias


def gelu(x):
    """flax nn.gelu default approximate=True (tanh form)."""
    return F.gelu(x, approximate="tanh")


def rope_sincos(Hp: int, Wp: int, head_dim: int, base: float, dtype):
    """layers/rope_position_encoding.py:36-40 (periods), :64-73 (coords, 'separate'), :117-123 (angles, tile x2)."""
    periods = base ** (2.0 * torch.arange(head_dim // 4, dtype=dtype) / (head_dim // 2))
    ch = torch.arange(0.5, Hp, dtype=dtype) / Hp
    cw = torch.arange(0.5, Wp, dtype=dtype) / Wp
    coords = torch.stack(torch.meshgrid(ch, cw, indexing="ij"), dim=-1).reshape(-1, 2)
    coords = 2.0 * coords - 1.0
    ang = 2 * math.pi * coords[:, :, None] / periods[None, None, :]
    ang = ang.reshape(ang.shape[0], -1)
    ang = torch.cat([ang, ang], dim=-1)
    return torch.sin(ang), torch.cos(ang)


def rope_rotate_half(x):
    """layers/attention.py:14-16."""
    x1, x2 = x.chunk(2, dim=-1)
    return torch.cat([-x2, x1], dim=-1)


def rope_apply(x, sin, cos):
    """layers/attention.py:19-20."""
    return x * cos + rope_rotate_half(x) * sin


def attention(qkv, heads, sin, cos, emu: Emu):
    """layers/attention.py:106-118 (compute_attention) + :69-90 (apply_rope) + flax dot_product_attention."""
    n, N, D3 = qkv.shape
    hd = D3 // 3 // heads
    qkv = qkv.reshape(n, N, 3, heads, hd)
    q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]          # [n, N, H, hd]
    q, k = q.transpose(1, 2), k.transpose(1, 2)                  # [n, H, N, hd]
    prefix = N - sin.shape[0]
    q = torch.cat([q[:, :, :prefix], rope_apply(q[:, :, prefix:], sin, cos)], dim=2)
    k = torch.cat([k[:, :, :prefix], rope_apply(k[:, :, prefix:], sin, cos)], dim=2)
    q, k = emu.act(q), emu.act(k)
    v = v.transpose(1, 2)
    s = torch.einsum("bhqd,bhkd->bhqk", q / math.sqrt(hd), k)
    a = emu.act(torch.softmax(s, dim=-1))
    o = torch.einsum("bhqk,bhkd->bhqd", a, v)
    return o.transpose(1, 2).reshape(n, N, heads * hd)


def block_forward(P: dict, b: str, x, sin, cos, cfg: ModelCfg, emu: Emu):
    """layers/block.py:195-201 (deterministic branch): x + ls1(attn(norm1 x)); x + ls2(mlp(norm2 x))."""
    y = emu.act(layer_norm(x, P[b + "norm1/scale"], P[b + "norm1/bias"], cfg.ln_eps))
    qkv_bias = P[b + "attn/qkv/bias"]
    if cfg.mask_k_bias:                         # upstream LinearKMaskedBias: bias * [1 | 0 | 1] (oracle only)
        D_ = qkv_bias.shape[0] // 3
        qkv_bias = torch.cat([qkv_bias[:D_], torch.zeros_like(qkv_bias[D_:2 * D_]), qkv_bias[2 * D_:]])
    qkv = emu.act(y @ emu.w(P[b + "attn/qkv/kernel"]) + qkv_bias)
    o = emu.act(attention(qkv, cfg.heads, sin, cos, emu))
    p = o @ emu.w(P[b + "attn/proj/kernel"]) + P[b + "attn/proj/bias"]
    x = x + P[b + "ls1/gamma"] * emu.grad(p)
    z = emu.act(layer_norm(x, P[b + "norm2/scale"], P[b + "norm2/bias"], cfg.ln_eps))
    if cfg.ffn_layer == "swiglu":                                                # layers/ffn_layers.py:71-76
        x1 = emu.grad(z @ emu.w(P[b + "mlp/w1/kernel"]) + P[b + "mlp/w1/bias"])
        x2 = emu.grad(z @ emu.w(P[b + "mlp/w2/kernel"]) + P[b + "mlp/w2/bias"])
        h = emu.act(F.silu(x1) * x2)
        m = emu.grad(h @ emu.w(P[b + "mlp/w3/kernel"]) + P[b + "mlp/w3/bias"])
        return x + P[b + "ls2/gamma"] * m
    u1 = emu.grad(z @ emu.w(P[b + "mlp/Dense_0/kernel"]) + P[b + "mlp/Dense_0/bias"])
    h = emu.act(gelu(u1))                                                        # layers/ffn_layers.py:36-40
    u2 = emu.grad(h @ emu.w(P[b + "mlp/Dense_1/kernel"]) + P[b + "mlp/Dense_1/bias"])
    m = gelu(u2) if cfg.mlp_second_act else u2                                   # layers/ffn_layers.py:43-47
    x = x + P[b + "ls2/gamma"] * m
    return x


def patch_embed(P: dict, x, cfg: ModelCfg, emu: Emu):
    """layers/patch_embed.py:45-55: NHWC conv, kernel = stride = patch  ==  [n*P, p*p*3] x [p*p*3, D] + bias."""
    n, H, W, Cc = x.shape
    p = cfg.patch
    assert H % p == 0 and W % p == 0
    Hp, Wp = H // p, W // p
    patches = x.reshape(n, Hp, p, Wp, p, Cc).permute(0, 1, 3, 2, 4, 5).reshape(n, Hp * Wp, p * p * Cc)
    w = emu.w(P["patch_embed/proj/kernel"]).reshape(p * p * Cc, -1)
    return patches @ w + P["patch_embed/proj/bias"], (Hp, Wp)


def backbone_forward(P: dict, x_list, masks_list, cfg: ModelCfg, emu: Emu = Emu(False)):
    """models/vision_transformer.py:173-247 (prepare_tokens_with_masks + forward_features_list).

    x_list: list of [n, H, W, 3] crops (already in compute dtype); masks_list: list of bool [n, P] or None.
    Returns one dict per crop set with x_norm_clstoken [n, D] and x_norm_patchtokens [n, P, D].
    """
    toks, ropes = [], []
    for x, masks in zip(x_list, masks_list):
        t, (Hp, Wp) = patch_embed(P, x, cfg, emu)
        if masks is not None:                                   # :178-184
            t = torch.where(masks[..., None], P["mask_token"].to(t.dtype)[None], t)
            cls = P["cls_token"]
        else:                                                   # :185-187
            cls = P["cls_token"] + 0 * P["mask_token"]
        parts = [cls.expand(t.shape[0], -1, -1)]
        if cfg.n_storage:                                       # :189-201 register tokens between cls and patches
            parts.append(P["storage_tokens"].to(t.dtype).expand(t.shape[0], -1, -1))
        t = torch.cat(parts + [t], dim=1)
        toks.append(t)
        sc = rope_sincos(Hp, Wp, cfg.head_dim, cfg.rope_base, t.dtype)
        ropes.append((sc[0].to(t.device), sc[1].to(t.device)))
    for i in range(cfg.depth):
        toks = [block_forward(P, f"blocks_{i}/", t, s, c, cfg, emu) for t, (s, c) in zip(toks, ropes)]
    outs = []
    for t in toks:
        xn = layer_norm(t, P["norm/scale"], P["norm/bias"], cfg.ln_eps)      # :234
        R = cfg.n_storage                                                    # :231-245
        outs.append({"x_norm_clstoken": xn[:, 0], "x_storage_tokens": xn[:, 1:1 + R], "x_norm_patchtokens": xn[:, 1 + R:],
                     "x_prenorm": t})
    return outs


def head_forward(P: dict, x, emu: Emu = Emu(False), last_layer: bool = True):
    """layers/dino_head.py:78-85: MLP(GELU) -> x / (||x|| + 1e-12) -> bias-free prototype layer
    (last_layer=False is the reference's `no_last_layer=True`)."""
    x = emu.act(x)
    u = emu.act(gelu(emu.grad(x @ emu.w(P["mlp/layers_0/kernel"]) + P["mlp/layers_0/bias"])))
    u = emu.act(gelu(emu.grad(u @ emu.w(P["mlp/layers_2/kernel"]) + P["mlp/layers_2/bias"])))
    u = u @ emu.w(P["mlp/layers_4/kernel"]) + P["mlp/layers_4/bias"]
    nrm = torch.linalg.norm(u, ord=2, dim=-1, keepdim=True)
    u = emu.act(u / (nrm + 1e-12))
    if not last_layer:
        return u
    return u @ emu.w(P["last_layer/kernel"])
