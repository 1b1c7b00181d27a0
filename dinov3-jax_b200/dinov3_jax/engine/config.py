"""Static description of one training configuration for the B200 engine (derived from the reference's YAML keys).

Field names follow dinov3_jax/configs/ssl_default_config.yaml; arch table follows
dinov3_jax/models/vision_transformer.py:325-408.
"""
from __future__ import annotations

from dataclasses import dataclass, replace

ARCHS = {  # name -> (embed_dim, depth, heads)
    "vit_small": (384, 12, 6),
    "vit_base": (768, 12, 12),
    "vit_large": (1024, 24, 16),
    "vit_so400m": (1152, 27, 18),
    "vit_huge2": (1280, 32, 20),
    "vit_giant2": (1536, 40, 24),
}


@dataclass(frozen=True)
class EngineConfig:
    embed_dim: int = 384
    depth: int = 12
    heads: int = 6
    patch: int = 16
    ffn_ratio: float = 4.0
    global_size: int = 224
    local_size: int = 96
    n_global: int = 2
    n_local: int = 8
    n_prototypes: int = 65536
    head_hidden: int = 2048
    head_bottleneck: int = 256
    layerscale: float = 1e-5
    rope_base: float = 100.0
    student_temp: float = 0.1
    dino_loss_weight: float = 1.0
    koleo_loss_weight: float = 0.1
    ibot_loss_weight: float = 1.0
    clip_grad: float = 3.0
    ln_eps: float = 1e-6
    mlp_second_act: bool = True      # reference applies GELU after fc2 too (layers/ffn_layers.py:47)
    layerwise_decay: float = 0.9
    patch_embed_lr_mult: float = 0.2
    dino_head_wd_multiplier: float = 1.0
    adamw_beta1: float = 0.9
    adamw_beta2: float = 0.999
    mask_probability: float = 0.5
    mask_ratio: tuple = (0.1, 0.5)

    @property
    def head_dim(self) -> int:
        return self.embed_dim // self.heads

    @property
    def hidden(self) -> int:
        return int(self.embed_dim * self.ffn_ratio)

    def patches(self, size: int) -> int:
        return (size // self.patch) ** 2

    def tokens(self, size: int) -> int:
        return self.patches(size) + 1


def config_for(arch: str, **kw) -> EngineConfig:
    d, l, h = ARCHS[arch]
    return replace(EngineConfig(embed_dim=d, depth=l, heads=h), **kw)


def from_oracle_cfg(c) -> EngineConfig:
    """Build from any object with the same field names (tests pass oracle.arch.ModelCfg)."""
    names = EngineConfig.__dataclass_fields__.keys()
    return EngineConfig(**{k: getattr(c, k) for k in names if hasattr(c, k)})
