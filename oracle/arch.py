"""Model / recipe configuration shared by the oracle and its callers (test infrastructure).

Arch table follows dinov3_jax/models/vision_transformer.py:325-408 (vit_small ... vit_giant2); recipe defaults follow
dinov3_jax/configs/ssl_default_config.yaml (line numbers cited per field).
"""
from __future__ import annotations

from dataclasses import dataclass, replace

# name -> (embed_dim, depth, heads)   models/vision_transformer.py:325-397
ARCHS = {
    "vit_small": (384, 12, 6),
    "vit_base": (768, 12, 12),
    "vit_large": (1024, 24, 16),
    "vit_giant2": (1536, 40, 24),
}


@dataclass(frozen=True)
class ModelCfg:
    embed_dim: int = 384
    depth: int = 12
    heads: int = 6
    patch: int = 16                 # student.patch_size            ssl_default_config.yaml:92
    ffn_ratio: float = 4.0          # student.ffn_ratio             :95
    global_size: int = 224          # crops.global_crops_size       :129
    local_size: int = 96            # crops.local_crops_size        :130
    n_global: int = 2
    n_local: int = 8                # crops.local_crops_number      :128
    n_prototypes: int = 65536       # dino/ibot.head_n_prototypes   :23,50
    head_hidden: int = 2048         # head_hidden_dim               :25,52
    head_bottleneck: int = 256      # head_bottleneck_dim           :24,51
    layerscale: float = 1e-5        # student.layerscale            :98
    rope_base: float = 100.0        # student.pos_embed_rope_base   :100
    student_temp: float = 0.1       # loss/dino_clstoken_loss.py:16, loss/ibot_patch_loss.py:20
    dino_loss_weight: float = 1.0   # :21
    koleo_loss_weight: float = 0.1  # :28
    ibot_loss_weight: float = 1.0   # :42
    mask_probability: float = 0.5   # ibot.mask_sample_probability  :44
    mask_ratio: tuple = (0.1, 0.5)  # ibot.mask_ratio_min_max       :43
    clip_grad: float = 3.0          # optim.clip_grad               :148
    ln_eps: float = 1e-6            # models/vision_transformer.py:40 (layernormbf16: 1e-5, :41)
    n_storage: int = 0              # student.n_storage_tokens (models/vision_transformer.py:106-111)
    ffn_layer: str = "mlp"          # "mlp" | "swiglu" (layers/ffn_layers.py:52-76; oracle only)
    mask_k_bias: bool = False       # student.mask_k_bias: the k third of the qkv bias is masked to zero (upstream DINOv3)
    swiglu_align: int = 8           # swiglu / swiglu32 / swiglu64 / swiglu128 (models/vision_transformer.py:30-36)
    mlp_second_act: bool = True     # layers/ffn_layers.py:47 applies GELU after fc2 as well

    @property
    def head_dim(self) -> int:
        return self.embed_dim // self.heads

    @property
    def hidden(self) -> int:
        return int(self.embed_dim * self.ffn_ratio)

    @property
    def swiglu_hidden(self) -> int:
        d = int(self.hidden * 2 / 3)                       # layers/ffn_layers.py:64-65
        return d + (-d % self.swiglu_align)

    @property
    def prefix(self) -> int:
        return 1 + self.n_storage

    def tokens(self, size: int) -> int:
        return (size // self.patch) ** 2 + self.prefix

    @property
    def n_patches_global(self) -> int:
        return (self.global_size // self.patch) ** 2


def cfg_for(arch: str, **kw) -> ModelCfg:
    d, l, h = ARCHS[arch]
    return replace(ModelCfg(embed_dim=d, depth=l, heads=h), **kw)


def tiny_cfg(**kw) -> ModelCfg:
    """Seconds-on-CPU configuration used by the parity tests and golden fixtures (head_dim stays 64)."""
    base = ModelCfg(embed_dim=128, depth=2, heads=2, global_size=64, local_size=32, n_local=8, n_prototypes=512,
                    head_hidden=256, head_bottleneck=64)
    return replace(base, **kw)
