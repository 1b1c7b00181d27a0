"""CPU cross-check of the oracle's ViT forward against an INDEPENDENT implementation of upstream DINOv3: Hugging Face
transformers' `DINOv3ViTModel` (a PyTorch port of facebookresearch/dinov3, the model the JAX reference was ported from).
This does not involve the reference checkout; it pins the pieces the numpy shim can only restate — RoPE frequencies /
patch-centre coordinates / rotate-half convention, which tokens are rotated, attention scaling, LayerScale placement,
token order (cls, registers, patches), mask-token substitution, the flax-vs-torch conv kernel layout — against a second
code base.  Upstream applies GELU once in the MLP (the JAX reference applies it twice), so the oracle runs with
mlp_second_act=False here; layer-norm eps 1e-5 = the reference's `layernormbf16`."""
import pytest
import torch

hf = pytest.importorskip("transformers.models.dinov3_vit")


def _to_hf_state_dict(bp: dict, depth: int, D: int) -> dict:
    sd = {"embeddings.cls_token": bp["cls_token"], "embeddings.mask_token": bp["mask_token"].reshape(1, 1, D),
          "embeddings.register_tokens": bp["storage_tokens"],
          "embeddings.patch_embeddings.weight": bp["patch_embed/proj/kernel"].permute(3, 2, 0, 1).contiguous(),
          "embeddings.patch_embeddings.bias": bp["patch_embed/proj/bias"], "norm.weight": bp["norm/scale"], "norm.bias": bp["norm/bias"]}
    for i in range(depth):
        b, h = f"blocks_{i}/", f"model.layer.{i}."
        qkv_w, qkv_b = bp[b + "attn/qkv/kernel"], bp[b + "attn/qkv/bias"]
        for j, name in enumerate(("q_proj", "k_proj", "v_proj")):            # layers/attention.py:106-112: [.., 3, H, hd] split
            sd[h + f"attention.{name}.weight"] = qkv_w[:, j * D:(j + 1) * D].t().contiguous()
            sd[h + f"attention.{name}.bias"] = qkv_b[j * D:(j + 1) * D]
        sd[h + "attention.o_proj.weight"] = bp[b + "attn/proj/kernel"].t().contiguous()
        sd[h + "attention.o_proj.bias"] = bp[b + "attn/proj/bias"]
        sd[h + "norm1.weight"], sd[h + "norm1.bias"] = bp[b + "norm1/scale"], bp[b + "norm1/bias"]
        sd[h + "norm2.weight"], sd[h + "norm2.bias"] = bp[b + "norm2/scale"], bp[b + "norm2/bias"]
        sd[h + "layer_scale1.lambda1"], sd[h + "layer_scale2.lambda1"] = bp[b + "ls1/gamma"], bp[b + "ls2/gamma"]
        sd[h + "mlp.up_proj.weight"] = bp[b + "mlp/Dense_0/kernel"].t().contiguous()
        sd[h + "mlp.up_proj.bias"] = bp[b + "mlp/Dense_0/bias"]
        sd[h + "mlp.down_proj.weight"] = bp[b + "mlp/Dense_1/kernel"].t().contiguous()
        sd[h + "mlp.down_proj.bias"] = bp[b + "mlp/Dense_1/bias"]
    return sd


@pytest.mark.parametrize("size,n_storage,mask_k_bias", [(64, 4, False), (48, 0, False), (48, 4, True)])
def test_oracle_vit_matches_huggingface_dinov3(size, n_storage, mask_k_bias):
    from oracle.arch import ModelCfg
    from oracle.model import backbone_forward, formula_images, formula_params, sub
    D, depth, heads = 128, 2, 2
    cfg = ModelCfg(embed_dim=D, depth=depth, heads=heads, global_size=size, local_size=32, n_storage=n_storage, ln_eps=1e-5,
                   mlp_second_act=False, mask_k_bias=mask_k_bias, n_prototypes=16, head_hidden=16, head_bottleneck=8)
    bp = sub(formula_params(cfg, 6), "student_backbone")
    if not n_storage:
        bp["storage_tokens"] = torch.zeros(1, 0, D, dtype=torch.float64)
    hcfg = hf.DINOv3ViTConfig(patch_size=16, hidden_size=D, intermediate_size=4 * D, num_hidden_layers=depth, num_attention_heads=heads,
                              hidden_act="gelu_pytorch_tanh", layer_norm_eps=1e-5, rope_theta=100.0, image_size=size, query_bias=True,
                              key_bias=not mask_k_bias, value_bias=True, proj_bias=True, mlp_bias=True, layerscale_value=1.0,
                              num_register_tokens=n_storage, use_gated_mlp=False)
    model = hf.DINOv3ViTModel(hcfg).double().eval()
    sd = _to_hf_state_dict(bp, depth, D)
    if mask_k_bias:                                     # HF drops the key bias altogether (upstream masks it to zero)
        sd = {k: v for k, v in sd.items() if not k.endswith("k_proj.bias")}
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected and all("inv_freq" in k for k in missing), (missing, unexpected)
    n, P = 3, (size // 16) ** 2
    x = formula_images((n, size, size, 3), 77)
    masks = (torch.arange(n * P).reshape(n, P) * 7 % 5 == 0)
    for mk in (masks, None):
        if not n_storage:
            bp_o = {k: v for k, v in bp.items() if k != "storage_tokens"}
        else:
            bp_o = bp
        want = backbone_forward(bp_o, [x], [mk], cfg)[0]
        with torch.no_grad():
            got = model(pixel_values=x.permute(0, 3, 1, 2).contiguous(), bool_masked_pos=mk).last_hidden_state
        ref = torch.cat([want["x_norm_clstoken"][:, None], want["x_storage_tokens"], want["x_norm_patchtokens"]], dim=1)
        assert got.shape == ref.shape
        err = (got - ref).abs().max().item() / ref.abs().max().item()
        assert err < 2e-6, err            # HF builds its sin / cos tables in float32 even for a float64 model


def test_oracle_swiglu_block_matches_huggingface_gated_mlp():
    """Oracle-only groundwork for SURVEY §8f.1: SwiGLU FFN  w3(silu(w1 x) * w2 x)  (layers/ffn_layers.py:52-76, hidden =
    2/3 * 4D rounded up to `align_to`) inside the whole ViT against HF's gated-MLP DINOv3 (gate_proj / up_proj / down_proj
    with SiLU)."""
    from oracle.arch import ModelCfg
    from oracle.model import backbone_forward, formula_images, formula_params, sub
    D, depth, heads, size = 128, 2, 2, 48
    cfg = ModelCfg(embed_dim=D, depth=depth, heads=heads, global_size=size, local_size=32, n_storage=4, ln_eps=1e-5,
                   ffn_layer="swiglu", swiglu_align=64, n_prototypes=16, head_hidden=16, head_bottleneck=8)
    assert cfg.swiglu_hidden == 384                                     # int(512 * 2 / 3) = 341 -> 384
    bp = sub(formula_params(cfg, 8), "student_backbone")
    hcfg = hf.DINOv3ViTConfig(patch_size=16, hidden_size=D, intermediate_size=cfg.swiglu_hidden, num_hidden_layers=depth,
                              num_attention_heads=heads, hidden_act="silu", layer_norm_eps=1e-5, rope_theta=100.0, image_size=size,
                              query_bias=True, key_bias=True, value_bias=True, num_register_tokens=4, use_gated_mlp=True)
    model = hf.DINOv3ViTModel(hcfg).double().eval()
    mlp_free = {k: v for k, v in bp.items() if "/mlp/" not in k}
    for i in range(depth):                                              # placeholders so the shared mapper can run
        mlp_free[f"blocks_{i}/mlp/Dense_0/kernel"] = torch.zeros(D, 1, dtype=torch.float64)
        mlp_free[f"blocks_{i}/mlp/Dense_0/bias"] = torch.zeros(1, dtype=torch.float64)
        mlp_free[f"blocks_{i}/mlp/Dense_1/kernel"] = torch.zeros(1, D, dtype=torch.float64)
        mlp_free[f"blocks_{i}/mlp/Dense_1/bias"] = torch.zeros(D, dtype=torch.float64)
    sd = {k: v for k, v in _to_hf_state_dict(mlp_free, depth, D).items() if ".mlp." not in k}
    for i in range(depth):
        b, h = f"blocks_{i}/mlp/", f"model.layer.{i}.mlp."
        for ours, theirs in (("w1", "gate_proj"), ("w2", "up_proj"), ("w3", "down_proj")):
            sd[h + theirs + ".weight"] = bp[b + ours + "/kernel"].t().contiguous()
            sd[h + theirs + ".bias"] = bp[b + ours + "/bias"]
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected and all("inv_freq" in k for k in missing), (missing, unexpected)
    x = formula_images((2, size, size, 3), 78)
    want = backbone_forward(bp, [x], [None], cfg)[0]
    with torch.no_grad():
        got = model(pixel_values=x.permute(0, 3, 1, 2).contiguous()).last_hidden_state
    ref = torch.cat([want["x_norm_clstoken"][:, None], want["x_storage_tokens"], want["x_norm_patchtokens"]], dim=1)
    assert (got - ref).abs().max().item() / ref.abs().max().item() < 2e-6
