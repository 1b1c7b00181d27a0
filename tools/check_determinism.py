"""1 GPU: run-to-run reproducibility of the step's gradients (same engine thrice, then a second engine instance)."""
import os, sys, dataclasses
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "dinov3-jax_b200")); sys.path.insert(0, ROOT)
import torch
from dinov3_jax import _native
from dinov3_jax.engine import Engine, config_for
from dinov3_jax.engine.synth import reference_like_params, synthetic_batch
_native.init(0)
arch = sys.argv[1] if len(sys.argv) > 1 else "vit_large"
depth = int(os.environ.get("CHK_DEPTH", "2")); B = int(os.environ.get("CHK_B", "4"))
cfg = dataclasses.replace(config_for(arch, n_prototypes=8192, layerscale=0.1), depth=depth)
params = reference_like_params(cfg, 0)
batch = synthetic_batch(cfg, B, seed=10)


def rel(a, b):
    num = sum(float(((a[k] - b[k]) ** 2).sum()) for k in a); den = sum(float((a[k] ** 2).sum()) for k in a)
    w = max(((float((a[k] - b[k]).norm() / (a[k].norm() + 1e-30)), k) for k in a))
    return (num / den) ** 0.5, w


runs = []
for inst in range(2):
    eng = Engine(cfg, B, max_masked=int(batch["mask_indices_list"].shape[0]))
    eng.params.load_reference_tree(params)
    eng.set_batch(batch)
    for it in range(3 if inst == 0 else 1):
        eng.forward_backward(0.05)
        torch.cuda.synchronize()
        runs.append({k: v.float().cpu() for k, v in eng.params.export_reference_tree("grad").items()})
    del eng
print("same engine, run 0 vs 1:", rel(runs[0], runs[1]))
print("same engine, run 1 vs 2:", rel(runs[1], runs[2]))
print("second engine instance vs run 0:", rel(runs[0], runs[3]))
# which tensors differ first (walk from the heads down)
for k in ("student_dino_head/last_layer/kernel", "student_dino_head/mlp/layers_0/kernel", "student_ibot_head/mlp/layers_0/kernel",
          "student_backbone/norm/scale", f"student_backbone/blocks_{depth - 1}/mlp/Dense_1/kernel", f"student_backbone/blocks_{depth - 1}/mlp/Dense_0/kernel",
          f"student_backbone/blocks_{depth - 1}/attn/proj/kernel", f"student_backbone/blocks_{depth - 1}/attn/qkv/kernel",
          f"student_backbone/blocks_{depth - 1}/norm1/scale", "student_backbone/blocks_0/attn/qkv/kernel"):
    a, b = runs[0][k], runs[1][k]
    print(f"  {k:60s} {float((a - b).norm() / (a.norm() + 1e-30)):.3e}")

# ---- where does the run-to-run difference enter?  Snapshot intermediate buffers of two consecutive runs.
eng = Engine(cfg, B, max_masked=int(batch["mask_indices_list"].shape[0]))
eng.params.load_reference_tree(params)
eng.set_batch(batch)
snaps = []
for it in range(2):
    eng.forward_backward(0.05)
    torch.cuda.synchronize()
    M = eng.M
    snaps.append({
        "t_ibot_logits": eng.h_t_ibot.logits[:M].clone(), "sk_ibot.mx": eng.sk_ibot.mx.clone(), "sk_ibot.s": eng.sk_ibot.s.clone(),
        "sk_ibot.a": eng.sk_ibot.a[:M].clone(), "s_ibot_logits": eng.h_s_ibot.logits[:M].clone(), "ibot dS": eng.h_s_ibot.dS[:M].float().clone(),
        "ibot dYn": eng.h_s_ibot.dYn[:M].float().clone(), "ibot dU3": eng.h_s_ibot.dU3[:M].float().clone(), "ibot dA0": eng.h_s_ibot.dA0[:M].clone(),
        "dino dS": eng.h_s_dino.dS.float().clone(), "dino dA0": eng.h_s_dino.dA0.clone(), "student Xn": eng.student.Xn.clone(),
        "student QKV[0]": eng.student.QKV[0].float().clone(), "student O[0]": eng.student.O[0].float().clone(),
        "metrics": eng.metrics.clone(),
    })
for k in snaps[0]:
    a, b = snaps[0][k], snaps[1][k]
    d = float((a - b).norm() / (a.norm() + 1e-30))
    print(f"  {k:20s} rel diff {d:.3e}   max abs {float((a - b).abs().max()):.3e}")
