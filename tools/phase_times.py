"""GPU: time the phases of one ViT-L/16 B=64 step with CUDA events (teacher fwd, student fwd, heads+losses, backward, optimizer)."""
import os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "dinov3-jax_b200"))
import torch
from dinov3_jax.engine import Engine, config_for
from dinov3_jax.engine.synth import synthetic_batch, init_reference_like
from dinov3_jax import ops

arch = sys.argv[1] if len(sys.argv) > 1 else "vit_large"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
cfg = config_for(arch)
batch = synthetic_batch(cfg, B, 0)
eng = Engine(cfg, B, max_masked=int(batch["mask_indices_list"].shape[0]))
init_reference_like(eng)
eng.set_batch(batch)
hyper = dict(teacher_temp=0.04, lr=1e-4, wd=0.04, last_layer_lr=0.0, momentum=0.996)
for _ in range(3): eng.train_step(None, **hyper)
marks = []
def mark(name):
    e = torch.cuda.Event(enable_timing=True); e.record(); marks.append((name, e))
orig = {k: getattr(eng, k) for k in ("_backbone_fwd", "_head_fwd", "_sinkhorn_pair", "_head_bwd", "_block_bwd", "optimizer_step")}
def wrap(name, label_fn):
    f = orig[name]
    def g(*a, **k):
        r = f(*a, **k); mark(label_fn(*a, **k)); return r
    setattr(eng, name, g)
wrap("_backbone_fwd", lambda st, imgs, masks, teacher: "backbone fwd teacher" if teacher else "backbone fwd student")
wrap("_head_fwd", lambda hb, module, R, teacher, stash: f"heads fwd {'teacher' if teacher else 'student'}")
wrap("_sinkhorn_pair", lambda *a, **k: "sinkhorn")
wrap("_head_bwd", lambda *a, **k: "heads bwd (+CE, before)")
wrap("_block_bwd", lambda i, *a: "blocks bwd")
wrap("optimizer_step", lambda *a, **k: "optimizer (sumsq+adamw+ema)")
mark("start")
eng.train_step(None, **hyper)
mark("end")
torch.cuda.synchronize()
agg = {}
for (n0, e0), (n1, e1) in zip(marks, marks[1:]):
    agg[n1] = agg.get(n1, 0.0) + e0.elapsed_time(e1)
tot = marks[0][1].elapsed_time(marks[-1][1])
print(f"{arch} B={B}: step {tot:.2f} ms")
for k, v in agg.items():
    print(f"  {k:34s} {v:8.2f} ms  {100*v/tot:5.1f}%")
