"""GPU probe: engine forward/backward/optimizer vs the CPU oracle on a tiny configuration."""
import os, sys, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "dinov3-jax_b200")); sys.path.insert(0, ROOT)
import torch
from oracle import tiny_cfg, cfg_for
from oracle.model import init_params, Emu
from oracle.batch import synthetic_batch
from oracle.step import ssl_forward, train_step, init_opt_state
from dinov3_jax.engine import Engine, from_oracle_cfg


def run(cfg, B, label, perturb=0.05, check_step=True, oracle_device="cpu"):
    print(f"== {label}: D={cfg.embed_dim} L={cfg.depth} K={cfg.n_prototypes} B={B} ls={cfg.layerscale}", flush=True)
    P = init_params(cfg, 0, perturb=perturb)
    batch = synthetic_batch(cfg, B, 0)
    eng = Engine(from_oracle_cfg(cfg), B, max_masked=max(int(batch["mask_indices_list"].shape[0]), 1))
    eng.params.load_reference_tree(P)
    hyper = dict(lr=1e-3, wd=0.04, last_layer_lr=5e-4, momentum=0.99, teacher_temp=0.05)
    eng.set_batch(batch)
    t0 = time.time()
    eng.forward_backward(hyper["teacher_temp"]); torch.cuda.synchronize()
    print(f"  engine fwd+bwd {time.time()-t0:.3f}s", flush=True)
    grads_e = {k: v.cpu() for k, v in eng.params.export_reference_tree("grad").items()}
    eng.optimizer_step(hyper["lr"], hyper["wd"], hyper["last_layer_lr"], hyper["momentum"]); torch.cuda.synchronize()
    met = eng.read_metrics()
    newp_e = {k: v.cpu() for k, v in eng.params.export_reference_tree("param").items()}
    od = torch.device(oracle_device)
    Po = {k: v.to(od) for k, v in P.items()}; bo = {k: (v.to(od) if torch.is_tensor(v) else v) for k, v in batch.items()}
    for name, emu in (("fp32 oracle", Emu(False)), ("bf16-emulating oracle", Emu(True))):
        t0 = time.time()
        newp, st, loss, m, grads = train_step(Po, init_opt_state(Po), bo, cfg, emu=emu, **hyper)
        print(f"  -- vs {name} ({time.time()-t0:.1f}s): loss oracle={loss.item():.6f} engine={met['total_loss']:.6f} rel={abs(loss.item()-met['total_loss'])/abs(loss.item()):.2e}")
        for k in ("dino_local_crops_loss", "dino_global_crops_loss", "koleo_loss", "ibot_loss", "student_backbone_grad_norm", "student_dino_head_grad_norm", "student_ibot_head_grad_norm"):
            o = float(m[k]); e = met[k]
            print(f"     {k}: oracle={o:.6f} engine={e:.6f} rel={abs(o-e)/(abs(o)+1e-12):.2e}")
        worst = []
        for k, g in grads.items():
            g = g.cpu(); ge = grads_e[k].reshape(g.shape)
            r = ((ge - g).norm() / (g.norm() + 1e-30)).item()
            worst.append((r, k, g.norm().item()))
        worst.sort(reverse=True)
        tot = torch.sqrt(sum(((grads_e[k].reshape(g.shape) - g.cpu()) ** 2).sum() for k, g in grads.items())) / torch.sqrt(sum((g.cpu() ** 2).sum() for g in grads.values()))
        print(f"     grads: global rel={tot.item():.3e}; worst tensors:")
        for r, k, n in worst[:6]:
            print(f"       {k}: rel={r:.3e} |g|={n:.3e}")
        if check_step:
            wp = []
            for k, v in newp.items():
                v = v.cpu(); d0 = (v - P[k]).norm().item()
                r = ((newp_e[k].reshape(v.shape) - v).norm() / (d0 + 1e-30)).item()   # error relative to the update size
                wp.append((r, k, d0))
            wp.sort(reverse=True)
            print("     update error / update norm, worst:", ", ".join(f"{k.split('/',1)[1][-28:]}={r:.2e}" for r, k, _ in wp[:4]))


if __name__ == "__main__":
    torch.manual_seed(0)
    which = sys.argv[1:] or ["tiny", "tiny_ls1"]
    if "tiny" in which:
        run(tiny_cfg(), 4, "tiny")
    if "tiny_ls1" in which:
        run(tiny_cfg(layerscale=1.0), 4, "tiny layerscale=1")
    if "vits" in which:
        run(cfg_for("vit_small", n_prototypes=4096), 4, "vit_small K=4096", oracle_device="cuda")
