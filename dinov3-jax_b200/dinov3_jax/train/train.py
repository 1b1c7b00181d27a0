"""Training driver with the reference's entry points (dinov3_jax/train/train.py): `get_args_parser`, `main`,
`do_train`, `build_schedulers`, `build_optimizer`, `train_step`.

Launch one process per GPU:  torchrun --nproc-per-node N -m dinov3_jax.train.train --config-file cfg.yaml --opts k=v
The loop keeps the reference's shape (:622-706) minus its per-step host syncs: metrics are read every `print_freq`.
"""
from __future__ import annotations

import argparse
import math
import os
import time

import torch

from ..configs import DinoV3SetupArgs, setup_config
from .cosine_lr_scheduler import CosineScheduler
from .ssl_meta_arch import SSLMetaArch


def get_args_parser(add_help: bool = True):
    p = argparse.ArgumentParser("DINOv3 training (B200 engine)", add_help=add_help)
    p.add_argument("--config-file", default="", metavar="FILE")
    p.add_argument("--no-resume", action="store_true")
    p.add_argument("--eval-only", action="store_true")
    p.add_argument("--eval", type=str, default="")
    p.add_argument("--opts", default=[], nargs="+", help="key=value overrides")
    p.add_argument("--output-dir", default="", type=str)
    p.add_argument("--seed", default=0, type=int)
    p.add_argument("--max-iters", default=0, type=int, help="stop after this many iterations (0 = full schedule)")
    p.add_argument("--print-freq", default=10, type=int)
    return p


def build_schedulers(config):
    """train/train.py:127-182: lr, weight decay, teacher momentum, teacher temperature, last-layer lr."""
    L = config.train.OFFICIAL_EPOCH_LENGTH
    total = config.optim["epochs"] * L
    lr = dict(base_value=config.optim["lr"], final_value=config.optim["min_lr"], total_iters=total,
              warmup_iters=config.optim["warmup_epochs"] * L, start_warmup_value=0,
              trunc_extra=config.optim["schedule_trunc_extra"])
    wd = dict(base_value=config.optim["weight_decay"], final_value=config.optim["weight_decay_end"], total_iters=total,
              trunc_extra=config.optim["schedule_trunc_extra"])
    mom = dict(base_value=config.teacher["momentum_teacher"], final_value=config.teacher["final_momentum_teacher"],
               total_iters=total, trunc_extra=config.optim["schedule_trunc_extra"])
    tw = config.teacher["warmup_teacher_temp_epochs"] * L
    temp = dict(base_value=config.teacher["teacher_temp"], final_value=config.teacher["teacher_temp"], total_iters=tw,
                warmup_iters=tw, start_warmup_value=config.teacher["warmup_teacher_temp"])
    lr_s, wd_s, mom_s, temp_s, last_s = (CosineScheduler(**lr), CosineScheduler(**wd), CosineScheduler(**mom),
                                         CosineScheduler(**temp), CosineScheduler(**lr))
    last_s.schedule[: config.optim["freeze_last_layer_epochs"] * L] = 0     # :169-173
    return lr_s, wd_s, mom_s, temp_s, last_s


def build_optimizer(config, param_groups, lr_schedule=None, wd_schedule=None, last_layer_lr_schedule=None):
    """train/train.py:75-122 builds optax.multi_transform(adamw per group).  The B200 optimizer is the fused
    clip+AdamW+EMA kernel inside the engine; this returns the description it is driven by."""
    return {"groups": param_groups, "lr": lr_schedule, "wd": wd_schedule, "last_layer_lr": last_layer_lr_schedule,
            "b1": config.optim.adamw_beta1, "b2": config.optim.adamw_beta2, "clip_grad": config.optim.clip_grad}


def train_step(engine, batch, teacher_temp, iteration, schedules):
    """One optimisation step (train/train.py:491-565 with the intended update semantics)."""
    lr_s, wd_s, mom_s, _, last_s = schedules
    engine.train_step(batch, teacher_temp=float(teacher_temp), lr=float(lr_s[iteration]), wd=float(wd_s[iteration]),
                      last_layer_lr=float(last_s[iteration]), momentum=float(mom_s[iteration]))


def do_train(config, model: SSLMetaArch, resume: bool = False, data_loader=None, max_iters: int = 0,
             print_freq: int = 10):
    """train/train.py:319-713.  `data_loader` yields the reference's collate dicts; default: synthetic batches."""
    from .. import distributed
    from ..engine.synth import init_reference_like, synthetic_batch
    comm = None
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if distributed.is_enabled() and distributed.get_world_size() > 1:
        from ..fsdp.runtime import Comm
        comm = Comm()
    engine = model.build_engine(device=f"cuda:{local_rank}", comm=comm)
    init_reference_like(engine, seed=config.train.seed)
    schedules = build_schedulers(config)
    total = len(schedules[0].schedule)
    n_iters = min(total, max_iters) if max_iters else total
    if data_loader is None:
        fixed = synthetic_batch(engine.cfg, config.train.batch_size_per_gpu, seed=distributed.get_rank(), pin=True)
        data_loader = (fixed for _ in range(n_iters))
    # ---- resume / periodic checkpoints (train/train.py:447-469,695-706; <output_dir>/ckpt/<iteration>)
    from ..checkpointer import (engine_state, find_latest_checkpoint, keep_checkpoint_copy, keep_last_n_checkpoints,
                                load_checkpoint, load_engine_state, save_checkpoint)
    ckpt_dir = os.path.join(getattr(config.train, "output_dir", None) or ".", "ckpt")
    ck_cfg = config.get("checkpointing", None) if hasattr(config, "get") else getattr(config, "checkpointing", None)
    start_iter = 0
    last = find_latest_checkpoint(ckpt_dir) if resume else None
    if last is not None:
        ck = load_checkpoint(last, strict_loading=False)
        load_engine_state(engine, ck["model_params"], ck.get("optimizer_state"))
        start_iter = int(ck["iteration"]) + 1
        if distributed.is_main_process():
            print(f"checkpoint found {last}: resuming at iteration {start_iter}", flush=True)
    meters, nan_streak, t0 = {}, 0, time.time()
    for it, data in zip(range(start_iter, n_iters), data_loader):
        train_step(engine, data, schedules[3][it], it, schedules)
        if ck_cfg is not None and (it + 1) % int(ck_cfg.period) == 0:
            params_tree, opt_tree = engine_state(engine)                 # collective under FSDP (all-gathers the shards)
            if distributed.is_main_process():
                save_checkpoint(os.path.join(ckpt_dir, str(it)), iteration=it, params=params_tree,
                                optimizer_state=opt_tree, overwrite=True)
                keep_last_n_checkpoints(ckpt_dir, ck_cfg.max_to_keep)
                if "keep_every" in ck_cfg and (it + 1) % int(ck_cfg.keep_every) == 0:
                    keep_checkpoint_copy(os.path.join(ckpt_dir, str(it)))
        if it % print_freq == 0 or it == n_iters - 1:
            m = engine.read_metrics()                  # the only device->host sync of the loop
            if math.isnan(m["total_loss"]):            # NaN guard of train/train.py:656-667, evaluated on read
                nan_streak += 1
                if nan_streak > 2:
                    raise RuntimeError("loss is NaN for more than 2 consecutive reads")
            else:
                nan_streak = 0
            meters = m
            if distributed.is_main_process():
                dt = (time.time() - t0) / (it + 1)
                print(f"it {it}: loss {m['total_loss']:.4f} dino_l {m['dino_local_crops_loss']:.4f} dino_g "
                      f"{m['dino_global_crops_loss']:.4f} koleo {m['koleo_loss']:.4f} ibot {m['ibot_loss']:.4f} "
                      f"({dt * 1e3:.1f} ms/it)", flush=True)
    return meters


def main(argv=None):
    args = get_args_parser().parse_args(argv)
    if "RANK" in os.environ and int(os.environ.get("WORLD_SIZE", "1")) > 1:
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl")
    config = setup_config(DinoV3SetupArgs(config_file=args.config_file or None, output_dir=args.output_dir, opts=args.opts))
    model = SSLMetaArch(config)
    return do_train(config, model, resume=not args.no_resume, max_iters=args.max_iters, print_freq=args.print_freq)


if __name__ == "__main__":
    main()
