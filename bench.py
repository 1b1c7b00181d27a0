#!/usr/bin/env python
"""bench.py — DINOv3 SSL training step on B200: global-crops/sec, ViT-L/16, 224^2, 2 global + 8 local crops, bf16.

Contract (driver): `python bench.py --gpus N --steps K --warmup W` (torchrun launches N ranks for N > 1) prints ONE
JSON line on rank 0.  `--impl reference` times the CPU restatement of the reference step (oracle/, "port") on the
host cores instead (the reference's JAX stack is not installable here).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "dinov3-jax_b200"))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

ARCH_FLOPS = {}  # filled by block_flops()


def f_img(D, L, Ng, Nl, n_local=8):
    """Algorithmic attention+MLP FLOPs per image: teacher fwd on 2 global crops + 3x student."""
    f_blk = lambda N: 24 * N * D * D + 4 * N * N * D
    return L * (2 * f_blk(Ng) + 3 * (2 * f_blk(Ng) + n_local * f_blk(Nl)))


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), d.get("hbm_gbs", 6650.0), "measured"
    return 1590.0, 1400.0, 6650.0, "fallback"


def ncu_traffic():
    """DRAM traffic of the roofline kernel from the committed `ncu --set full` capture of GEMM launches with KNOWN shapes
    (profiles/r02_ncu_gemm_labeled.txt, tools/ncu_gemm_labeled.py): per labelled launch the measured dram read+write
    bytes, the algorithmic bytes and their ratio; `traffic` in the JSON line is the first one (the qkv projection)."""
    import re
    p = os.path.join(ROOT, "profiles", "r02_ncu_gemm_labeled.txt")
    if not os.path.exists(p):
        return None
    launches = []
    for line in open(p):
        m = re.match(r"\[(.*?)\] launch \d+: .*? = +([\d.]+) MB .*algorithmic ([\d.]+) MB -> traffic ratio ([\d.]+)", line)
        if m:
            launches.append({"launch": m.group(1), "dram_bytes": float(m.group(2)) * 1e6, "algorithmic_bytes": float(m.group(3)) * 1e6,
                             "traffic_ratio": float(m.group(4))})
    if not launches:
        return None
    return {"bytes_per_launch": launches[0]["dram_bytes"], "labelled_launches": launches,
            "source": "profiles/r02_ncu_gemm_labeled.txt (ncu --set full --clock-control none, tools/ncu_gemm_labeled.py)"}


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for t, line in self.rows:
            if t < t0 or t > t1 + 0.2:
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def log(msg):
    if int(os.environ.get("RANK", "0")) == 0:
        print(f"[bench {time.strftime('%H:%M:%S')}] {msg}", file=sys.stderr, flush=True)


def hyper(it):
    # ssl_default_config.yaml:124-128,142-146 at iteration `it` of the default schedule (values only matter for parity)
    return dict(teacher_temp=0.04, lr=1e-4, wd=0.04, last_layer_lr=0.0, momentum=0.996)


def cpu_reference_rate(arch, threads, sample_B, steps=1, warmup=0, **cfg_kw):
    """Times oracle.step.train_step (CPU restatement of the reference step) on `sample_B` images per step."""
    from oracle import cfg_for
    from oracle.batch import synthetic_batch
    from oracle.model import init_params
    from oracle.step import init_opt_state, train_step
    threads = max(1, min(threads, 32))   # beyond ~32 threads the fp32 restatement stops scaling (many small ops)
    torch.set_num_threads(threads)
    cfg = cfg_for(arch, **cfg_kw)
    P = init_params(cfg, 0)
    st = init_opt_state(P)
    batch = synthetic_batch(cfg, sample_B, 0)
    times, first_loss = [], None
    for i in range(warmup + steps):
        t0 = time.time()
        P, st, loss, m, _ = train_step(P, st, batch, cfg, **hyper(i))
        times.append(time.time() - t0)
        if first_loss is None:
            first_loss = float(loss)         # loss of the un-updated parameters (what parity_check compares)
    dt = sum(times[warmup:]) / steps
    return 2 * sample_B / dt, dt, float(loss), threads, first_loss


def oracle_kw(args):
    return dict(n_prototypes=args.prototypes, patch=args.patch, local_size=args.local_size)


def parity_check(args, cfg, loss_cpu, local_rank):
    """Engine vs oracle on the cpu_baseline sample: same architecture, K, parameters (oracle.init_params seed 0) and
    batch (oracle.batch.synthetic_batch seed 0) — the loss of the un-updated parameters, relative difference."""
    from dinov3_jax.engine import Engine
    from oracle import cfg_for
    from oracle.batch import synthetic_batch as oracle_batch
    from oracle.model import init_params
    ocfg = cfg_for(args.arch, **oracle_kw(args))
    B = args.cpu_sample_batch
    batch = oracle_batch(ocfg, B, 0)
    eng = Engine(cfg, B, device=f"cuda:{local_rank}", max_masked=max(int(batch["mask_indices_list"].shape[0]), 1))
    eng.params.load_reference_tree(init_params(ocfg, 0))
    eng.set_batch(batch)
    eng.forward_backward(hyper(0)["teacher_temp"])
    loss = eng.read_metrics()["total_loss"]
    rel = abs(loss - loss_cpu) / abs(loss_cpu)
    return {"engine_loss": loss, "oracle_loss": loss_cpu, "rel": rel, "ok": bool(rel < 1e-3),
            "what": f"{args.arch} full depth, K={cfg.n_prototypes}, {B} images, fp32 oracle vs bf16 engine, tolerance 1e-3"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--arch", default="vit_large")
    ap.add_argument("--batch", type=int, default=64, help="images per GPU (ssl_default_config.yaml:75)")
    ap.add_argument("--prototypes", type=int, default=65536)
    ap.add_argument("--patch", type=int, default=16)
    ap.add_argument("--local-size", type=int, default=96, help="98 for patch 14 (96 is not divisible, layers/patch_embed.py:48-49)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--gram", choices=["off", "ema", "frozen"], default="off",
                    help="add the Gram-anchoring term (gram.use_loss; SURVEY 8f.2): ema = EMA teacher's patches as targets, "
                         "frozen = a third backbone pass with a snapshot of the teacher; off = the BASELINE headline workload")
    ap.add_argument("--remat", action="store_true", help="activation rematerialisation (train.checkpointing): BASELINE configs[4]")
    ap.add_argument("--cpu-sample-batch", type=int, default=2, help="images per CPU-baseline step (bounded sample: ~15-25 s of CPU work)")
    ap.add_argument("--cpu-budget-s", type=float, default=150.0, help="wall-clock bound of the --impl reference arm")
    ap.add_argument("--no-checks", action="store_true", help="skip parity_check / fsdp_check")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cores = os.cpu_count() or 1

    from dinov3_jax.engine.config import ARCHS
    D, L, H = ARCHS[args.arch]
    Ng, Nl = (224 // args.patch) ** 2 + 1, (args.local_size // args.patch) ** 2 + 1
    flops_per_gcrop = f_img(D, L, Ng, Nl) / 2
    cfg_desc = {"workload": f"{args.arch}/{args.patch} student+teacher, 2x224^2 + 8x{args.local_size}^2 crops, {args.batch} img/GPU, "
                            f"K={args.prototypes} prototypes, DINO+iBOT+KoLeo, clip+AdamW+EMA (BASELINE configs[3] per-GPU shape)",
                "global_batch": args.batch * world, "parallelism": f"fsdp{world}" if world > 1 else "single",
                "l2": "inputs+activations per step >> 126 MB L2 (no flush needed)"}

    if args.impl == "reference":
        if rank != 0:
            return
        # The reference's JAX stack cannot be installed offline, so this arm times the CPU restatement of
        # the same step (oracle/, kind "port") on a BOUNDED SAMPLE of the workload: `--cpu-sample-batch` images per
        # step instead of 64.  Everything printed describes what actually ran.
        budget = float(args.cpu_budget_s)
        t_probe = time.time()
        _, dt1, _, used, _ = cpu_reference_rate(args.arch, cores, args.cpu_sample_batch, steps=1, warmup=0, **oracle_kw(args))   # untimed probe = warm-up 1
        steps = max(1, min(args.steps, int((budget - (time.time() - t_probe)) / max(dt1, 1e-3)) - max(args.warmup - 1, 0)))
        warm = max(0, min(args.warmup - 1, int(budget / max(dt1, 1e-3)) - steps))
        val, dt, _, used, _ = cpu_reference_rate(args.arch, cores, args.cpu_sample_batch, steps=steps, warmup=warm, **oracle_kw(args))
        ref_cfg = dict(cfg_desc)
        ref_cfg.update({"sample": f"{args.cpu_sample_batch} images per step (bounded sample of the {args.batch}-img/GPU step; "
                                  f"crops/s on the CPU is taken as batch-size independent)",
                        "global_batch": args.cpu_sample_batch, "parallelism": "single host process", "same_config": False,
                        "extrapolated": True, "requested": {"gpus": args.gpus, "steps": args.steps, "warmup": args.warmup}})
        print(json.dumps({
            "impl": "reference", "metric": "global_crops_per_sec", "value": val, "unit": "global-crops/s",
            "n_gpus": 1, "steps": steps, "warmup": warm + 1, "ms_per_step": dt * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": ref_cfg,
            "cpu_baseline": {"value": val, "unit": "global-crops/s", "cores": used, "kind": "port",
                             "sample": f"torch-CPU fp32 restatement of the reference train_step (oracle/), {args.arch}, "
                                       f"{args.cpu_sample_batch} images/step, {steps} timed steps after {warm + 1} warm-up; "
                                       f"the reference's JAX stack is not installable offline"},
            "e2e": {"value": val, "unit": "global-crops/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the B200 engine has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    comm = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        from dinov3_jax.fsdp.runtime import Comm
        comm = Comm(dist.group.WORLD)
    from dinov3_jax import _native, ops
    from dinov3_jax.engine import Engine, config_for
    from dinov3_jax.engine.synth import synthetic_batch, init_reference_like
    _native.init(local_rank)
    cfg = config_for(args.arch, n_prototypes=args.prototypes, patch=args.patch, local_size=args.local_size)
    if args.gram != "off":
        import dataclasses
        cfg = dataclasses.replace(cfg, gram_use_loss=True, gram_ema_teacher=args.gram == "ema", gram_it_load_ema_teacher=0,
                                  gram_remove_only_teacher_neg=True)
        cfg_desc["workload"] += f" + Gram anchoring ({args.gram} teacher)"
        cfg_desc["gram"] = args.gram
    B = args.batch
    log(f"building synthetic batch B={B}")
    batch = synthetic_batch(cfg, B, seed=rank, pin=True)
    M = int(batch["mask_indices_list"].shape[0])
    eng = Engine(cfg, B, device=f"cuda:{local_rank}", max_masked=M, comm=comm, remat=args.remat)
    cfg_desc["activation_remat"] = bool(args.remat)
    if world > 1:
        cfg_desc["grad_reduce_scatter"] = ("push over NVLink peer memory (GEMM epilogue + d3_scatter_add_peers)" if eng.fsdp.push else "nccl reduce_scatter")
    cfg_desc["wgrad_stream"] = bool(eng.wgrad_overlap)
    log(f"engine allocated ({torch.cuda.memory_allocated() / 2**30:.1f} GiB); initialising {eng.params.n_params() / 1e6:.1f} M student params")
    init_reference_like(eng, seed=0)
    log("params loaded")

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def step_device(i):
        eng.train_step(None, **hyper(i))

    def step_e2e(i):
        eng.train_step(batch, **hyper(i))     # pinned host -> device copies inside
        return eng.read_metrics()["total_loss"]  # device -> host read of the loss

    # ---- warm-up (device-resident inputs)
    eng.set_batch(batch)
    for i in range(max(args.warmup, 3)):
        tw = time.time()
        step_device(i)
        torch.cuda.synchronize()
        log(f"warm-up step {i}: {time.time() - tw:.3f} s")
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    _native.reset_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    e0.record()
    for i in range(args.steps):
        step_device(i)
    e1.record()
    barrier()
    t1 = time.time()
    launches = _native.launch_count()
    ms = e0.elapsed_time(e1) / args.steps
    log(f"timed region: {ms:.2f} ms/step")
    clocks = sampler.stop(t0, t1) if sampler else None
    # ---- end-to-end: host buffers, H2D inside, loss read back every step
    for i in range(2):
        step_e2e(i)
    barrier()
    e0.record()
    loss = 0.0
    for i in range(args.steps):
        loss = step_e2e(i)
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1) / args.steps
    log(f"e2e: {ms_e2e:.2f} ms/step")
    if world > 1:
        t = torch.tensor([ms, ms_e2e], device="cuda")
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms, ms_e2e = t.tolist()
    h2d = sum(batch[k].numel() * batch[k].element_size() for k in ("collated_global_crops", "collated_local_crops", "collated_masks", "mask_indices_list"))
    # ---- roofline leg: one instrumented step, CUDA events around every tensor-core GEMM launch
    # (single stream for this step: with the weight-gradient stream active a launch's event pair would also time its
    # wait for SMs held by the other stream's kernel)
    overlap, eng.wgrad_overlap = eng.wgrad_overlap, False
    fwd_overlap, eng.fwd_overlap = eng.fwd_overlap, False
    ops.PROFILE = []
    step_device(0)
    torch.cuda.synchronize()
    prof, ops.PROFILE = ops.PROFILE, None
    eng.wgrad_overlap = overlap
    eng.fwd_overlap = fwd_overlap
    g_flops = sum(p[1] for p in prof)
    g_ms = sum(p[2].elapsed_time(p[3]) for p in prof)
    burst, sustained, hbm, how = peaks()
    gemm_tf = g_flops / (g_ms * 1e-3) / 1e12 if g_ms else 0.0
    traffic = ncu_traffic()
    value = 2 * B * world / (ms * 1e-3)
    e2e_value = 2 * B * world / (ms_e2e * 1e-3)
    step_tf = value * flops_per_gcrop / world / 1e12
    out = {
        "metric": "global_crops_per_sec", "value": value, "unit": "global-crops/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic", "config": cfg_desc, "clocks": clocks, "gpu_launches": int(launches),
        "e2e": {"value": e2e_value, "unit": "global-crops/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 32,
                "ms_per_step": ms_e2e, "loss": loss},
        "roofline": {"bound": "tensor", "kernel": "gemm2sm_kernel + gemm1sm_kernel (every tcgen05 GEMM launch of one step)",
                     "achieved": gemm_tf, "peak": sustained, "unit": "TFLOP/s", "frac": gemm_tf / sustained,
                     "traffic": traffic["bytes_per_launch"] if traffic else None,
                     "traffic_unit": "bytes of ONE labelled launch (dram__bytes_read.sum + dram__bytes_write.sum, ncu --set full): "
                                     "labelled_launches[0], the qkv projection",
                     "traffic_labelled_launches": traffic["labelled_launches"] if traffic else None,
                     "traffic_source": traffic["source"] if traffic else None, "peak_source": f"bf16_tflops_sustained ({how})", "launches": len(prof),
                     "share_of_step": g_ms / ms if ms else None},
        "step_roofline": {"achieved": step_tf, "peak": sustained, "unit": "TFLOP/s", "frac": step_tf / sustained,
                          "note": "attention+MLP algorithmic FLOPs (BASELINE.md §3) / step time / GPU"},
    }
    if world > 1 and not args.no_checks:
        from dinov3_jax.fsdp.selfcheck import fsdp_equals_single_gpu
        try:
            out["fsdp_check"] = fsdp_equals_single_gpu(comm, f"cuda:{local_rank}")
        except Exception as e:                      # the timed numbers above stand; the check reports its own failure
            out["fsdp_check"] = {"ok": False, "error": repr(e)[:300]}
        log(f"fsdp_check: {out['fsdp_check']}")
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            log("cpu baseline (oracle port) ...")
            v, dt, _, cores, loss_cpu = cpu_reference_rate(args.arch, cores, args.cpu_sample_batch, steps=2, warmup=0, **oracle_kw(args))
            log(f"cpu baseline: {dt:.1f} s/step")
            out["cpu_baseline"] = {"value": v, "unit": "global-crops/s", "cores": cores, "kind": "port",
                                   "sample": f"oracle train_step (torch-CPU fp32 restatement), {args.arch}, "
                                             f"{args.cpu_sample_batch} images per step, 2 steps of {dt:.1f} s"}
            if not args.no_checks:
                # the same sample through the engine (same parameters, same batch): full-size model, K prototypes
                try:
                    out["parity_check"] = parity_check(args, cfg, loss_cpu, local_rank)
                except Exception as e:
                    out["parity_check"] = {"ok": False, "error": repr(e)[:300]}
                log(f"parity_check: {out['parity_check']}")
        print(json.dumps(out))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
