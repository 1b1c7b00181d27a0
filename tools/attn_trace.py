"""GPU: clock64() phase trace of one attention-backward CTA (N=197 and packed N=37)."""
import os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "dinov3-jax_b200"))
import torch
from dinov3_jax import ops, _native as N
lib = N.init()
for (n, Ntok, H) in [(128, 197, 16), (512, 37, 16)]:
    D = 64 * H
    qkv = torch.randn(n * Ntok, 3 * D, device="cuda").to(torch.bfloat16); do = torch.randn(n * Ntok, D, device="cuda").to(torch.bfloat16)
    o = torch.empty(n * Ntok, D, device="cuda", dtype=torch.bfloat16); lse = torch.zeros(n, H, Ntok, device="cuda"); delta = torch.zeros(n, H, Ntok, device="cuda")
    dqkv = torch.empty_like(qkv)
    ops.attn_fwd(qkv, o, lse, n, Ntok, D, H)
    for _ in range(2): ops.attn_bwd(qkv, o, do, lse, delta, dqkv, n, Ntok, D, H)
    buf = torch.zeros(64, dtype=torch.int64, device="cuda")
    N.check(lib.d3_debug_attn_trace(buf.data_ptr()))
    ops.attn_bwd(qkv, o, do, lse, delta, dqkv, n, Ntok, D, H)
    torch.cuda.synchronize()
    N.check(lib.d3_debug_attn_trace(None))
    t = buf.cpu().reshape(32, 2)
    t0 = int(t[0, 0])
    names = {0: "start", 1: "after tmem alloc+sync", 2: "Q/dO/KV landed (t0)", 20: "loop done", 21: "last acc done", 22: "dK/dV stored", 23: "dQ stored", 24: "exit"}
    for it in range(4):
        names.update({3 + 4 * it: f"it{it} before wait", 4 + 4 * it: f"it{it} S/dP ready", 5 + 4 * it: f"it{it} elementwise done", 6 + 4 * it: f"it{it} after sync"})
    print(f"== N={Ntok}: cycles since CTA start (thread 0 | thread 200)")
    for i in range(25):
        if int(t[i, 0]) or int(t[i, 1]):
            a = int(t[i, 0]) - t0 if int(t[i, 0]) else -1
            b = int(t[i, 1]) - t0 if int(t[i, 1]) else -1
            print(f"  {names.get(i, i):28s} {a:8d} {b:8d}")
