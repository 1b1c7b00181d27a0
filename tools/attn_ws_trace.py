"""GPU: clock64() phase trace of CTA 0 of the warp-specialised attention forward (first 16 query tiles it processes)."""
import os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "dinov3-jax_b200"))
import torch
from dinov3_jax import ops, _native as N
lib = N.init()
EV = ["S ready", "S issued", "PV ready", "PV issued", "S seen", "max done", "P arrived", "O seen", "slot freed", "stored", "1st ld", "it0", "it1", "it2", "it3"]
for (n, Ntok, H) in [(128, 197, 16), (512, 37, 16)]:
    D = 64 * H
    qkv = torch.randn(n * Ntok, 3 * D, device="cuda").to(torch.bfloat16)
    o = torch.empty(n * Ntok, D, device="cuda", dtype=torch.bfloat16); lse = torch.zeros(n, H, Ntok, device="cuda")
    for _ in range(2): ops.attn_fwd(qkv, o, lse, n, Ntok, D, H)
    buf = torch.zeros(64 + 320, dtype=torch.int64, device="cuda")
    N.check(lib.d3_debug_attn_trace(buf.data_ptr()))
    ops.attn_fwd(qkv, o, lse, n, Ntok, D, H)
    torch.cuda.synchronize()
    N.check(lib.d3_debug_attn_trace(None))
    t = buf[64:].cpu().reshape(16, 20)[:, :15]
    t0 = int(t[t > 0].min())
    print(f"== N={Ntok}: cycles since the first event of CTA 0; rows = units (query tiles) in processing order")
    print("  unit " + " ".join(f"{e:>10s}" for e in EV))
    for u in range(16):
        print(f"  {u:4d} " + " ".join(f"{(int(x) - t0) if int(x) else -1:10d}" for x in t[u]))
