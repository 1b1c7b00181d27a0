"""Four tcgen05 GEMM launches of a ViT-L/16 B=64 student block with KNOWN shapes, for one `ncu --set full` capture
(`ncu --set full --clock-control none -k regex:gemm2sm -o gpurun_out/r02_ncu_gemm python tools/ncu_gemm_labeled.py`).
Launch order (each after one un-profiled warm-up of the same call): qkv forward, fc1 forward (+bias, GELU, stash), proj
forward (+bias, LayerScale, fp32 residual), fc2 weight gradient.  tools/ncu_summary.py labels the capture with these."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "dinov3-jax_b200"))
import torch
from dinov3_jax import ops
dev, bf = "cuda", torch.bfloat16
T, D, Hd = 44160, 1024, 4096
X = torch.randn(T, D, device=dev).to(bf); Hh = torch.randn(T, Hd, device=dev).to(bf)
Wqkv = torch.randn(D, 3 * D, device=dev).to(bf); W1 = torch.randn(D, Hd, device=dev).to(bf); Wp = torch.randn(D, D, device=dev).to(bf)
b3, b1, bd, gam = torch.randn(3 * D, device=dev), torch.randn(Hd, device=dev), torch.randn(D, device=dev), torch.randn(D, device=dev)
QKV = torch.empty(T, 3 * D, device=dev, dtype=bf); U1 = torch.empty(T, Hd, device=dev, dtype=bf); G = torch.empty(T, Hd, device=dev, dtype=bf)
Xf = torch.randn(T, D, device=dev); Xo = torch.empty(T, D, device=dev); dU2 = torch.randn(T, D, device=dev).to(bf)
gW2 = torch.zeros(Hd, D, device=dev)
calls = [
    lambda: ops.gemm(X, Wqkv, QKV, b_mn=True, bias=b3),
    lambda: ops.gemm(X, W1, G, b_mn=True, bias=b1, gelu=True, store_pre=U1),
    lambda: ops.gemm(X, Wp, Xo, b_mn=True, bias=bd, gamma=gam, resid=Xf),
    lambda: ops.gemm(Hh, dU2, gW2, a_mn=True, b_mn=True, accum=True),
]
for c in calls:
    c()
torch.cuda.synchronize()
for c in calls:
    c()
torch.cuda.synchronize()
print("done")
