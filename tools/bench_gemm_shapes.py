"""GPU microbench: every GEMM call shape of one ViT-L/16 B=64 training step, with its real epilogue, timed in isolation."""
import os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "dinov3-jax_b200"))
import torch
from dinov3_jax import ops

dev = "cuda"
bf, f32 = torch.bfloat16, torch.float32
D, Hd = 1024, 4096
Tt, Ts = 25216, 44160


def timeit(fn, iters=8):
    for _ in range(3): fn()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def run(tile_n=0, only=None):
    T = Ts
    X = torch.randn(T, D, device=dev).to(bf); H = torch.randn(T, Hd, device=dev).to(bf)
    Wqkv = torch.randn(D, 3 * D, device=dev).to(bf); Wp = torch.randn(D, D, device=dev).to(bf)
    W1 = torch.randn(D, Hd, device=dev).to(bf); W2 = torch.randn(Hd, D, device=dev).to(bf)
    b3, b1, bd = torch.randn(3 * D, device=dev), torch.randn(Hd, device=dev), torch.randn(D, device=dev)
    gam = torch.randn(D, device=dev)
    QKV = torch.empty(T, 3 * D, device=dev, dtype=bf); Xf = torch.randn(T, D, device=dev); Xo = torch.empty(T, D, device=dev)
    P = torch.empty(T, D, device=dev, dtype=bf); U1 = torch.empty(T, Hd, device=dev, dtype=bf); Hh = torch.empty(T, Hd, device=dev, dtype=bf)
    dU2 = torch.randn(T, D, device=dev).to(bf); dU1 = torch.empty(T, Hd, device=dev, dtype=bf); dQKV = torch.randn(T, 3 * D, device=dev).to(bf)
    dZ = torch.empty(T, D, device=dev, dtype=bf)
    gW2 = torch.empty(Hd, D, device=dev); gW1 = torch.empty(D, Hd, device=dev); gWqkv = torch.empty(D, 3 * D, device=dev); gWp = torch.empty(D, D, device=dev)
    cases = [
        ("fwd qkv   [T,D]x[D,3D] +bias",            2 * T * D * 3 * D, lambda: ops.gemm(X, Wqkv, QKV, b_mn=True, bias=b3, tile_n=tile_n)),
        ("fwd proj  +bias,pre,gamma,resid(f32)",     2 * T * D * D,     lambda: ops.gemm(X, Wp, Xo, b_mn=True, bias=bd, store_pre=P, gamma=gam, resid=Xf, tile_n=tile_n)),
        ("fwd fc1   +bias,pre,gelu",                 2 * T * D * Hd,    lambda: ops.gemm(X, W1, Hh, b_mn=True, bias=b1, gelu=True, store_pre=U1, tile_n=tile_n)),
        ("fwd fc1   +bias,gelu (teacher)",           2 * T * D * Hd,    lambda: ops.gemm(X, W1, Hh, b_mn=True, bias=b1, gelu=True, tile_n=tile_n)),
        ("fwd fc2   +bias,pre,gelu,gamma,resid",     2 * T * D * Hd,    lambda: ops.gemm(H, W2, Xo, b_mn=True, bias=bd, gelu=True, store_pre=P, gamma=gam, resid=Xf, tile_n=tile_n)),
        ("dgrad fc2 [T,D]x[D,4D] *gelu'(u1)",        2 * T * D * Hd,    lambda: ops.gemm(dU2, W2, dU1, dgelu_of=U1, tile_n=tile_n)),
        ("dgrad fc1 [T,4D]x[4D,D]",                  2 * T * D * Hd,    lambda: ops.gemm(H, W1, dZ, tile_n=tile_n)),
        ("dgrad proj[T,D]x[D,D]",                    2 * T * D * D,     lambda: ops.gemm(dU2, Wp, dZ, tile_n=tile_n)),
        ("dgrad qkv [T,3D]x[3D,D]",                  2 * T * D * 3 * D, lambda: ops.gemm(dQKV, Wqkv, dZ, tile_n=tile_n)),
        ("wgrad fc2 [4D,T]x[T,D] f32",               2 * T * D * Hd,    lambda: ops.gemm(H, dU2, gW2, a_mn=True, b_mn=True, accum=True, tile_n=tile_n)),
        ("wgrad fc1 [D,T]x[T,4D] f32",               2 * T * D * Hd,    lambda: ops.gemm(X, H, gW1, a_mn=True, b_mn=True, accum=True, tile_n=tile_n)),
        ("wgrad qkv [D,T]x[T,3D] f32",               2 * T * D * 3 * D, lambda: ops.gemm(X, dQKV, gWqkv, a_mn=True, b_mn=True, accum=True, tile_n=tile_n)),
        ("wgrad proj[D,T]x[T,D] f32",                2 * T * D * D,     lambda: ops.gemm(X, dU2, gWp, a_mn=True, b_mn=True, accum=True, tile_n=tile_n)),
    ]
    tot_f, tot_ms = 0, 0
    for name, fl, fn in cases:
        if only and only not in name: continue
        ms = timeit(fn)
        tot_f += fl; tot_ms += ms
        print(f"  {name:44s} {ms:7.3f} ms  {fl/ms/1e9:7.1f} TFLOP/s", flush=True)
    print(f"  {'sum (student block fwd+bwd GEMMs)':44s} {tot_ms:7.3f} ms  {tot_f/tot_ms/1e9:7.1f} TFLOP/s")


if __name__ == "__main__":
    print(f"== student stream T={Ts}, tile_n auto"); run(0)
    if len(sys.argv) > 1:
        for bn in (128, 256):
            print(f"== forced tile_n={bn}"); run(bn, only="wgrad")
