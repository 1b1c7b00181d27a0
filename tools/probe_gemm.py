"""GPU probe: d3_gemm_bf16 vs torch fp32 matmul on bf16 operands, all operand-major combinations.
Run one combination per process (argv: a_mn b_mn) so a hang or fault in one does not hide the others."""
import os, sys, subprocess, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "dinov3-jax_b200"))


def run_combo(a_mn, b_mn):
    import torch
    from dinov3_jax import ops, _native as N
    torch.manual_seed(0)
    dev = "cuda"
    shapes = [(128, 64, 64), (128, 256, 64), (256, 256, 128), (128, 128, 512), (384, 320, 192), (296, 200, 136),
              (1000, 1152, 384), (4096, 1024, 1024), (25216, 3072, 1024)]
    for (M, Nn, K) in shapes:
        for bn in (0, 64, 128, 256, 512):
            if bn == 0 and M > 5000:
                pass
            A = torch.randn(M, K, device=dev).to(torch.bfloat16)
            B = torch.randn(K, Nn, device=dev).to(torch.bfloat16)   # logical [K,N]
            ref = A.float() @ B.float()
            A_st = A.t().contiguous() if a_mn else A
            B_st = B if b_mn else B.t().contiguous()
            out = torch.full((M, Nn), float("nan"), device=dev, dtype=torch.float32)
            try:
                ops.gemm(A_st, B_st, out, a_mn=a_mn, b_mn=b_mn, tile_n=bn)
                torch.cuda.synchronize()
            except Exception as e:  # noqa
                print(f"  a_mn={a_mn} b_mn={b_mn} M={M} N={Nn} K={K} bn={bn}: EXC {e}", flush=True)
                return 1
            err = (out - ref).abs().max().item()
            scale = ref.abs().max().item()
            nan = torch.isnan(out).sum().item()
            status = "ok" if (err <= 2e-3 * scale and nan == 0) else "BAD"
            line = f"  a_mn={a_mn} b_mn={b_mn} M={M} N={Nn} K={K} bn={bn}: maxerr={err:.4g} scale={scale:.3g} nan={nan} {status}"
            if status == "BAD":
                bad = ((out - ref).abs() > 2e-3 * scale) | torch.isnan(out)
                idx = bad.nonzero()
                rows = idx[:, 0].unique()[:8].tolist(); cols = idx[:, 1].unique()[:8].tolist()
                line += f" nbad={bad.sum().item()} rows={rows} cols={cols} out00={out[0,0].item():.4g} ref00={ref[0,0].item():.4g}"
            print(line, flush=True)
    # epilogue check (forward layout: A K-major, B as given)
    M, Nn, K = 504, 384, 256
    A = torch.randn(M, K, device=dev).to(torch.bfloat16); B = (torch.randn(K, Nn, device=dev) * 0.1).to(torch.bfloat16)
    A_st = A.t().contiguous() if a_mn else A
    B_st = B if b_mn else B.t().contiguous()
    bias = torch.randn(Nn, device=dev); gamma = torch.randn(Nn, device=dev); resid = torch.randn(M, Nn, device=dev)
    acc = A.float() @ B.float()
    u = acc + bias
    gel = torch.nn.functional.gelu(u, approximate="tanh")
    want = resid + gamma * gel
    out = torch.empty(M, Nn, device=dev); pre = torch.empty(M, Nn, device=dev, dtype=torch.bfloat16)
    ops.gemm(A_st, B_st, out, a_mn=a_mn, b_mn=b_mn, bias=bias, gelu=True, store_pre=pre, gamma=gamma, resid=resid)
    torch.cuda.synchronize()
    print(f"  epilogue bias+gelu+gamma+resid: maxerr={(out-want).abs().max().item():.4g}  pre err={(pre.float()-u).abs().max().item():.4g}")
    # dgelu multiply + bf16 out
    ub = torch.randn(M, Nn, device=dev).to(torch.bfloat16)
    uf = ub.float().requires_grad_(True)
    torch.nn.functional.gelu(uf, approximate="tanh").sum().backward()
    want2 = acc * uf.grad
    out2 = torch.empty(M, Nn, device=dev, dtype=torch.bfloat16)
    ops.gemm(A_st, B_st, out2, a_mn=a_mn, b_mn=b_mn, dgelu_of=ub)
    torch.cuda.synchronize()
    print(f"  epilogue dgelu bf16-out: maxerr={(out2.float()-want2).abs().max().item():.4g} scale={want2.abs().max().item():.3g}")
    # accumulate
    out3 = torch.ones(M, Nn, device=dev)
    ops.gemm(A_st, B_st, out3, a_mn=a_mn, b_mn=b_mn, accum=True, alpha=0.5)
    torch.cuda.synchronize()
    print(f"  epilogue accum alpha: maxerr={(out3-(1+0.5*acc)).abs().max().item():.4g}")
    # CTA-pair kernel with the TMA epilogue (tile_n=512), ragged M, every epilogue combination the engine uses
    for (M, Nn, K) in [(1000, 512, 256), (777, 1024, 320), (25216, 1024, 1024)]:
        A = torch.randn(M, K, device=dev).to(torch.bfloat16); B = (torch.randn(K, Nn, device=dev) * 0.1).to(torch.bfloat16)
        A_st = A.t().contiguous() if a_mn else A
        B_st = B if b_mn else B.t().contiguous()
        bias = torch.randn(Nn, device=dev); gamma = torch.randn(Nn, device=dev); resid = torch.randn(M, Nn, device=dev)
        acc = A.float() @ B.float(); u = acc + bias
        gel = torch.nn.functional.gelu(u, approximate="tanh")
        def chk(name, got, want, tol):
            e = ((got.float() - want).norm() / want.norm()).item()
            print(f"  tma-epi M={M} N={Nn} K={K} {name}: rel={e:.3e} nan={int(torch.isnan(got.float()).sum())} {'ok' if e < tol else 'BAD'}", flush=True)
        o = torch.full((M, Nn), float('nan'), device=dev, dtype=torch.bfloat16)
        ops.gemm(A_st, B_st, o, a_mn=a_mn, b_mn=b_mn, bias=bias, tile_n=512); chk("bias->bf16", o, u, 4e-3)
        o = torch.full((M, Nn), float('nan'), device=dev, dtype=torch.bfloat16); pre = torch.full((M, Nn), float('nan'), device=dev, dtype=torch.bfloat16)
        ops.gemm(A_st, B_st, o, a_mn=a_mn, b_mn=b_mn, bias=bias, gelu=True, store_pre=pre, tile_n=512); chk("bias,gelu,pre out", o, gel, 4e-3); chk("bias,gelu,pre pre", pre, u, 4e-3)
        o = torch.full((M, Nn), float('nan'), device=dev); pre = torch.full((M, Nn), float('nan'), device=dev, dtype=torch.bfloat16)
        ops.gemm(A_st, B_st, o, a_mn=a_mn, b_mn=b_mn, bias=bias, store_pre=pre, gamma=gamma, resid=resid, tile_n=512); chk("proj-like f32", o, resid + gamma * u, 1e-5); chk("proj-like pre", pre, u, 4e-3)
        o = torch.full((M, Nn), float('nan'), device=dev)
        ops.gemm(A_st, B_st, o, a_mn=a_mn, b_mn=b_mn, bias=bias, gelu=True, store_pre=pre, gamma=gamma, resid=resid, tile_n=512); chk("fc2-like f32", o, resid + gamma * gel, 5e-4)
        ub = torch.randn(M, Nn, device=dev).to(torch.bfloat16); uf = ub.float().requires_grad_(True)
        torch.nn.functional.gelu(uf, approximate="tanh").sum().backward()
        o = torch.full((M, Nn), float('nan'), device=dev, dtype=torch.bfloat16)
        ops.gemm(A_st, B_st, o, a_mn=a_mn, b_mn=b_mn, dgelu_of=ub, tile_n=512); chk("dgelu->bf16", o, acc * uf.grad, 4e-3)
        o = torch.full((M, Nn), float('nan'), device=dev)
        ops.gemm(A_st, B_st, o, a_mn=a_mn, b_mn=b_mn, tile_n=512); chk("plain f32", o, acc, 1e-5)
    # split-K (fp32 atomic reduction into a zeroed output), both kernels
    for (M, Nn, K, bn, sk) in [(256, 256, 4096, 512, 8), (1024, 1024, 8192, 512, 4), (384, 320, 2048, 128, 5), (1024, 1024, 44160, 0, 0)]:
        A = torch.randn(M, K, device=dev).to(torch.bfloat16); B = torch.randn(K, Nn, device=dev).to(torch.bfloat16)
        A_st = A.t().contiguous() if a_mn else A
        B_st = B if b_mn else B.t().contiguous()
        ref = A.float() @ B.float()
        out = torch.zeros(M, Nn, device=dev)
        ops.gemm(A_st, B_st, out, a_mn=a_mn, b_mn=b_mn, accum=True, tile_n=bn, split_k=sk)
        torch.cuda.synchronize()
        err = (out - ref).abs().max().item(); scale = ref.abs().max().item()
        print(f"  split-K M={M} N={Nn} K={K} bn={bn} split={sk}: maxerr={err:.4g} scale={scale:.3g} {'ok' if err <= 2e-3*scale else 'BAD'}", flush=True)
    # timing of the big shape
    M, Nn, K = 25216, 4096, 1024
    A = torch.randn(M, K, device=dev).to(torch.bfloat16); B = torch.randn(K, Nn, device=dev).to(torch.bfloat16)
    A_st = A.t().contiguous() if a_mn else A
    B_st = B if b_mn else B.t().contiguous()
    out = torch.empty(M, Nn, device=dev, dtype=torch.bfloat16)
    for bn in (128, 256, 512):
        for _ in range(3):
            ops.gemm(A_st, B_st, out, a_mn=a_mn, b_mn=b_mn, tile_n=bn)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(10):
            ops.gemm(A_st, B_st, out, a_mn=a_mn, b_mn=b_mn, tile_n=bn)
        e.record(); torch.cuda.synchronize()
        ms = s.elapsed_time(e) / 10
        print(f"  perf M={M} N={Nn} K={K} bn={bn}: {ms:.3f} ms  {2*M*Nn*K/ms/1e9:.1f} TFLOP/s", flush=True)
    return 0


if __name__ == "__main__":
    if len(sys.argv) == 3:
        sys.exit(run_combo(bool(int(sys.argv[1])), bool(int(sys.argv[2]))))
    for a_mn in (0, 1):
        for b_mn in (0, 1):
            print(f"== combo a_mn={a_mn} b_mn={b_mn}", flush=True)
            t0 = time.time()
            try:
                r = subprocess.run([sys.executable, __file__, str(a_mn), str(b_mn)], timeout=240)
                print(f"== exit {r.returncode} in {time.time()-t0:.1f}s", flush=True)
            except subprocess.TimeoutExpired:
                print("== TIMEOUT (hang)", flush=True)
