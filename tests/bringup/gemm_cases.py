"""GPU bring-up for the tcgen05 TF32 GEMM: each case runs in its own process (a trap poisons the context).
usage: python tests/bringup/gemm_cases.py <case>|all
"""
import os, subprocess, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

CASES = ["nt_small", "nt_big", "nn_small", "tn_small", "tt_small", "dgrad_big", "wgrad_big", "batched_qk", "batched_pv", "epilogues", "edges", "perf"]


def rel_err(a, b):
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def report(name, got, ref, tol=3e-3):
    e = rel_err(got, ref)
    bad = not (e < tol)
    print(f"  {name}: rel_err={e:.3e} {'FAIL' if bad else 'ok'}", flush=True)
    if bad:
        d = (got - ref).abs()
        flat = d.flatten()
        idx = flat.argmax().item()
        print(f"    shape={tuple(got.shape)} worst idx={idx} got={got.flatten()[idx].item()} ref={ref.flatten()[idx].item()} nan={torch.isnan(got).sum().item()}")
        if got.dim() == 2:
            rows_bad = (d > tol * ref.abs().max()).any(1).nonzero().flatten()[:16].tolist()
            cols_bad = (d > tol * ref.abs().max()).any(0).nonzero().flatten()[:16].tolist()
            print(f"    bad rows (first 16): {rows_bad}\n    bad cols (first 16): {cols_bad}")
            print("    got[0,:8]", got[0, :8].tolist()); print("    ref[0,:8]", ref[0, :8].tolist())
    return not bad


def run_case(case):
    from uvc_b200 import ops
    torch.backends.cuda.matmul.allow_tf32 = False
    dev = "cuda"
    g = torch.Generator(device=dev); g.manual_seed(0)
    rn = lambda *s: torch.randn(*s, device=dev, generator=g)
    ok = True
    if case in ("nt_small", "nt_big"):
        M, N, K = (256, 128, 64) if case == "nt_small" else (25216, 1152, 384)
        A, B = rn(M, K), rn(N, K); D = torch.empty(M, N, device=dev)
        ops.gemm(A, B, D, M, N, K); torch.cuda.synchronize()
        ok &= report(case, D, A @ B.t())
    elif case == "nn_small":      # A K-major, B MN-major: D = A @ Bm, Bm stored [K,N]
        M, N, K = 256, 256, 96
        A, Bm = rn(M, K), rn(K, N); D = torch.empty(M, N, device=dev)
        ops.gemm(A, ops.operand(Bm, mn_major=True), D, M, N, K); torch.cuda.synchronize()
        ok &= report(case, D, A @ Bm)
    elif case == "tn_small":      # A MN-major (stored [K,M]), B K-major
        M, N, K = 256, 128, 96
        Am, B = rn(K, M), rn(N, K); D = torch.empty(M, N, device=dev)
        ops.gemm(ops.operand(Am, mn_major=True), B, D, M, N, K); torch.cuda.synchronize()
        ok &= report(case, D, Am.t() @ B.t())
    elif case == "tt_small":      # both MN-major: D = Am^T @ Bm
        M, N, K = 256, 256, 160
        Am, Bm = rn(K, M), rn(K, N); D = torch.empty(M, N, device=dev)
        ops.gemm(ops.operand(Am, mn_major=True), ops.operand(Bm, mn_major=True), D, M, N, K); torch.cuda.synchronize()
        ok &= report(case, D, Am.t() @ Bm)
    elif case == "dgrad_big":     # dX[M,K'] = dY[M,N'] @ W[N',K']
        M, Np, Kp = 25216, 1536, 384
        dY, W = rn(M, Np), rn(Np, Kp); dX = torch.empty(M, Kp, device=dev)
        ops.gemm(dY, ops.operand(W, mn_major=True), dX, M, Kp, Np); torch.cuda.synchronize()
        ok &= report(case, dX, dY @ W)
    elif case == "wgrad_big":     # dW[N',K'] = dY^T @ X  (split-K, atomic)
        M, Np, Kp = 25216, 1536, 384
        dY, X = rn(M, Np), rn(M, Kp); dW = torch.zeros(Np, Kp, device=dev)
        ops.gemm(ops.operand(dY, mn_major=True), ops.operand(X, mn_major=True), dW, Np, Kp, M, splits=16, flags=ops.EPI_ATOMIC); torch.cuda.synchronize()
        ok &= report(case, dW, dY.t() @ X)
        dW2 = torch.zeros(Np, Kp, device=dev)
        ops.gemm(ops.operand(dY, mn_major=True), ops.operand(X, mn_major=True), dW2, Np, Kp, M); torch.cuda.synchronize()
        ok &= report(case + "_nosplit", dW2, dY.t() @ X)
    elif case == "batched_qk":    # S[b,h] = scale * Q K^T out of a [B,197,3,H,64] qkv buffer
        Bz, H, Nt, d = 4, 6, 197, 64; Cc = H * d
        qkv = rn(Bz, Nt, 3, H, d); ldS = 208
        S = torch.zeros(Bz, H, Nt, ldS, device=dev)
        q = qkv[:, :, 0]; k = qkv[:, :, 1]
        Aop = ops.Operand(q.data_ptr(), 3 * Cc, d, Nt * 3 * Cc, 0, 0)
        Bop = ops.Operand(k.data_ptr(), 3 * Cc, d, Nt * 3 * Cc, 0, 0)
        ops.gemm(Aop, Bop, S, Nt, Nt, d, ldd=ldS, d_bs=(Nt * ldS, H * Nt * ldS), batch=(H, Bz), alpha=0.125); torch.cuda.synchronize()
        ref = torch.einsum("bnhd,bmhd->bhnm", q, k) * 0.125
        ok &= report(case, S[..., :Nt].contiguous(), ref)
        print("    pad cols untouched:", bool((S[..., Nt:] == 0).all().item()))
    elif case == "batched_pv":    # ctx[b,n,h,:] = P[b,h] @ V[b,h]
        Bz, H, Nt, d = 4, 6, 197, 64; Cc = H * d; ldS = 208
        qkv = rn(Bz, Nt, 3, H, d); v = qkv[:, :, 2]
        P = torch.zeros(Bz, H, Nt, ldS, device=dev); P[..., :Nt] = torch.softmax(rn(Bz, H, Nt, Nt), -1)
        P[..., Nt:] = float("nan")   # padding must never be read (TMA clips K at 197)
        ctx = torch.empty(Bz, Nt, H, d, device=dev)
        Aop = ops.Operand(P.data_ptr(), ldS, Nt * ldS, H * Nt * ldS, 0, 0)
        Bop = ops.Operand(v.data_ptr(), 3 * Cc, d, Nt * 3 * Cc, 1, 0)
        ops.gemm(Aop, Bop, ctx, Nt, d, Nt, ldd=Cc, d_bs=(d, Nt * Cc), batch=(H, Bz)); torch.cuda.synchronize()
        ref = torch.einsum("bhnm,bmhd->bnhd", P[..., :Nt], v)
        ok &= report(case, ctx, ref)
    elif case == "epilogues":
        M, N, K = 1000, 768, 192
        A, B, bias, R = rn(M, K), rn(N, K) * 0.1, rn(N), rn(M, N)
        pre = A @ B.t() + bias
        D = torch.empty(M, N, device=dev); aux = torch.empty(M, N, device=dev)
        ops.gemm(A, B, D, M, N, K, bias=bias, aux=aux, flags=ops.EPI_GELU); torch.cuda.synchronize()
        pp = pre.clone().requires_grad_(True); torch.nn.functional.gelu(pp).sum().backward()
        ok &= report("gelu", D, torch.nn.functional.gelu(pre)); ok &= report("gelu_aux", aux, pp.grad)
        ops.gemm(A, B, D, M, N, K, bias=bias, R=R, beta=0.5); torch.cuda.synchronize()
        ok &= report("bias_residual", D, pre + 0.5 * R)
        al = torch.tensor([0.25], device=dev); be = torch.tensor([2.0], device=dev)
        ops.gemm(A, B, D, M, N, K, R=R, alpha=2.0, alpha_dev=al, beta_dev=be); torch.cuda.synchronize()
        ok &= report("alpha_beta_dev", D, 0.5 * (A @ B.t()) + 2.0 * R)
        u = rn(M, N)
        ops.gemm(A, B, D, M, N, K, aux=u, flags=ops.EPI_GELU_BWD); torch.cuda.synchronize()
        ok &= report("gelu_bwd", D, (A @ B.t()) * u)
        D0 = rn(M, N); D1 = D0.clone()
        ops.gemm(A, B, D1, M, N, K, R=D1); torch.cuda.synchronize()
        ok &= report("accumulate_inplace", D1, D0 + A @ B.t())
    elif case == "edges":
        for (M, N, K) in [(8, 1000, 192), (197, 197, 64), (130, 36, 200), (128, 128, 8), (1, 4, 4), (300, 260, 33 * 4)]:
            A, B = rn(M, K), rn(N, K); ldd = (N + 3) // 4 * 4
            D = torch.zeros(M, ldd, device=dev)
            ops.gemm(A, B, D, M, N, K, ldd=ldd); torch.cuda.synchronize()
            ok &= report(f"edge {M}x{N}x{K}", D[:, :N].contiguous(), A @ B.t())
    elif case == "perf":
        for (M, N, K, amn, bmn, splits, tag) in [(25216, 1152, 384, 0, 0, 1, "qkv fwd"), (25216, 1536, 384, 0, 0, 1, "fc1 fwd"),
                                                (25216, 384, 1536, 0, 0, 1, "fc2 fwd"), (25216, 384, 1536, 0, 1, 1, "fc1 dgrad"),
                                                (1536, 384, 25216, 1, 1, 16, "fc1 wgrad"), (8192, 8192, 8192, 0, 0, 1, "square 8k")]:
            A = rn(K, M) if amn else rn(M, K); B = rn(K, N) if bmn else rn(N, K); D = torch.zeros(M, N, device=dev)
            Ao, Bo = ops.operand(A, mn_major=bool(amn)), ops.operand(B, mn_major=bool(bmn))
            fl = ops.EPI_ATOMIC if splits > 1 else 0
            for _ in range(3): ops.gemm(Ao, Bo, D, M, N, K, splits=splits, flags=fl)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10): ops.gemm(Ao, Bo, D, M, N, K, splits=splits, flags=fl)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            print(f"  perf {tag}: {M}x{N}x{K} {ms*1e3:.1f} us  {2*M*N*K/ms/1e9:.1f} TFLOP/s", flush=True)
        torch.backends.cuda.matmul.allow_tf32 = True
        A, B = rn(8192, 8192), rn(8192, 8192)
        for _ in range(3): A @ B
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): A @ B
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"  cuBLAS tf32 8192^3: {ms*1e3:.1f} us {2*8192**3/ms/1e9:.1f} TFLOP/s")
    elif case == "v2":
        # CTA-pair kernel (run with UVC_GEMM_V2=2 and UVC_GEMM_V2_BN in {128,192,256}): majors, ragged edges, split-K, epilogues
        for (M, N, K, amn, bmn, splits) in [(512, 256, 64, 0, 0, 1), (1000, 388, 100, 0, 0, 1), (300, 260, 132, 0, 1, 1), (776, 512, 96, 1, 0, 1),
                                            (1536, 384, 2500, 1, 1, 5), (25216, 1152, 384, 0, 0, 1), (25216, 384, 1536, 0, 1, 1), (130, 36, 200, 0, 0, 1)]:
            A = rn(K, M) if amn else rn(M, K); B = rn(K, N) if bmn else rn(N, K)
            D = torch.zeros(M, N, device=dev)
            fl = ops.EPI_ATOMIC if splits > 1 else 0
            ops.gemm(ops.operand(A, mn_major=bool(amn)), ops.operand(B, mn_major=bool(bmn)), D, M, N, K, splits=splits, flags=fl); torch.cuda.synchronize()
            ref = (A.t() if amn else A) @ (B if bmn else B.t())
            ok &= report(f"v2 {M}x{N}x{K} a_mn={amn} b_mn={bmn} splits={splits}", D, ref)
        M, N, K = 1000, 768, 192
        A, B, bias, R = rn(M, K), rn(N, K) * 0.1, rn(N), rn(M, N)
        pre = A @ B.t() + bias
        D = torch.empty(M, N, device=dev); aux = torch.empty(M, N, device=dev)
        ops.gemm(A, B, D, M, N, K, bias=bias, aux=aux, flags=ops.EPI_GELU); torch.cuda.synchronize()
        pp = pre.clone().requires_grad_(True); torch.nn.functional.gelu(pp).sum().backward()
        ok &= report("v2 gelu", D, torch.nn.functional.gelu(pre)); ok &= report("v2 gelu_aux", aux, pp.grad)
        ops.gemm(A, B, D, M, N, K, bias=bias, R=R, beta=0.5); torch.cuda.synchronize()
        ok &= report("v2 bias_residual", D, pre + 0.5 * R)
        u = rn(M, N)
        ops.gemm(A, B, D, M, N, K, aux=u, flags=ops.EPI_GELU_BWD | ops.EPI_ROUND_TF32); torch.cuda.synchronize()
        ok &= report("v2 gelu_bwd", D, (A @ B.t()) * u)
        D0 = rn(M, N); D1 = D0.clone()
        ops.gemm(A, B, D1, M, N, K, R=D1); torch.cuda.synchronize()
        ok &= report("v2 accumulate_inplace", D1, D0 + A @ B.t())
        # many back-to-back launches (pipeline phase bookkeeping across units, TMEM alloc/free across kernels)
        A, B = rn(4096, 384), rn(1536, 384); D = torch.empty(4096, 1536, device=dev)
        for _ in range(20): ops.gemm(A, B, D, 4096, 1536, 384)
        torch.cuda.synchronize()
        ok &= report("v2 repeat", D, A @ B.t())
    elif case == "perf2":
        shapes = [(25216, 1152, 384, 0, 0, 1, 0, "qkv fwd"), (25216, 1536, 384, 0, 0, 1, ops.EPI_GELU, "fc1 fwd+gelu"), (25216, 384, 1536, 0, 0, 1, 0, "fc2 fwd"),
                  (25216, 384, 384, 0, 0, 1, 0, "proj fwd"), (25216, 384, 1536, 0, 1, 1, 0, "fc1 dgrad"), (25216, 1536, 384, 0, 1, 1, 0, "fc2 dgrad"),
                  (25216, 384, 1152, 0, 1, 1, 0, "qkv dgrad"), (1536, 384, 25216, 1, 1, 0, ops.EPI_ATOMIC, "fc1 wgrad"), (384, 1536, 25216, 1, 1, 0, ops.EPI_ATOMIC, "fc2 wgrad"),
                  (1152, 384, 25216, 1, 1, 0, ops.EPI_ATOMIC, "qkv wgrad"), (384, 384, 25216, 1, 1, 0, ops.EPI_ATOMIC, "proj wgrad"), (8192, 8192, 8192, 0, 0, 1, 0, "square 8k")]
        for (M, N, K, amn, bmn, splits, fl, tag) in shapes:
            A = rn(K, M) if amn else rn(M, K); B = rn(K, N) if bmn else rn(N, K); D = torch.zeros(M, N, device=dev)
            aux = torch.empty(M, N, device=dev) if fl & ops.EPI_GELU else None
            bias = rn(N) if fl & ops.EPI_GELU else None
            Ao, Bo = ops.operand(A, mn_major=bool(amn)), ops.operand(B, mn_major=bool(bmn))
            for sp in ([1] if splits else [4, 8, 16]):
                for _ in range(3): ops.gemm(Ao, Bo, D, M, N, K, splits=sp, flags=fl, aux=aux, bias=bias)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(10): ops.gemm(Ao, Bo, D, M, N, K, splits=sp, flags=fl, aux=aux, bias=bias)
                e1.record(); torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / 10
                print(f"  perf2[{os.environ.get('UVC_GEMM_V2','1')}/{os.environ.get('UVC_GEMM_V2_BN','auto')}] {tag}: {M}x{N}x{K} splits={sp} {ms*1e3:.1f} us  {2*M*N*K/ms/1e9:.1f} TFLOP/s", flush=True)
    print(f"CASE {case}: {'PASS' if ok else 'FAIL'}", flush=True)
    return ok


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which != "all":
        sys.exit(0 if run_case(which) else 1)
    summary = {}
    for c in CASES:
        print(f"=== {c}", flush=True)
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), c], timeout=180)
            summary[c] = "PASS" if r.returncode == 0 else f"FAIL(rc={r.returncode})"
        except subprocess.TimeoutExpired:
            summary[c] = "TIMEOUT"
    print("SUMMARY", summary)
