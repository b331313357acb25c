"""Bring-up: per-shape timing of the fp16-operand GEMMs of one DeiT-Small block (forward, data gradients, weight gradients) with rotating buffers
(so the operands are not L2-resident from the previous iteration).  UVC_LIB_PATH selects an alternative build of the library."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from uvc_b200 import ops
g = torch.Generator(device="cuda"); g.manual_seed(0)
rn = lambda *s: torch.randn(*s, device="cuda", generator=g)
NB = 4


def t(fn, n=24):
    for i in range(4): fn(i)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


M = int(os.environ.get("M", 25216))
C, Fh = int(os.environ.get("C", 384)), int(os.environ.get("FH", 1536))
d1 = torch.tensor([0.7], device="cuda")
rows = []
def report(tag, us, flops):
    rows.append((tag, us, flops))
    print(f"{tag:34s} {us:7.1f} us   {flops / us / 1e6:7.0f} TFLOP/s", flush=True)

for (N, K, tag, kind) in [(3 * C, C, "qkv fwd (bias, f16 out)", "d16"), (C, C, "proj fwd (bias, +res, f32 out)", "res"), (Fh, C, "fc1 fwd (bias, gelu, h16 + gelu'16)", "gelu"),
                          (Fh, C, "fc1 fwd teacher (gelu, h16)", "gelu_t"), (C, Fh, "fc2 fwd (bias, +res, f32 out)", "res"),
                          (Fh, C, "fc2 dgrad (x gelu', colsum, f16 out)", "gbwd"), (C, Fh, "fc1 dgrad (f16 out)", "d16nb"), (C, C, "proj dgrad (f16 out)", "d16nb"),
                          (C, 3 * C, "qkv dgrad (f16 out)", "d16nb")]:
    A = [rn(M, K).half() for _ in range(NB)]; B = (rn(N, K) * 0.1).half(); bias = rn(N)
    D16 = [torch.empty(M, N, device="cuda", dtype=torch.float16) for _ in range(NB)]
    if kind == "d16":
        fn = lambda i: ops.gemm(A[i % NB], B, None, M, N, K, bias=bias, D16=D16[i % NB])
    elif kind == "d16nb":
        fn = lambda i: ops.gemm(A[i % NB], B, None, M, N, K, D16=D16[i % NB])
    elif kind == "res":
        R = [rn(M, N) for _ in range(NB)]; D = [torch.empty(M, N, device="cuda") for _ in range(NB)]
        fn = lambda i: ops.gemm(A[i % NB], B, D[i % NB], M, N, K, bias=bias, R=R[i % NB])
    elif kind == "gelu":
        aux = [torch.empty(M, N, device="cuda", dtype=torch.float16) for _ in range(NB)]
        fn = lambda i: ops.gemm(A[i % NB], B, None, M, N, K, bias=bias, flags=ops.EPI_GELU, D16=D16[i % NB], aux=aux[i % NB])
    elif kind == "gelu_t":
        fn = lambda i: ops.gemm(A[i % NB], B, None, M, N, K, bias=bias, flags=ops.EPI_GELU, D16=D16[i % NB])
    elif kind == "gbwd":
        aux = [torch.rand(M, N, device="cuda").half() for _ in range(NB)]; cs = torch.zeros(N, device="cuda")
        fn = lambda i: ops.gemm(A[i % NB], B, None, M, N, K, aux=aux[i % NB], flags=ops.EPI_GELU_BWD, D16=D16[i % NB], colsum=cs, colsum_scale=0.5, alpha_dev=d1)
    report(tag, t(fn), 2.0 * M * N * K)
    del A, D16

for (Nw, Kw, tag) in [(C, Fh, "fc2 wgrad"), (Fh, C, "fc1 wgrad"), (C, C, "proj wgrad"), (3 * C, C, "qkv wgrad")]:
    dY = [rn(M, Nw).half() for _ in range(NB)]; X = [rn(M, Kw).half() for _ in range(NB)]
    dW = torch.zeros(Nw, Kw, device="cuda")
    tiles = ((Nw + 127) // 128) * ((Kw + 127) // 128)
    s = max(1, min((2 * 148) // tiles, ((M + 63) // 64) // 4))
    fn = lambda i: ops.gemm(ops.operand(dY[i % NB], mn_major=True), ops.operand(X[i % NB], mn_major=True), dW, Nw, Kw, M, splits=s, flags=ops.GEMM_F16 | ops.EPI_ATOMIC, alpha=0.5)
    report(f"{tag} (split-K {s})", t(fn), 2.0 * M * Nw * Kw)
tot = sum(u for _, u, _ in rows); fl = sum(f for _, _, f in rows)
print(f"sum {tot:.1f} us, {fl / tot / 1e6:.0f} TFLOP/s overall")
