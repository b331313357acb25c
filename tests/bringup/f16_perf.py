import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from uvc_b200 import ops
g = torch.Generator(device="cuda"); g.manual_seed(0)
rn = lambda *s: torch.randn(*s, device="cuda", generator=g)
def t(fn, n=20):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
M = 25216
for (N, K, tag, kw) in [(1152, 384, "qkv", {}), (384, 384, "proj(+res)", {"res": 1}), (1536, 384, "fc1+gelu->h16", {"gelu": 1}), (384, 1536, "fc2(+res)", {"res": 1})]:
    A32, B32 = rn(M, K), rn(N, K) * 0.1
    A16, B16 = A32.half(), B32.half()
    bias = rn(N); R = rn(M, N) if kw.get("res") else None
    D = torch.empty(M, N, device="cuda"); D16 = torch.empty(M, N, device="cuda", dtype=torch.float16)
    if kw.get("gelu"):
        t32 = t(lambda: ops.gemm(A32, B32, D, M, N, K, bias=bias, flags=ops.EPI_GELU | ops.EPI_ROUND_TF32))
        t16 = t(lambda: ops.gemm(A16, B16, None, M, N, K, bias=bias, flags=ops.EPI_GELU, D16=D16))
        t16b = t(lambda: ops.gemm(A16, B16, D, M, N, K, bias=bias, flags=ops.EPI_GELU | ops.EPI_ROUND_TF32, D16=D16, aux=D.clone()))
        print(f"{tag:16s} tf32 {t32:6.1f} us | fp16 (h16 only) {t16:6.1f} us | fp16 (h16 + h32 + gelu') {t16b:6.1f} us")
    else:
        t32 = t(lambda: ops.gemm(A32, B32, D, M, N, K, bias=bias, R=R))
        t16 = t(lambda: ops.gemm(A16, B16, D, M, N, K, bias=bias, R=R))
        print(f"{tag:16s} tf32 {t32:6.1f} us | fp16 operands {t16:6.1f} us   ({2*M*N*K/t16/1e6:.0f} TFLOP/s)")
