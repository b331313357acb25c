import torch, sys
sys.path.insert(0, "/root/repo")
from uvc_b200 import ops
M, N, K = 25216, 1536, 384
g = torch.Generator(device="cuda"); g.manual_seed(0)
dY = torch.randn(M, K, device="cuda", generator=g); W = torch.randn(K, N, device="cuda", generator=g); aux = torch.randn(M, N, device="cuda", generator=g)
D = torch.empty(M, N, device="cuda"); cs = torch.zeros(N, device="cuda")
def t(fn, n=20):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
Wm = ops.operand(W, mn_major=True)
print("plain dgrad        %.1f us" % t(lambda: ops.gemm(dY, Wm, D, M, N, K)))
print("gelu' + round      %.1f us" % t(lambda: ops.gemm(dY, Wm, D, M, N, K, aux=aux, flags=ops.EPI_GELU_BWD | ops.EPI_ROUND_TF32)))
print("gelu' + round + cs %.1f us" % t(lambda: ops.gemm(dY, Wm, D, M, N, K, aux=aux, flags=ops.EPI_GELU_BWD | ops.EPI_ROUND_TF32, colsum=cs)))
