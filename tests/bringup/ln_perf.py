"""Bring-up: LayerNorm backward at the bench shape (25216 x 384), plain and with the fused column sums."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from uvc_b200 import ops
M, C = 25216, 384
dev = "cuda"
dy, x, r1, r2 = (torch.randn(M, C, device=dev) for _ in range(4))
mean, rstd = torch.randn(M, device=dev), torch.rand(M, device=dev) + 0.5
gamma = torch.randn(C, device=dev); s2 = torch.tensor([0.7], device=dev)
dg, db, c1, c2 = (torch.zeros(C, device=dev) for _ in range(4))
dx = torch.empty(M, C, device=dev)
def timeit(fn, n=30):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
tag = os.environ.get("UVC_LNB_BLOCKS", "6")
print(tag, "ln1-style (r1,r2):   %.1f us" % timeit(lambda: ops.layernorm_bwd(dy, x, mean, rstd, gamma, r1=r1, r2=r2, s2=s2, dgamma=dg, dbeta=db, dx=dx)))
print(tag, "ln2-style (r2,cs):   %.1f us" % timeit(lambda: ops.layernorm_bwd(dy, x, mean, rstd, gamma, r2=r2, s2=s2, dgamma=dg, dbeta=db, dx=dx, cs_r1=c1, cs_out=c2)))
print(tag, "no param grads:      %.1f us" % timeit(lambda: ops.layernorm_bwd(dy, x, mean, rstd, gamma, r1=r1, r2=r2, s2=s2, dx=dx)))
