"""Bring-up: LayerNorm forward / backward (fp16-operand variants) timing at the DeiT-Small bench shape with rotating buffers."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from uvc_b200 import ops
g = torch.Generator(device="cuda"); g.manual_seed(0)
rn = lambda *s: torch.randn(*s, device="cuda", generator=g)
M, C = int(os.environ.get("M", 25216)), int(os.environ.get("C", 384))
NB = 4
def t(fn, n=24):
    for i in range(4): fn(i)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
xs = [rn(M, C) for _ in range(NB)]; gam, bet = rn(C), rn(C)
y, mean, rstd = ops.layernorm_fwd_f16(xs[0], gam, bet, 1e-6)
dy = [rn(M, C).half() for _ in range(NB)]; gs = [rn(M, C) for _ in range(NB)]; r1s = [rn(M, C) for _ in range(NB)]; ts = [rn(M, C) for _ in range(NB)]
one = torch.ones(1, device="cuda")
dg, db, c1, c2 = (torch.zeros(C, device="cuda") for _ in range(4))
dots = torch.zeros(2, device="cuda")
us = t(lambda i: ops.layernorm_fwd_f16(xs[i % NB], gam, bet, 1e-6))
print(f"LN fwd (f32 -> f16)                    {us:6.1f} us  {(M*C*6)/us/1e3:6.0f} GB/s")
us = t(lambda i: ops.layernorm_bwd_f16(dy[i % NB], 1.0, xs[i % NB], mean, rstd, gam, r2=gs[i % NB], s2=one, dgamma=dg, dbeta=db, cs_r1=c1, cs_out=c2))
print(f"LN2 bwd (dy16, x, g -> dx, dx16, col sums) {us:6.1f} us  {(M*C*(2+4+4+4+2))/us/1e3:6.0f} GB/s")
us = t(lambda i: ops.layernorm_bwd_f16(dy[i % NB], 1.0, xs[i % NB], mean, rstd, gam, r1=r1s[i % NB], r2=gs[i % NB], s2=one, dgamma=dg, dbeta=db))
print(f"LN1 bwd (dy16, x, dx1, g -> dx, dx16)      {us:6.1f} us  {(M*C*(2+4+4+4+4+2))/us/1e3:6.0f} GB/s")
