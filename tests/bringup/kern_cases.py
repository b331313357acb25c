"""Bring-up: run the hot kernels of one DeiT-Small block once each inside a cudaProfilerStart/Stop window (for `ncu --profile-from-start off`)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from uvc_b200 import ops
g = torch.Generator(device="cuda"); g.manual_seed(0)
rn = lambda *s: torch.randn(*s, device="cuda", generator=g)
M, C, Fh, B, H, N = 25216, 384, 1536, 128, 6, 197
which = set((os.environ.get("CASES") or "fc1,proj,qkv,fc2dg,attn,ln,wgrad,blend,optim,loss,t2t").split(","))
ln1 = rn(M, C).half(); W1 = (rn(Fh, C) * 0.1).half(); b1 = rn(Fh)
h16 = torch.empty(M, Fh, device="cuda", dtype=torch.float16); aux = torch.empty_like(h16)
ctx = rn(M, C).half(); Wp = (rn(C, C) * 0.1).half(); bp = rn(C); x = rn(M, C); x1 = torch.empty(M, C, device="cuda")
Wq = (rn(3 * C, C) * 0.1).half(); bq = rn(3 * C); qkv = torch.empty(M, 3 * C, device="cuda", dtype=torch.float16)
g16 = rn(M, C).half(); W2T = (rn(Fh, C) * 0.1).half(); cs = torch.zeros(Fh, device="cuda"); dh = torch.empty_like(h16)
qkv_in = rn(M, 3 * C).half()
xs = rn(M, C); gam, bet = rn(C), rn(C)
hh = rn(M, Fh).half(); W2 = (rn(C, Fh) * 0.1).half(); b2 = rn(C); t_keep = torch.empty(M, C, device="cuda"); xo = torch.empty(M, C, device="cuda")
blend = torch.tensor([0.3, 0.7], device="cuda")
P = 22_050_664                       # DeiT-Small parameter arena
if "optim" in which:
    pw, pg, pm, pv = (rn(P) * 0.02 for _ in range(4)); pv = pv.abs()
    pflags = torch.full((P,), 7, dtype=torch.uint8, device="cuda"); acc = torch.zeros(1, device="cuda")
if "loss" in which:
    lg, tl = rn(B, 1000), rn(B, 1000); tg = torch.softmax(rn(B, 1000), -1)
if "t2t" in which:
    from uvc_b200.T2TViT.models import T2T_module
    t2t = T2T_module(embed_dim=384).cuda().eval()
    xt = rn(32, 3, 224, 224); rt = rn(32, 196, 384) * 0.01
def run():
    if "blend" in which: ops.gemm(hh, W2, xo, M, C, Fh, bias=b2, R=x, blend=blend, R2=x1, D2=t_keep)
    if "optim" in which:
        acc.zero_(); ops.sqnorm_accum_flags_(pg, pflags, acc); ops.clip_adamw_flags_(pw, pg, pm, pv, pflags, acc, 1.0, 1e-4, 0.9, 0.999, 1e-8, 0.05, 3)
    if "loss" in which: ops.distill_loss(lg, tl, tg, 0.1, 1.0)
    if "t2t" in which:
        tok, _ = t2t(xt); (tok * rt).sum().backward()
    if "fc1" in which: ops.gemm(ln1, W1, None, M, Fh, C, bias=b1, flags=ops.EPI_GELU, D16=h16, aux=aux)
    if "proj" in which: ops.gemm(ctx, Wp, x1, M, C, C, bias=bp, R=x)
    if "qkv" in which: ops.gemm(ln1, Wq, None, M, 3 * C, C, bias=bq, D16=qkv)
    if "fc2dg" in which: ops.gemm(g16, W2T, None, M, Fh, C, aux=aux, flags=ops.EPI_GELU_BWD, D16=dh, colsum=cs, colsum_scale=0.5)
    if "attn" in which:
        c16, lse = ops.attention_fwd_f16(qkv_in, B, H, N)
        ops.attention_bwd_f16(qkv_in, lse, c16, ctx, B, H, N)
    if "ln" in which:
        y, mean, rstd = ops.layernorm_fwd_f16(xs, gam, bet, 1e-6)
        dg, db, c1, c2 = (torch.zeros(C, device="cuda") for _ in range(4))
        ops.layernorm_bwd_f16(ctx, 1.0, xs, mean, rstd, gam, r2=x, s2=torch.ones(1, device="cuda"), dgamma=dg, dbeta=db, cs_r1=c1, cs_out=c2)
        ops.layernorm_bwd_f16(ctx, 1.0, xs, mean, rstd, gam, r1=x, r2=x1, s2=torch.ones(1, device="cuda"), dgamma=dg, dbeta=db)
    if "wgrad" in which:
        dW = torch.zeros(C, Fh, device="cuda")
        ops.gemm(ops.operand(g16, mn_major=True), ops.operand(h16, mn_major=True), dW, C, Fh, M, splits=8, flags=ops.GEMM_F16 | ops.EPI_ATOMIC)
for _ in range(3): run()
torch.cuda.synchronize()
torch.cuda.profiler.start()
run()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
