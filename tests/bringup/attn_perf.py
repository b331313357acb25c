"""GPU timing of the attention core at the bench shape (B=128, H=6, N=197, d=64): fused forward vs the GEMM-composed path, backward."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from uvc_b200 import ops

B, H, N, d = 128, 6, 197, 64
C = H * d
g = torch.Generator(device="cuda"); g.manual_seed(0)
qkv = ops.round_tf32(torch.randn(B * N, 3 * C, device="cuda", generator=g))
dctx = ops.round_tf32(torch.randn(B * N, C, device="cuda", generator=g))


def timeit(fn, n=20):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

tag = f"fused={os.environ.get('UVC_ATTN_FUSED', '1')}"
print(tag, "fwd save_P : %.1f us" % timeit(lambda: ops.attention_fwd(qkv, B, H, N, d)))
if os.environ.get('UVC_ATTN_FUSED', '1') != '0':
    print(tag, "fwd no P   : %.1f us" % timeit(lambda: ops.attention_fwd(qkv, B, H, N, d, save_P=False)))
ctx, P = ops.attention_fwd(qkv, B, H, N, d)
print(tag, "bwd        : %.1f us" % timeit(lambda: ops.attention_bwd(qkv, P, dctx, B, H, N, d)))
if os.environ.get('UVC_ATTN_FUSED', '1') != '0':
    ctx2, lse = ops.attention_fwd_lse(qkv, B, H, N, d)
    print(tag, "fwd lse    : %.1f us" % timeit(lambda: ops.attention_fwd_lse(qkv, B, H, N, d)))
    print(tag, "bwd fused  : %.1f us" % timeit(lambda: ops.attention_bwd_fused(qkv, lse, ctx2, dctx, B, H, N, d)))
