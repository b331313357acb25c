"""ncu target: a handful of hot-path GEMM shapes, each launched a few times.  usage: gemm_prof.py [shape-tag ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from uvc_b200 import ops

SHAPES = {   # tag: (M, N, K, a_mn, b_mn, splits, flags)
    "qkv_fwd": (25216, 1152, 384, 0, 0, 1, 0),
    "fc1_fwd": (25216, 1536, 384, 0, 0, 1, ops.EPI_GELU | ops.EPI_BIAS | ops.EPI_ROUND_TF32),
    "fc2_fwd": (25216, 384, 1536, 0, 0, 1, ops.EPI_BIAS | ops.EPI_RESIDUAL),
    "fc2_dgrad": (25216, 1536, 384, 0, 1, 1, ops.EPI_GELU_BWD | ops.EPI_ROUND_TF32),
    "fc1_wgrad": (1536, 384, 25216, 1, 1, 8, ops.EPI_ATOMIC),
    "fc2_wgrad": (384, 1536, 25216, 1, 1, 8, ops.EPI_ATOMIC),
}
dev = "cuda"
g = torch.Generator(device=dev); g.manual_seed(0)
rn = lambda *s: torch.randn(*s, device=dev, generator=g)
tags = sys.argv[1:] or list(SHAPES)
reps = int(os.environ.get("REPS", "3"))
for tag in tags:
    M, N, K, amn, bmn, sp, fl = SHAPES[tag]
    A = rn(K, M) if amn else rn(M, K); B = rn(K, N) if bmn else rn(N, K); D = torch.zeros(M, N, device=dev)
    kw = {}
    if fl & ops.EPI_BIAS: kw["bias"] = rn(N)
    if fl & (ops.EPI_GELU | ops.EPI_GELU_BWD): kw["aux"] = rn(M, N)
    if fl & ops.EPI_RESIDUAL: kw["R"] = rn(M, N)
    fl2 = fl & ~(ops.EPI_BIAS | ops.EPI_RESIDUAL)
    for _ in range(reps):
        ops.gemm(ops.operand(A, mn_major=bool(amn)), ops.operand(B, mn_major=bool(bmn)), D, M, N, K, splits=sp, flags=fl2, **kw)
    torch.cuda.synchronize()
    print("ran", tag, flush=True)
