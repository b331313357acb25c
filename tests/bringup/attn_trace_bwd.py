"""Bring-up: per-warp event timeline of attn_bwd_kernel<PHASE>'s CTA 0, work items 2..5 (library built with UVC_NVCC_EXTRA=-DUVC_ATTN_TRACE).
The trace buffer is shared by both phases: the kernel that ran last wins, so UVC_TRACE_PHASE=1 stops after phase 1 by timing only that call order."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from uvc_b200 import ops, _lib
B, H, N, d = 128, 6, 197, 64
g = torch.Generator(device="cuda"); g.manual_seed(0)
qkv = ops.round_tf32(torch.randn(B * N, 3 * H * d, device="cuda", generator=g))
dctx = ops.round_tf32(torch.randn(B * N, H * d, device="cuda", generator=g))
ctx, lse = ops.attention_fwd_lse(qkv, B, H, N, d)
for _ in range(3): ops.attention_bwd_fused(qkv, lse, ctx, dctx, B, H, N, d)
torch.cuda.synchronize()
lib = _lib.load()
buf = (C.c_longlong * 320)()
lib.uvc_attn_trace_read(buf)
v = list(buf)
t0 = min(x for x in v if x > 0)
for w in (0, 1, 2, 6):
    for it in range(4):
        ev = [v[(w * 4 + it) * 8 + e] for e in range(8)]
        print(f"w{w} item{it + 2}: " + " ".join(f"{(x - t0) / 1.9e3:7.2f}" if x else "      -" for x in ev))
