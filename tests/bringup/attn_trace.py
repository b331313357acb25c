"""Bring-up: per-warp event timeline of attn_fwd_kernel's CTA 0 (library built with UVC_NVCC_EXTRA=-DUVC_ATTN_TRACE)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from uvc_b200 import ops, _lib
B, H, N, d = 128, 6, 197, 64
g = torch.Generator(device="cuda"); g.manual_seed(0)
qkv = ops.round_tf32(torch.randn(B * N, 3 * H * d, device="cuda", generator=g))
for _ in range(5): ops.attention_fwd_lse(qkv, B, H, N, d)
torch.cuda.synchronize()
lib = _lib.load()
buf = (C.c_longlong * 320)()
print("rc", lib.uvc_attn_trace_read(buf))
v = list(buf)
t0 = min(x for x in v if x > 0)
names = {0: "TMA", 1: "MMA"}
for w in range(10):
    for it in range(4):
        ev = [v[(w * 4 + it) * 8 + e] for e in range(8)]
        print(f"w{w} {names.get(w, 'SM%d' % ((w - 2) // 4))} it{it}: " + " ".join(f"{(x - t0) / 1.9e3:7.2f}" if x else "      -" for x in ev))
