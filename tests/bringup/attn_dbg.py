import os, sys
sys.path.insert(0, "/root/repo")
import torch
from uvc_b200 import ops
B, H, N, d = 128, 6, 197, 64
C = H * d
g = torch.Generator(device="cuda"); g.manual_seed(0)
qkv = ops.round_tf32(torch.randn(B * N, 3 * C, device="cuda", generator=g))
def timeit(fn, n=30):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for dbg in [0, 1, 2, 3, 4, 8, 16, 24, 7, 31]:
    os.environ["UVC_ATTN_DBG"] = str(dbg)
    print("dbg", dbg, "fwd lse: %.1f us" % timeit(lambda: ops.attention_fwd_lse(qkv, B, H, N, d)), flush=True)
