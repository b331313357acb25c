"""Bring-up: time the four weight-gradient GEMM shapes of a DeiT-Small block (dW[N,K] += dY[M,N]^T X[M,K], M = 25216) under the current
UVC_GEMM_V2 setting (0 = 128x128 kernel only, 1 = default dispatch)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import ctypes as C
import torch
from uvc_b200 import ops, _lib
M = 25216
lib = _lib.load()
def wsplits(n, k, m):
    tiles = ((n + 127) // 128) * ((k + 127) // 128)
    s = (2 * 148) // tiles
    return max(1, min(s, ((m + 31) // 32) // 4))
def timeit(fn, n=20):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for name, N, K in [("qkv_w", 1152, 384), ("proj_w", 384, 384), ("fc1_w", 1536, 384), ("fc2_w", 384, 1536)]:
    dY = ops.round_tf32(torch.randn(M, N, device="cuda")); X = ops.round_tf32(torch.randn(M, K, device="cuda"))
    dW = torch.zeros(N, K, device="cuda")
    sp = wsplits(N, K, M)
    f = lambda: ops.gemm(ops.operand(dY, mn_major=True), ops.operand(X, mn_major=True), dW, N, K, M, flags=_lib.EPI_ATOMIC, splits=sp)
    us = timeit(f)
    print(f"V2={os.environ.get('UVC_GEMM_V2', '1')} {name}: {us:.1f} us  {2.0 * M * N * K / us / 1e6:.0f} TF/s", flush=True)
