"""Thin torch-tensor wrappers over the C ABI (one function per `uvc_*` entry point).

PyTorch is plumbing here: it owns device memory and the stream; every op below is a launch of a
hand-written sm_100a kernel in libuvc_sm100.so.  Nothing in this module computes with torch ops.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import EPI_ATOMIC, EPI_BIAS, EPI_GELU, EPI_GELU_BWD, EPI_RESIDUAL, GemmArgs, Operand  # noqa: F401


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    if t is None:
        return None
    assert t.is_cuda and t.dtype == torch.float32, "uvc_b200 ops take CUDA fp32 tensors"
    return C.c_void_p(t.data_ptr())


def operand(t, ld=None, bs1=0, bs2=0, mn_major=False):
    """Describe a GEMM operand living in tensor `t` (pointer = t.data_ptr())."""
    if ld is None:
        assert t.dim() == 2 and t.stride(1) == 1
        ld = t.stride(0)
    return Operand(t.data_ptr(), int(ld), int(bs1), int(bs2), 1 if mn_major else 0, 0)


def gemm(A, B, D, M, N, K, *, ldd=None, d_bs=(0, 0), batch=(1, 1), bias=None, R=None, ldr=None, r_bs=(0, 0),
         aux=None, ldaux=None, aux_bs=(0, 0), alpha=1.0, beta=1.0, alpha_dev=None, beta_dev=None, flags=0, splits=1):
    """D[z] = epilogue(alpha * A[z] @ B[z]^T) with A:[M,K], B:[N,K] (see include/uvc_b200.h)."""
    lib = _lib.load()
    a = GemmArgs()
    a.M, a.N, a.K = int(M), int(N), int(K)
    a.nb1, a.nb2 = int(batch[0]), int(batch[1])
    a.splits = int(splits)
    a.A = A if isinstance(A, Operand) else operand(A)
    a.B = B if isinstance(B, Operand) else operand(B)
    a.D = D.data_ptr()
    a.ldd = int(ldd if ldd is not None else D.stride(-2))
    a.d_bs1, a.d_bs2 = int(d_bs[0]), int(d_bs[1])
    if bias is not None:
        a.bias = bias.data_ptr(); flags |= EPI_BIAS
    if R is not None:
        a.R = R.data_ptr(); a.ldr = int(ldr if ldr is not None else R.stride(-2)); a.r_bs1, a.r_bs2 = int(r_bs[0]), int(r_bs[1])
        flags |= EPI_RESIDUAL
    if aux is not None:
        a.aux = aux.data_ptr(); a.ldaux = int(ldaux if ldaux is not None else aux.stride(-2)); a.aux_bs1, a.aux_bs2 = int(aux_bs[0]), int(aux_bs[1])
    a.alpha, a.beta = float(alpha), float(beta)
    a.alpha_dev = alpha_dev.data_ptr() if alpha_dev is not None else None
    a.beta_dev = beta_dev.data_ptr() if beta_dev is not None else None
    a.flags = int(flags)
    _lib.check(lib.uvc_gemm_tf32(C.byref(a), _stream()), "uvc_gemm_tf32")
    return D


def linear(x, w, bias=None, out=None, **kw):
    """y = x @ w^T + bias   (x:[M,K], w:[N,K]) — nn.Linear forward."""
    M, K = x.shape
    N = w.shape[0]
    if out is None:
        out = torch.empty(M, N, device=x.device, dtype=torch.float32)
    return gemm(x, w, out, M, N, K, bias=bias, **kw)
