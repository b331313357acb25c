"""Thin torch-tensor wrappers over the C ABI (one function per `uvc_*` entry point).

PyTorch is plumbing here: it owns device memory and the stream; every op below is a launch of a
hand-written sm_100a kernel in libuvc_sm100.so.  Nothing in this module computes with torch ops.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import EPI_ATOMIC, EPI_BIAS, EPI_BLEND, EPI_COLSUM, EPI_GELU, EPI_GELU_BWD, EPI_RESIDUAL, EPI_ROUND_TF32, GEMM_F16, EPI_AUX_F16, GemmArgs, Operand  # noqa: F401


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
         aux=None, ldaux=None, aux_bs=(0, 0), alpha=1.0, beta=1.0, alpha_dev=None, beta_dev=None, flags=0, splits=1, colsum=None, D16=None,
         colsum_scale=0.0, blend=None, R2=None, D2=None):
    """D[z] = epilogue(alpha * A[z] @ B[z]^T) with A:[M,K], B:[N,K] (see include/uvc_b200.h)."""
    lib = _lib.load()
    a = GemmArgs()
    a.M, a.N, a.K = int(M), int(N), int(K)
    a.nb1, a.nb2 = int(batch[0]), int(batch[1])
    a.splits = int(splits)
    a.A = A if isinstance(A, Operand) else operand(A)
    a.B = B if isinstance(B, Operand) else operand(B)
    if D is not None:
        a.D = D.data_ptr()
        a.ldd = int(ldd if ldd is not None else D.stride(-2))
    if D16 is not None:
        assert D16.dtype == torch.float16
        a.D16 = D16.data_ptr(); a.ldd16 = int(D16.stride(-2))
    a.d_bs1, a.d_bs2 = int(d_bs[0]), int(d_bs[1])
    if bias is not None:
        a.bias = bias.data_ptr(); flags |= EPI_BIAS
    if R is not None:
        a.R = R.data_ptr(); a.ldr = int(ldr if ldr is not None else R.stride(-2)); a.r_bs1, a.r_bs2 = int(r_bs[0]), int(r_bs[1])
        flags |= EPI_RESIDUAL
    if aux is not None:
        a.aux = aux.data_ptr(); a.ldaux = int(ldaux if ldaux is not None else aux.stride(-2)); a.aux_bs1, a.aux_bs2 = int(aux_bs[0]), int(aux_bs[1])
        if aux.dtype == torch.float16:
            flags |= EPI_AUX_F16
    a.alpha, a.beta = float(alpha), float(beta)
    a.alpha_dev = alpha_dev.data_ptr() if alpha_dev is not None else None
    a.beta_dev = beta_dev.data_ptr() if beta_dev is not None else None
    if (isinstance(A, torch.Tensor) and A.dtype == torch.float16):
        flags |= GEMM_F16
    if colsum is not None:
        a.colsum = colsum.data_ptr(); flags |= EPI_COLSUM
        a.colsum_scale = float(colsum_scale)
    if blend is not None:       # fused block-gate blend: D = blend[1] * t + blend[0] * R2, D2 = t
        a.blend_dev = blend.data_ptr(); a.R2 = R2.data_ptr(); a.ldr2 = int(R2.stride(-2)); flags |= EPI_BLEND
        if D2 is not None:
            a.D2 = D2.data_ptr(); a.ldd2 = int(D2.stride(-2))
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


# ------------------------------------------------------------------------------------------ row-wise ops
def _p(t):
    """device pointer of a CUDA fp32 tensor (or None)"""
    if t is None:
        return None
    if not (t.is_cuda and t.dtype == torch.float32):
        raise _lib.UvcError("uvc_b200 ops take CUDA fp32 tensors (there is no CPU path)")
    return t.data_ptr()


def _call(name, *args):
    lib = _lib.load()
    _lib.check(getattr(lib, name)(*args, _stream()), name)


def layernorm_fwd(x, gamma, beta, eps, y=None, ldx=None, M=None, save_stats=True, round_tf32=False):
    """x:[M,C] rows (row stride ldx) -> y:[M,C], (mean, rstd)."""
    C_ = gamma.numel()
    M = x.numel() // C_ if M is None else M
    ldx = C_ if ldx is None else ldx
    if y is None:
        y = torch.empty(M, C_, device=x.device)
    mean = torch.empty(M, device=x.device) if save_stats else None
    rstd = torch.empty(M, device=x.device) if save_stats else None
    _call("uvc_layernorm_fwd", _p(x), ldx, _p(gamma), _p(beta), float(eps), _p(y), C_, _p(mean), _p(rstd), M, C_, int(round_tf32))
    return y, mean, rstd


def layernorm_bwd(dy, x, mean, rstd, gamma, r1=None, r2=None, s2=None, dgamma=None, dbeta=None, ldx=None, lddx=None, dx=None, cs_r1=None, cs_out=None):
    C_ = gamma.numel()
    M = mean.numel()
    ldx = C_ if ldx is None else ldx
    lddx = C_ if lddx is None else lddx
    if dx is None:
        dx = torch.empty(M, C_, device=dy.device)
    if cs_r1 is not None or cs_out is not None:
        _call("uvc_layernorm_bwd_cs", _p(dy), C_, _p(x), ldx, _p(mean), _p(rstd), _p(gamma), _p(r1), _p(r2), _p(s2), _p(dx), lddx,
              _p(dgamma), _p(dbeta), _p(cs_r1), _p(cs_out), M, C_)
    else:
        _call("uvc_layernorm_bwd", _p(dy), C_, _p(x), ldx, _p(mean), _p(rstd), _p(gamma), _p(r1), _p(r2), _p(s2), _p(dx), lddx,
              _p(dgamma), _p(dbeta), M, C_)
    return dx


def _p16(t):
    if t is None:
        return None
    if not (t.is_cuda and t.dtype == torch.float16):
        raise _lib.UvcError("expected a CUDA fp16 tensor")
    return t.data_ptr()


def layernorm_fwd_f16(x, gamma, beta, eps, save_stats=True):
    """x:[M,C] fp32 -> y16:[M,C] fp16 (operand of the next kind::f16 GEMM), (mean, rstd)."""
    C_ = gamma.numel()
    M = x.numel() // C_
    y = torch.empty(M, C_, device=x.device, dtype=torch.float16)
    mean = torch.empty(M, device=x.device) if save_stats else None
    rstd = torch.empty(M, device=x.device) if save_stats else None
    _call("uvc_layernorm_fwd_f16", _p(x), C_, _p(gamma), _p(beta), float(eps), _p16(y), C_, _p(mean), _p(rstd), M, C_)
    return y, mean, rstd


def layernorm_bwd_f16(dy16, dy_scale, x, mean, rstd, gamma, r1=None, r2=None, s2=None, dgamma=None, dbeta=None, cs_r1=None, cs_out=None,
                      want_dx16=True, dx16_scale=1.0):
    """dy16 fp16 (scaled by 1/dy_scale) -> (dx fp32, dx16 = fp16(dx16_scale * dx) or None)"""
    C_ = gamma.numel()
    M = mean.numel()
    dx = torch.empty(M, C_, device=x.device)
    dx16 = torch.empty(M, C_, device=x.device, dtype=torch.float16) if want_dx16 else None
    _call("uvc_layernorm_bwd_f16", _p16(dy16), C_, float(dy_scale), _p(x), C_, _p(mean), _p(rstd), _p(gamma), _p(r1), _p(r2), _p(s2), _p(dx), _p16(dx16),
          float(dx16_scale), C_, _p(dgamma), _p(dbeta), _p(cs_r1), _p(cs_out), M, C_)
    return dx, dx16


def cvt_f16(src, transposed=True):
    """fp32 [rows, cols] -> (fp16 copy, fp16 transposed copy [cols, rows] or None)"""
    rows, cols = src.shape
    dst = torch.empty(rows, cols, device=src.device, dtype=torch.float16)
    dstT = torch.empty(cols, rows, device=src.device, dtype=torch.float16) if transposed else None
    _call("uvc_cvt_f16", _p(src), _p16(dst), _p16(dstT), rows, cols)
    return dst, dstT


def softmax_fwd_(S, n, round_tf32=False):
    ld = S.shape[-1]
    _call("uvc_softmax_fwd", _p(S), ld, S.numel() // ld, n, int(round_tf32))
    return S


def softmax_bwd_(P, dP, n, scale, round_tf32=False):
    ld = P.shape[-1]
    _call("uvc_softmax_bwd", _p(P), _p(dP), ld, P.numel() // ld, n, float(scale), int(round_tf32))
    return dP


def colsum_(X, out, scale_dev=None):
    M, N = X.shape
    _call("uvc_colsum", _p(X), X.stride(0), M, N, _p(scale_dev), _p(out))
    return out


def blend_fwd(t, x, d, out=None):
    out = torch.empty_like(x) if out is None else out
    _call("uvc_blend_fwd", _p(t), _p(x), _p(d), _p(out), x.numel())
    return out


def blend_dots_(g, t, x, dots):
    _call("uvc_blend_dots", _p(g), _p(t), _p(x), _p(dots), x.numel())
    return dots


def round_tf32(src, dst=None):
    """dst = src rounded to the nearest TF32 value (10-bit mantissa)."""
    dst = torch.empty_like(src) if dst is None else dst
    _call("uvc_round_tf32", _p(src), _p(dst), src.numel())
    return dst


def im2col16(x, patch=16, round_tf32=False):
    B, Cin, HW, _ = x.shape
    g = HW // patch
    out = torch.empty(B * g * g, Cin * patch * patch, device=x.device)
    _call("uvc_im2col16", _p(x), _p(out), B, Cin, HW, patch, int(round_tf32))
    return out


def assemble_tokens(pe, cls, pos, pscale=None, tmask=None):
    B, np_, C_ = pe.shape
    tok = torch.empty(B, np_ + 1, C_, device=pe.device)
    _call("uvc_assemble_tokens", _p(pe), _p(cls), _p(pos), _p(pscale), _p(tmask), _p(tok), B, np_, C_)
    return tok


def assemble_tokens_bwd(g, pe, pscale=None, tmask=None, want_dpos=True):
    B, np_, C_ = pe.shape
    dpe = torch.empty_like(pe)
    dscale = torch.zeros(np_, device=pe.device) if pscale is not None else None
    dtmask = torch.empty(B, np_, device=pe.device) if tmask is not None else None
    dpos = torch.zeros(np_ + 1, C_, device=pe.device) if want_dpos else None
    dcls = torch.zeros(C_, device=pe.device) if want_dpos else None
    _call("uvc_assemble_tokens_bwd", _p(g), _p(pe), _p(pscale), _p(tmask), _p(dpe), _p(dscale), _p(dtmask), _p(dpos), _p(dcls), B, np_, C_)
    return dpe, dscale, dtmask, dpos, dcls


def token_gate_fold(patch_w, patch_b, gate_w):
    """(v [Kp], c1 [1]) with scores = feat . v + c1 equal to Linear(C,1)(conv(x)) in fp32"""
    C_ = patch_w.shape[0]
    w2 = patch_w.reshape(C_, -1)
    v = torch.empty(w2.shape[1], device=patch_w.device)
    c1 = torch.empty(1, device=patch_w.device)
    _call("uvc_token_gate_fold", _p(w2), _p(patch_b), _p(gate_w), C_, w2.shape[1], _p(v), _p(c1))
    return v, c1


def token_gate_fwd(feat, v, c1, gate_b, pscale, noise, tau, k, B, np_, want_saved=True):
    """-> (mask [B,np], ysoft, ls, scores): Gumbel top-k straight-through token mask, fp32 scores from `feat` rows"""
    Kf = feat.shape[-1]
    dev = feat.device
    mask = torch.empty(B, np_, device=dev)
    ysoft, ls, scores = (torch.empty(B, np_, device=dev) for _ in range(3)) if want_saved else (None, None, None)
    _call("uvc_token_gate_fwd", _p(feat), Kf, Kf, _p(v), _p(c1), _p(gate_b), _p(pscale), _p(noise), float(tau), int(k), B, np_, _p(mask), _p(ysoft), _p(ls), _p(scores))
    return mask, ysoft, ls, scores


def token_gate_bwd(dmask, ysoft, ls, tau):
    B, np_ = ysoft.shape
    ds = torch.empty_like(ysoft)
    _call("uvc_token_gate_bwd", _p(dmask), _p(ysoft), _p(ls), float(tau), B, np_, _p(ds))
    return ds


def token_gate_apply_(dscores, x, gate_w, pscale, dx, d_gate_w, d_gate_b=None, d_pscale=None):
    B, np_, C_ = x.shape
    _call("uvc_token_gate_apply", _p(dscores), _p(x), _p(gate_w), _p(pscale), B, np_, C_, _p(dx), _p(d_gate_w), _p(d_gate_b), _p(d_pscale))


def scale_add_(y, x, s=1.0, s_dev=None):
    _call("uvc_scale_add", _p(y), _p(x), _p(s_dev), float(s), y.numel())
    return y


# ------------------------------------------------------------------------------------------ attention
def attn_ldp(N):
    return int(_lib.load().uvc_attn_ldp(int(N)))


def attention_fwd(qkv, B, H, N, d, scale=None, save_P=True):
    """qkv:[B*N, 3*H*d] -> (ctx:[B*N, H*d], P:[B,H,N,ldp] or None when save_P is False (inference: probabilities stay on chip))"""
    scale = d ** -0.5 if scale is None else scale
    P = torch.zeros(B, H, N, attn_ldp(N), device=qkv.device) if save_P else None
    ctx = torch.empty(B * N, H * d, device=qkv.device)
    _call("uvc_attention_fwd", _p(qkv), _p(P), _p(ctx), B, H, N, d, float(scale))
    return ctx, P


def attention_fwd_lse(qkv, B, H, N, d, scale=None):
    """training forward of the fused path: (ctx, lse[B,H,N]) — no probabilities are materialised"""
    scale = d ** -0.5 if scale is None else scale
    lse = torch.empty(B, H, N, device=qkv.device)
    ctx = torch.empty(B * N, H * d, device=qkv.device)
    _call("uvc_attention_fwd_lse", _p(qkv), _p(lse), _p(ctx), B, H, N, d, float(scale))
    return ctx, lse


def attention_bwd_fused(qkv, lse, ctx, dctx, B, H, N, d, scale=None, dbias=None):
    """dqkv (and, accumulated into `dbias` [3*H*d] if given, the qkv bias gradient = column sums of dqkv)"""
    scale = d ** -0.5 if scale is None else scale
    dqkv = torch.empty_like(qkv)
    ws = torch.empty(B, H, N, device=qkv.device)
    _call("uvc_attention_bwd_fused", _p(qkv), _p(lse), _p(ctx), _p(dctx), _p(ws), _p(dqkv), _p(dbias), B, H, N, d, float(scale))
    return dqkv


def attention_fwd_f16(qkv16, B, H, N, d=64, scale=None, want_lse=True):
    """fp16 operand storage: qkv16 [B*N, 3*H*64] fp16 -> (ctx16 [B*N, H*64] fp16, lse [B,H,N] fp32 or None)"""
    scale = d ** -0.5 if scale is None else scale
    lse = torch.empty(B, H, N, device=qkv16.device) if want_lse else None
    ctx = torch.empty(B * N, H * d, device=qkv16.device, dtype=torch.float16)
    _call("uvc_attention_fwd_f16", _p16(qkv16), _p16(ctx), _p(lse), B, H, N, d, float(scale))
    return ctx, lse


def attention_bwd_f16(qkv16, lse, ctx16, dctx16, B, H, N, d=64, scale=None, dbias=None, db_scale=1.0):
    scale = d ** -0.5 if scale is None else scale
    dqkv = torch.empty_like(qkv16)
    ws = torch.empty(B, H, N, device=qkv16.device)
    _call("uvc_attention_bwd_f16", _p16(qkv16), _p(lse), _p16(ctx16), _p16(dctx16), _p(ws), _p16(dqkv), _p(dbias), float(db_scale), B, H, N, d, float(scale))
    return dqkv


def attention_bwd(qkv, P, dctx, B, H, N, d, scale=None):
    scale = d ** -0.5 if scale is None else scale
    dP = torch.empty_like(P)
    dqkv = torch.empty_like(qkv)
    _call("uvc_attention_bwd", _p(qkv), _p(P), _p(dctx), _p(dP), _p(dqkv), B, H, N, d, float(scale))
    return dqkv


# ------------------------------------------------------------------------------------------ loss / optimiser
def distill_loss(logits, teacher_logits, targets, alpha, T, grad_scale=1.0, want_grad=True):
    """returns (loss_out[3] = (loss, base, kd), dlogits)"""
    B, NC = logits.shape
    out = torch.empty(3, device=logits.device)
    dl = torch.empty_like(logits) if want_grad else None
    _call("uvc_distill_loss", _p(logits), _p(teacher_logits), _p(targets), B, NC, float(alpha), float(T), float(grad_scale), _p(out), _p(dl))
    return out, dl


def sqnorm_accum_(g, acc):
    _call("uvc_sqnorm_accum", _p(g), g.numel(), _p(acc))
    return acc


def clip_adamw_(p, g, m, v, sqnorm_acc, max_norm, lr, beta1, beta2, eps, weight_decay, step, mask=None):
    _call("uvc_clip_adamw", _p(p), _p(g), _p(m), _p(v), _p(mask), p.numel(), _p(sqnorm_acc), float(max_norm), float(lr),
          float(beta1), float(beta2), float(eps), float(weight_decay), int(step))


def sqnorm_accum_flags_(g, flags, acc):
    _call("uvc_sqnorm_accum_flags", _p(g), flags.data_ptr(), g.numel(), _p(acc))
    return acc


def clip_adamw_flags_(p, g, m, v, flags, sqnorm_acc, max_norm, lr, beta1, beta2, eps, weight_decay, step):
    """flat-arena sweep with one option byte per element: bit 0 keep (mask), bit 1 decay, bit 2 active"""
    assert flags.dtype == torch.uint8 and flags.numel() == p.numel()
    _call("uvc_clip_adamw_flags", _p(p), _p(g), _p(m), _p(v), flags.data_ptr(), p.numel(), _p(sqnorm_acc), float(max_norm), float(lr),
          float(beta1), float(beta2), float(eps), float(weight_decay), int(step))


def mixup_(x, y, num_classes, lam, smoothing, box=None):
    """in-place batch mixup (box None) or cutmix (box = (yl, yh, xl, xh)) of x [B,C,H,W] with x.flip(0); returns the mixed smoothed targets [B, NC]"""
    B, C_, H, W = x.shape
    assert x.is_contiguous() and y.dtype == torch.int64 and y.is_cuda
    tgt = torch.empty(B, num_classes, device=x.device)
    yl, yh, xl, xh = box if box is not None else (0, 0, 0, 0)
    _call("uvc_mixup", _p(x), y.data_ptr(), _p(tgt), B, C_, H, W, int(num_classes), float(lam), float(smoothing), 1 if box is not None else 0,
          int(yl), int(yh), int(xl), int(xh))
    return tgt
