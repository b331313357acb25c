"""GPU parity of the 16-bit operand-storage ops (fp16 in HBM, kind::f16 MMAs, fp32 accumulation) against plain PyTorch fp32 math
on the SAME fp16 input values.  Products of fp16 values are exact in fp32, so a GEMM differs from the reference only by summation order;
outputs stored as fp16 carry one extra rounding (2^-11 relative)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

F16_OUT_TOL = 1.5e-3     # result rounded to fp16 once more, relative to the output's max magnitude


def rel(a, b):
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


@pytest.fixture(scope="module")
def ops():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    from uvc_b200 import ops as o
    return o


def rn(*s, seed=0):
    g = torch.Generator(device="cuda"); g.manual_seed(seed + sum(s))
    return torch.randn(*s, device="cuda", generator=g)


@pytest.mark.parametrize("M,N,K,splits", [(384, 1536, 25216, 8), (1152, 384, 25216, 10), (296, 264, 1000, 4), (128, 128, 64, 1), (64, 192, 200, 1),
                                           (1000, 384, 128, 1)])
def test_gemm_fp16_mn_major_splitk(ops, M, N, K, splits):
    """weight-gradient shape: both operands read transposed (MN-major fp16, 128 B swizzle), split-K with fp32 atomics, alpha folded in"""
    Am, Bm = rn(K, M).half(), rn(K, N).half()
    D = torch.zeros(M, N, device="cuda")
    flags = ops.GEMM_F16 | (ops.EPI_ATOMIC if splits > 1 else 0)
    ops.gemm(ops.operand(Am, mn_major=True), ops.operand(Bm, mn_major=True), D, M, N, K, splits=splits, flags=flags, alpha=0.25)
    ref = 0.25 * (Am.float().t() @ Bm.float())
    assert rel(D, ref) < 2e-5


def test_gemm_fp16_mixed_majors(ops):
    """K-major A with MN-major B (a data-gradient GEMM that reads the weight as stored) and the reverse, ragged sizes"""
    M, N, K = 333, 200, 264
    A, Bm = rn(M, K).half(), rn(K, N).half()
    D = torch.zeros(M, N, device="cuda")
    ops.gemm(A, ops.operand(Bm, mn_major=True), D, M, N, K, flags=ops.GEMM_F16)
    assert rel(D, A.float() @ Bm.float()) < 2e-5
    M = 336
    Am, B = rn(K, M).half(), rn(N, K).half()
    D = torch.zeros(M, N, device="cuda")
    ops.gemm(ops.operand(Am, mn_major=True), B, D, M, N, K, flags=ops.GEMM_F16)
    assert rel(D, Am.float().t() @ B.float().t()) < 2e-5


def test_gemm_fp16_d16_only_residual_colsum(ops):
    """the CTA-pair kernel's epilogues on fp16 operands: fp16-only output, fp32 output with residual, GELU' multiply + scaled column sums"""
    M, N, K = 2000, 768, 192
    A, B, bias, R = rn(M, K).half(), (rn(N, K) * 0.1).half(), rn(N), rn(M, N)
    pre = A.float() @ B.float().t()
    D16 = torch.empty(M, N, device="cuda", dtype=torch.float16)
    ops.gemm(A, B, None, M, N, K, bias=bias, D16=D16)
    assert rel(D16.float(), pre + bias) < F16_OUT_TOL
    D = torch.empty(M, N, device="cuda")
    ops.gemm(A, B, D, M, N, K, bias=bias, R=R, beta=0.5)
    assert rel(D, pre + bias + 0.5 * R) < 2e-5
    u = rn(M, N, seed=3).half()
    cs = torch.zeros(N, device="cuda")
    ops.gemm(A, B, None, M, N, K, aux=u, flags=ops.EPI_GELU_BWD, D16=D16, colsum=cs, colsum_scale=0.125)
    want = pre * u.float()
    assert rel(D16.float(), want) < F16_OUT_TOL
    assert rel(cs, 0.125 * want.sum(0)) < 1e-4


@pytest.mark.parametrize("M,C", [(25216, 384), (1000, 192), (788, 768)])
def test_layernorm_f16_variants(ops, M, C):
    x, g, b = rn(M, C) * 2 + 0.3, rn(C), rn(C)
    y, mean, rstd = ops.layernorm_fwd(x, g, b, 1e-6)
    y16, mean16, rstd16 = ops.layernorm_fwd_f16(x, g, b, 1e-6)
    assert torch.equal(mean, mean16) and torch.equal(rstd, rstd16)
    assert torch.equal(y16, y.half())
    S = 1024.0
    dy = rn(M, C, seed=1) * 1e-3
    dy16 = (dy * S).half()
    r1, r2, s2 = rn(M, C, seed=2) * 1e-3, rn(M, C, seed=3) * 1e-3, torch.tensor([0.3], device="cuda")
    dg, db = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    dg2, db2 = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    want = ops.layernorm_bwd(dy16.float() / S, x, mean, rstd, g, r1=r1, r2=r2, s2=s2, dgamma=dg, dbeta=db)
    dx, dx16 = ops.layernorm_bwd_f16(dy16, 1.0 / S, x, mean, rstd, g, r1=r1, r2=r2, s2=s2, dgamma=dg2, dbeta=db2, dx16_scale=S)
    assert rel(dx, want) < 1e-6 and rel(dg2, dg) < 1e-5 and rel(db2, db) < 1e-5
    assert rel(dx16.float() / S, want) < 1e-3


def test_cvt_f16_with_transpose(ops):
    for rows, cols in [(1152, 384), (384, 1536), (1000, 192), (33, 70)]:
        w = rn(rows, cols)
        a, at = ops.cvt_f16(w)
        assert torch.equal(a, w.half()) and torch.equal(at, w.half().t().contiguous())
        a2, none = ops.cvt_f16(w, transposed=False)
        assert none is None and torch.equal(a2, a)


def _attn_ref(qkv, B, H, N, d):
    t = qkv.view(B, N, 3, H, d).permute(2, 0, 3, 1, 4)
    s_ = (t[0] @ t[1].transpose(-2, -1)) * d ** -0.5
    return s_, (s_.softmax(-1) @ t[2]).transpose(1, 2).reshape(B * N, H * d)


@pytest.mark.parametrize("B,H,N", [(2, 6, 197), (64, 6, 197), (3, 3, 50), (5, 2, 130), (2, 4, 208), (40, 12, 197), (1, 1, 1)])
def test_attention_f16_forward(ops, B, H, N):
    d = 64
    qkv16 = rn(B * N, 3 * H * d).half()
    s_, ref = _attn_ref(qkv16.float(), B, H, N, d)
    ctx, lse = ops.attention_fwd_f16(qkv16, B, H, N)
    assert rel(ctx.float(), ref) < 2e-3
    assert rel(lse * 0.6931471805599453, torch.logsumexp(s_, -1)) < 1e-4
    ctx2, none = ops.attention_fwd_f16(qkv16, B, H, N, want_lse=False)
    assert none is None and torch.equal(ctx2, ctx)


@pytest.mark.parametrize("B,H,N", [(2, 6, 197), (64, 6, 197), (3, 3, 50), (5, 2, 130), (2, 4, 208)])
def test_attention_f16_backward(ops, B, H, N):
    """fp16 forward (lse) + recompute backward against autograd of the plain fp32 formula on the same fp16 values; the upstream gradient
    carries a loss scale that the bias-gradient output takes back out"""
    d = 64
    C = H * d
    S = 256.0
    qkv16 = rn(B * N, 3 * C).half()
    ctx, lse = ops.attention_fwd_f16(qkv16, B, H, N)
    q = qkv16.float().requires_grad_(True)
    _, ref = _attn_ref(q, B, H, N, d)
    dctx = rn(B * N, C, seed=7) * 0.01
    dctx16 = (dctx * S).half()
    ref.backward(dctx16.float() / S)
    db = torch.ones(3 * C, device="cuda")
    dqkv = ops.attention_bwd_f16(qkv16, lse, ctx, dctx16, B, H, N, dbias=db, db_scale=1.0 / S)
    assert rel(db - 1, q.grad.sum(0)) < 6e-3
    g = q.grad.view(B * N, 3, C)
    got = dqkv.float().view(B * N, 3, C) / S
    for i, nm in enumerate("qkv"):
        assert rel(got[:, i], g[:, i]) < 6e-3, nm


def test_attention_f16_propagates_nan(ops):
    B, H, N, d = 2, 3, 197, 64
    qkv = rn(B * N, 3 * H * d).half()
    qkv[5, 64 + 3] = float("nan")                    # image 0, token 5, head 1 of q
    ctx, lse = ops.attention_fwd_f16(qkv, B, H, N)
    assert torch.isnan(ctx[5, 64:128]).all() and torch.isnan(lse.view(B, H, N)[0, 1, 5])
    assert not torch.isnan(ctx[6]).any() and not torch.isnan(ctx[5, :64]).any()


@pytest.mark.parametrize("M,C,use_r1,use_r2,cs", [(8195, 384, True, True, False), (4099, 384, False, True, True), (2051, 768, True, False, True),
                                                  (1027, 192, True, True, True), (25216, 384, False, True, True)])
def test_layernorm_bwd_streamed_kernel_equals_register_kernel(ops, monkeypatch, M, C, use_r1, use_r2, cs):
    """The engine's LayerNorm backward at bench sizes runs `layernorm_bwd_stream_kernel` (bulk-asynchronous staging of 8-row stages through shared
    memory); small / strided problems run the register-resident kernel.  Same arithmetic: dx, the fp16 operand copy, dgamma / dbeta and the fused
    column sums must agree to summation order, including a ragged last stage (M % 8 != 0) and rows wider / narrower than the bench's."""
    x, g, b = rn(M, C) * 2 + 0.3, rn(C), rn(C)
    _, mean, rstd = ops.layernorm_fwd_f16(x, g, b, 1e-6)
    S = 256.0
    dy16 = (rn(M, C, seed=1) * 1e-3 * S).half()
    r1 = rn(M, C, seed=2) * 1e-3 if use_r1 else None
    r2 = rn(M, C, seed=3) * 1e-3 if use_r2 else None
    s2 = torch.tensor([0.7], device="cuda") if use_r2 else None
    out = {}
    for mode in ("reg", "stream"):
        if mode == "reg":
            monkeypatch.setenv("UVC_LN_BWD_REG", "1")
        else:
            monkeypatch.delenv("UVC_LN_BWD_REG", raising=False)
        dg, db, c1, c2 = (torch.zeros(C, device="cuda") for _ in range(4))
        dx, dx16 = ops.layernorm_bwd_f16(dy16, 1.0 / S, x, mean, rstd, g, r1=r1, r2=r2, s2=s2, dgamma=dg, dbeta=db, cs_r1=c1 if cs else None,
                                         cs_out=c2 if cs else None, dx16_scale=S)
        out[mode] = (dx, dx16.float(), dg, db, c1, c2)
    names = ("dx", "dx16", "dgamma", "dbeta", "cs_r1", "cs_out")
    for n, a, bb in zip(names, out["stream"], out["reg"]):
        if bb.abs().max() == 0:
            assert a.abs().max() == 0, n
        else:
            # dx: summation order only; dx16: a 1-ulp difference of dx may flip an fp16 rounding (2^-11 of that element); sums: atomics order
            assert rel(a, bb) < {"dx": 2e-6, "dx16": 6e-4}.get(n, 2e-4), (n, rel(a, bb))
