"""GPU parity of every C-ABI op against plain PyTorch fp32 on the same inputs (TF32 tensor-core ops: 3e-3 of
the output's max magnitude; fp32 CUDA-core ops: 2e-5)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

TF32_TOL = 3e-3     # one TF32 GEMM: inputs carry 10 mantissa bits
FP32_TOL = 2e-5


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


@pytest.mark.parametrize("M,N,K", [(256, 128, 64), (1576, 576, 192), (130, 36, 200), (1, 4, 4), (8, 1000, 192)])
def test_gemm_nt(ops, M, N, K):
    A, B = rn(M, K), rn(N, K)
    ldd = (N + 3) // 4 * 4
    D = torch.zeros(M, ldd, device="cuda")
    ops.gemm(A, B, D, M, N, K, ldd=ldd)
    assert rel(D[:, :N], A @ B.t()) < TF32_TOL


def test_gemm_transposed_operands_and_splitk(ops):
    M, N, K = 300, 260, 1000
    Am, Bm = rn(K, M), rn(K, N)
    D = torch.zeros(M, N, device="cuda")
    ops.gemm(ops.operand(Am, mn_major=True), ops.operand(Bm, mn_major=True), D, M, N, K, splits=4, flags=ops.EPI_ATOMIC)
    assert rel(D, Am.t() @ Bm) < TF32_TOL
    A = rn(M, K)
    ops.gemm(A, ops.operand(Bm, mn_major=True), D, M, N, K)
    assert rel(D, A @ Bm) < TF32_TOL


def test_gemm_epilogues(ops):
    M, N, K = 1000, 768, 192
    A, B, bias, R = rn(M, K), rn(N, K) * 0.1, rn(N), rn(M, N)
    pre = A @ B.t() + bias
    D = torch.empty(M, N, device="cuda"); aux = torch.empty(M, N, device="cuda")
    ops.gemm(A, B, D, M, N, K, bias=bias, aux=aux, flags=ops.EPI_GELU)
    pp = pre.clone().requires_grad_(True); F.gelu(pp).sum().backward()
    assert rel(D, F.gelu(pre)) < TF32_TOL and rel(aux, pp.grad) < TF32_TOL      # aux = gelu'(pre-activation)
    ops.gemm(A, B, D, M, N, K, bias=bias, R=R, beta=0.5)
    assert rel(D, pre + 0.5 * R) < TF32_TOL
    u = rn(M, N, seed=3)
    ops.gemm(A, B, D, M, N, K, aux=u, flags=ops.EPI_GELU_BWD)                  # multiply by the saved derivative
    assert rel(D, (A @ B.t()) * u) < TF32_TOL


def test_gemm_gelu_pair_with_fp16_aux(ops):
    """UVC_EPI_AUX_F16: the forward epilogue stores gelu'(pre-activation) as fp16, the backward epilogue multiplies by it (ragged M, N)."""
    M, N, K = 1000, 772, 192
    A, B, bias = rn(M, K), rn(N, K) * 0.1, rn(N)
    pre = A @ B.t() + bias
    D = torch.empty(M, N, device="cuda"); aux = torch.zeros(M, N, device="cuda", dtype=torch.float16)
    ops.gemm(A, B, D, M, N, K, bias=bias, aux=aux, flags=ops.EPI_GELU)
    pp = pre.clone().requires_grad_(True); F.gelu(pp).sum().backward()
    assert rel(D, F.gelu(pre)) < TF32_TOL and rel(aux.float(), pp.grad) < TF32_TOL
    G, W = rn(M, K, seed=5), rn(N, K, seed=6) * 0.1                      # dh = (G W^T) .* gelu'
    dh = torch.empty(M, N, device="cuda")
    ops.gemm(G, W, dh, M, N, K, aux=aux, flags=ops.EPI_GELU_BWD)
    assert rel(dh, (G @ W.t()) * pp.grad) < TF32_TOL


def test_fused_attention_propagates_nan(ops):
    """A NaN in q must come out as NaN in ctx and lse (the integer TF32 rounding of the TMEM operands must not launder it into a zero)."""
    B, H, N, d = 2, 3, 197, 64
    qkv = ops.round_tf32(rn(B * N, 3 * H * d))
    qkv[5, 64 + 3] = float("nan")                    # image 0, token 5, head 1 of q
    ctx, lse = ops.attention_fwd_lse(qkv, B, H, N, d)
    assert torch.isnan(ctx[5, 64:128]).all() and torch.isnan(lse.view(B, H, N)[0, 1, 5])
    assert not torch.isnan(ctx[6]).any() and not torch.isnan(ctx[5, :64]).any()
    qkv2 = ops.round_tf32(rn(B * N, 3 * H * d))
    qkv2[7, H * d + 2 * 64 + 1] = float("nan")       # image 0, key token 7, head 2: every query of that head sees it
    ctx2, _ = ops.attention_fwd_lse(qkv2, B, H, N, d)
    assert torch.isnan(ctx2[:N, 128:192]).all() and not torch.isnan(ctx2[N:]).any()


def test_gemm_rejects_bad_arguments(ops):
    from uvc_b200._lib import UvcError
    A, B = rn(16, 6), rn(16, 6)          # ld = 6 is not a multiple of 4
    with pytest.raises(UvcError):
        ops.gemm(A, B, torch.empty(16, 16, device="cuda"), 16, 16, 6)
    with pytest.raises(UvcError):
        ops.gemm(rn(16, 8), rn(16, 8), torch.empty(16, 16, device="cuda"), 16, 16, 8, splits=2)   # split-K without ATOMIC


@pytest.mark.parametrize("M,C,eps", [(1576, 192, 1e-6), (394, 384, 1e-6), (197, 768, 1e-5), (3, 1024, 1e-6)])
def test_layernorm_fwd_bwd(ops, M, C, eps):
    x, g, b, dy = rn(M, C) * 2 + 0.3, 1 + 0.1 * rn(C), 0.1 * rn(C, seed=1), rn(M, C, seed=2)
    y, mean, rstd = ops.layernorm_fwd(x, g, b, eps)
    xr, gr, br = x.clone().requires_grad_(True), g.clone().requires_grad_(True), b.clone().requires_grad_(True)
    yr = F.layer_norm(xr, (C,), gr, br, eps)
    assert rel(y, yr.detach()) < FP32_TOL
    yr.backward(dy)
    r1, r2, s2 = rn(M, C, seed=4), rn(M, C, seed=5), torch.tensor([0.37], device="cuda")
    dg, db = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    dx = ops.layernorm_bwd(dy, x, mean, rstd, g, r1=r1, r2=r2, s2=s2, dgamma=dg, dbeta=db)
    assert rel(dx, xr.grad + r1 + 0.37 * r2) < FP32_TOL
    assert rel(dg, gr.grad) < 1e-4 and rel(db, br.grad) < 1e-4


def test_layernorm_strided_cls_rows(ops):
    B, N, C = 8, 197, 192
    x, g, b = rn(B, N, C), 1 + 0.1 * rn(C), 0.1 * rn(C, seed=1)
    y, mean, rstd = ops.layernorm_fwd(x, g, b, 1e-6, ldx=N * C, M=B)
    assert rel(y, F.layer_norm(x[:, 0], (C,), g, b, 1e-6)) < FP32_TOL


def test_softmax_fwd_bwd(ops):
    rows, n, ld = 4 * 3 * 197, 197, 200
    S = torch.full((rows, ld), float("nan"), device="cuda"); S[:, :n] = rn(rows, n) * 3
    ref = torch.softmax(S[:, :n], -1)
    P = ops.softmax_fwd_(S.clone(), n)
    assert rel(P[:, :n], ref) < FP32_TOL and torch.isnan(P[:, n:]).all()
    dP = torch.zeros(rows, ld, device="cuda"); dP[:, :n] = rn(rows, n, seed=2)
    s = S[:, :n].clone().requires_grad_(True)
    (torch.softmax(s * 0.125, -1) * dP[:, :n]).sum().backward()
    P2 = torch.zeros(rows, ld, device="cuda"); P2[:, :n] = torch.softmax(S[:, :n] * 0.125, -1)
    dS = ops.softmax_bwd_(P2, dP.clone(), n, 0.125)
    assert rel(dS[:, :n], s.grad) < 1e-4


def test_colsum_blend_scale_add(ops):
    X = rn(1576, 576)
    out = torch.zeros(576, device="cuda")
    ops.colsum_(X, out)
    assert rel(out, X.sum(0)) < 1e-4
    t, x, g = rn(8, 197, 192), rn(8, 197, 192, seed=1), rn(8, 197, 192, seed=2)
    d = torch.tensor([0.3, 0.7], device="cuda")
    assert rel(ops.blend_fwd(t, x, d), 0.7 * t + 0.3 * x) < FP32_TOL
    dots = torch.zeros(2, device="cuda")
    ops.blend_dots_(g, t, x, dots)
    assert rel(dots, torch.stack([(g * x).sum(), (g * t).sum()])) < 1e-4
    y = x.clone()
    ops.scale_add_(y, t, 0.5, torch.tensor([2.0], device="cuda"))
    assert rel(y, x + t) < FP32_TOL


def test_im2col_and_token_assembly(ops):
    B, C = 3, 192
    x = rn(B, 3, 224, 224)
    cols = ops.im2col16(x, 16)
    ref = F.unfold(x, 16, stride=16).transpose(1, 2).reshape(B * 196, 768)
    assert torch.equal(cols, ref)
    pe, cls, pos = rn(B, 196, C), rn(C), rn(197, C, seed=1)
    ps, tm = torch.sigmoid(rn(196, seed=2)), (rn(B, 196, seed=3) > 0).float()
    tok = ops.assemble_tokens(pe, cls, pos, ps, tm)
    per, psr, tmr = pe.clone().requires_grad_(True), ps.clone().requires_grad_(True), tm.clone().requires_grad_(True)
    clsr, posr = cls.clone().requires_grad_(True), pos.clone().requires_grad_(True)
    ref = torch.cat([clsr.expand(B, 1, C), per * psr.view(1, -1, 1) * tmr.unsqueeze(-1)], 1) + posr
    assert rel(tok, ref.detach()) < FP32_TOL
    g = rn(B, 197, C, seed=4)
    ref.backward(g)
    dpe, dscale, dtmask, dpos, dcls = ops.assemble_tokens_bwd(g, pe, ps, tm)
    assert rel(dpe, per.grad) < FP32_TOL and rel(dscale, psr.grad) < 1e-4 and rel(dtmask, tmr.grad) < 1e-4
    assert rel(dpos, posr.grad) < 1e-4 and rel(dcls, clsr.grad) < 1e-4
    tok2 = ops.assemble_tokens(pe, cls, pos)
    assert rel(tok2, torch.cat([cls.expand(B, 1, C), pe], 1) + pos) < FP32_TOL


@pytest.mark.parametrize("B,H", [(4, 3), (2, 6)])
def test_attention_fwd_bwd(ops, B, H):
    N, d = 197, 64
    C = H * d
    qkv = ops.round_tf32(rn(B * N, 3 * C))      # in the engine the QKV GEMM epilogue rounds its output to TF32
    ctx, P = ops.attention_fwd(qkv, B, H, N, d)
    q = qkv.clone().requires_grad_(True)
    t = q.view(B, N, 3, H, d).permute(2, 0, 3, 1, 4)
    attn = ((t[0] @ t[1].transpose(-2, -1)) * d ** -0.5).softmax(-1)
    ref = (attn @ t[2]).transpose(1, 2).reshape(B * N, C)
    assert rel(ctx, ref.detach()) < TF32_TOL
    assert rel(P[..., :N], attn.detach()) < TF32_TOL
    dctx = ops.round_tf32(rn(B * N, C, seed=7))
    ref.backward(dctx)
    dqkv = ops.attention_bwd(qkv, P, dctx, B, H, N, d)
    assert rel(dqkv, q.grad) < 2 * TF32_TOL


@pytest.mark.parametrize("T,alpha,kd", [(1.0, 0.1, True), (3.0, 0.5, True), (1.0, 0.0, False)])
def test_distill_loss(ops, T, alpha, kd):
    B, NC = 16, 1000
    s, t = rn(B, NC) * 2, rn(B, NC, seed=1) * 2
    y = torch.softmax(rn(B, NC, seed=2) * 3, -1)
    sr = s.clone().requires_grad_(True)
    base = torch.sum(-y * F.log_softmax(sr, -1), -1).mean()
    if kd:
        k = F.kl_div(F.log_softmax(sr / T, 1), F.log_softmax(t / T, 1), reduction="sum", log_target=True) * (T * T) / sr.numel()
        loss = base * (1 - alpha) + k * alpha
    else:
        k, loss = torch.zeros(()), base
    loss.backward()
    out, dl = ops.distill_loss(s, t if kd else None, y, alpha, T)
    assert abs(out[0].item() - loss.item()) < 1e-5 * max(1, abs(loss.item()))
    assert abs(out[1].item() - base.item()) < 1e-5 * max(1, abs(base.item())) and abs(out[2].item() - float(k)) < 1e-6
    assert rel(dl, sr.grad) < 1e-4


def test_clip_adamw_matches_torch(ops):
    torch.manual_seed(0)
    shapes = [(1000, 192), (577,), (3, 5, 7)]
    ps = [torch.randn(s, device="cuda") for s in shapes]
    ref = [p.clone().requires_grad_(True) for p in ps]
    opt = torch.optim.AdamW(ref, lr=1e-3, weight_decay=0.05)
    n = sum((p.numel() + 3) // 4 * 4 for p in ps)
    flat_p, flat_g = torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    flat_m, flat_v = torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    offs, o = [], 0
    for p in ps:
        flat_p[o:o + p.numel()] = p.flatten(); offs.append(o); o += (p.numel() + 3) // 4 * 4
    for step in range(1, 4):
        gs = [torch.randn_like(p) * 3 for p in ps]
        for r, g, p, o in zip(ref, gs, ps, offs):
            r.grad = g.clone(); flat_g[o:o + p.numel()] = g.flatten()
        total = torch.nn.utils.clip_grad_norm_(ref, 1.0)
        opt.step()
        acc = torch.zeros(1, device="cuda")
        ops.sqnorm_accum_(flat_g, acc)
        assert abs(acc.sqrt().item() - total.item()) < 1e-4 * total.item()
        ops.clip_adamw_(flat_p, flat_g, flat_m, flat_v, acc, 1.0, 1e-3, 0.9, 0.999, 1e-8, 0.05, step)
        for r, p, o in zip(ref, ps, offs):
            assert rel(flat_p[o:o + p.numel()].view(p.shape), r.detach()) < 1e-5

@pytest.mark.parametrize("M,N,K", [(1000, 768, 192), (2048, 384, 100), (130, 36, 64)])
def test_gemm_colsum_epilogue(ops, M, N, K):
    """UVC_EPI_COLSUM: the bias gradient (column sums of the GEMM output) accumulated by the epilogue, plain and after GELU'."""
    A, B = rn(M, K), rn(N, K) * 0.1
    D = torch.empty(M, N, device="cuda"); cs = torch.ones(N, device="cuda")
    ops.gemm(A, B, D, M, N, K, colsum=cs)
    ref = A @ B.t()
    assert rel(D, ref) < TF32_TOL and rel(cs - 1, ref.sum(0)) < TF32_TOL
    u = rn(M, N, seed=5)
    cs2 = torch.zeros(N, device="cuda")
    ops.gemm(A, B, D, M, N, K, aux=u, flags=ops.EPI_GELU_BWD | ops.EPI_ROUND_TF32, colsum=cs2)
    ref2 = ref * u
    assert rel(D, ref2) < TF32_TOL and rel(cs2, ref2.sum(0)) < TF32_TOL


@pytest.mark.parametrize("M,N,K,amn,bmn,splits", [(25216, 1152, 384, 0, 0, 1), (25216, 384, 1536, 0, 1, 1), (1536, 384, 5000, 1, 1, 8), (1000, 388, 100, 0, 0, 1),
                                                   (776, 512, 96, 1, 0, 1)])
def test_gemm_cta_pair_kernel_shapes(ops, M, N, K, amn, bmn, splits):
    """hot-path shapes that dispatch to the persistent CTA-pair kernel (all operand majors, ragged edges, split-K)"""
    A = rn(K, M) if amn else rn(M, K); B = rn(K, N) if bmn else rn(N, K)
    D = torch.zeros(M, N, device="cuda")
    ops.gemm(ops.operand(A, mn_major=bool(amn)), ops.operand(B, mn_major=bool(bmn)), D, M, N, K, splits=splits, flags=ops.EPI_ATOMIC if splits > 1 else 0)
    ref = (A.t() if amn else A) @ (B if bmn else B.t())
    assert rel(D, ref) < TF32_TOL


def test_layernorm_bwd_fused_column_sums(ops):
    """uvc_layernorm_bwd_cs also emits the bias gradients of the neighbouring Linears (column sums of r1 and of dx)"""
    M, Cc = 1576, 384
    x, g, dy, r1 = rn(M, Cc) * 2 + 0.3, 1 + 0.1 * rn(Cc), rn(M, Cc, seed=2), rn(M, Cc, seed=3)
    y, mean, rstd = ops.layernorm_fwd(x, g, 0.1 * rn(Cc, seed=1), 1e-6)
    dx0 = ops.layernorm_bwd(dy, x, mean, rstd, g, r1=r1)
    c1, co = torch.zeros(Cc, device="cuda"), torch.ones(Cc, device="cuda")
    dg, db = torch.zeros(Cc, device="cuda"), torch.zeros(Cc, device="cuda")
    dx = ops.layernorm_bwd(dy, x, mean, rstd, g, r1=r1, dgamma=dg, dbeta=db, cs_r1=c1, cs_out=co)
    assert torch.equal(dx, dx0)
    assert rel(c1, r1.sum(0)) < FP32_TOL and rel(co - 1, dx0.double().sum(0).float()) < 1e-4 and rel(db, dy.sum(0)) < FP32_TOL


@pytest.mark.parametrize("B,H,N", [(64, 6, 197), (3, 3, 50), (5, 2, 130), (2, 4, 208), (40, 12, 197)])
def test_attention_fused_forward(ops, B, H, N):
    """fused tcgen05 attention forward (scores / probabilities stay in tensor memory): with and without the saved probabilities,
    more heads than SMs (persistent loop), one- and two-tile sequences"""
    d = 64
    C = H * d
    qkv = ops.round_tf32(rn(B * N, 3 * C))
    t = qkv.view(B, N, 3, H, d).permute(2, 0, 3, 1, 4)
    attn = ((t[0] @ t[1].transpose(-2, -1)) * d ** -0.5).softmax(-1)
    ref = (attn @ t[2]).transpose(1, 2).reshape(B * N, C)
    ctx, P = ops.attention_fwd(qkv, B, H, N, d)
    assert rel(ctx, ref) < TF32_TOL
    assert rel(P[..., :N], attn) < TF32_TOL
    assert (P[..., N:] == 0).all()
    ctx2, none = ops.attention_fwd(qkv, B, H, N, d, save_P=False)
    assert none is None and torch.equal(ctx2, ctx)


@pytest.mark.parametrize("B,H,N", [(2, 6, 197), (64, 6, 197), (3, 3, 50), (5, 2, 130), (2, 4, 208)])
def test_attention_fused_backward(ops, B, H, N):
    """fused forward (lse) + fused recompute backward against autograd of the plain fp32 formula"""
    d = 64
    C = H * d
    qkv = ops.round_tf32(rn(B * N, 3 * C))
    ctx, lse = ops.attention_fwd_lse(qkv, B, H, N, d)
    q = qkv.clone().requires_grad_(True)
    t = q.view(B, N, 3, H, d).permute(2, 0, 3, 1, 4)
    s_ = (t[0] @ t[1].transpose(-2, -1)) * d ** -0.5
    ref = (s_.softmax(-1) @ t[2]).transpose(1, 2).reshape(B * N, C)
    assert rel(ctx, ref.detach()) < TF32_TOL
    assert rel(lse * 0.6931471805599453, torch.logsumexp(s_.detach(), -1)) < 1e-4
    dctx = ops.round_tf32(rn(B * N, C, seed=7))
    ref.backward(dctx)
    db = torch.ones(3 * C, device="cuda")
    dqkv = ops.attention_bwd_fused(qkv, lse, ctx, dctx, B, H, N, d, dbias=db)
    assert rel(db - 1, q.grad.sum(0)) < 2 * TF32_TOL          # fused qkv bias gradient (column sums of dqkv)
    g = q.grad.view(B * N, 3, C)
    got = dqkv.view(B * N, 3, C)
    for i, nm in enumerate("qkv"):
        assert rel(got[:, i], g[:, i]) < 2 * TF32_TOL, nm


@pytest.mark.parametrize("M,N,K", [(25216, 1152, 384), (1000, 388, 104), (25216, 384, 1536), (300, 64, 64)])
def test_gemm_fp16_operands(ops, M, N, K):
    """fp16 operand storage (kind::f16, fp32 accumulate): fp32 and fp16 outputs, fused bias + GELU, against fp32 math on the same fp16 values"""
    A, B = rn(M, K).half(), (rn(N, K) * 0.1).half()
    bias = rn(N)
    D = torch.empty(M, N, device="cuda"); D16 = torch.empty(M, N, device="cuda", dtype=torch.float16)
    ops.gemm(A, B, D, M, N, K, bias=bias, D16=D16)
    ref = A.float() @ B.float().t() + bias
    assert rel(D, ref) < 1e-5 * max(1, K // 64) + 2e-6           # exact products, fp32 accumulation: only summation-order error
    assert rel(D16.float(), ref) < 1e-3
    aux = torch.empty(M, N, device="cuda")
    ops.gemm(A, B, None, M, N, K, bias=bias, aux=aux, flags=ops.EPI_GELU, D16=D16)
    assert rel(D16.float(), F.gelu(ref)) < 1.5e-3
