"""Token (patch) slimming gate, mode 2 of `--enable_patch_gating` (reference models/model_distilled.py:446-456):

    scores = Linear(C,1)(patch_embed(x) [* patch gate])            -> [B, 196]
    mask   = straight-through top-k of softmax((log_softmax(scores) + Gumbel) / tau)
    mask[:, 0] = 1                                                  (hits patch 0, as in the reference)

Only the Gumbel draw stays in torch (two tiny launches: it must consume the reference's random stream); scores, log-softmax, softmax, the
exact-rank top-k and the straight-through mask are ONE kernel per forward (`uvc_token_gate_fwd`, one CTA per image, no `.tolist()` round trip),
and the whole backward of the gate is two (`uvc_token_gate_bwd`, `uvc_token_gate_apply`).  The scores are fp32 dot products over the im2col
rows the patch GEMM reads (the gate's Linear folded into the conv weight, `uvc_token_gate_fold`), so the kept-token indices do not depend on
tensor-core rounding.  The patch embedding itself is produced by the same im2col + tcgen05 GEMM the engine uses, through its own small
autograd node so the gate's gradient reaches `patch_embed.proj`; the returned mask is applied inside the engine's token-assembly kernel, and
the embeddings enter the engine as `pe_in` (the engine then skips its own im2col + patch GEMM, and returns d_pe to this node's backward).
"""
import torch
import torch.nn.functional as F

from .. import ops


class _PatchEmbedFn(torch.autograd.Function):
    """patch_embed.proj as im2col + tcgen05 GEMM.  Its weight / bias gradients are written into the model's flat gradient arena (the slots the
    engine would have filled had it run the patch conv itself), so the one-collective DDP exchange and the flat optimiser sweep see them."""

    @staticmethod
    def forward(ctx, model, x, w, b, patch):
        cols = ops.im2col16(x.contiguous(), patch)
        w2 = w.reshape(w.shape[0], -1)
        pe = ops.linear(cols, w2, b)
        ctx.save_for_backward(cols, w2)
        ctx.model, ctx.params = model, (w, b)
        model._tg_cols = cols           # the gate scores are fp32 dot products over these rows (token_gate_mask, same forward)
        return pe.view(x.shape[0], -1, w.shape[0])

    @staticmethod
    def backward(ctx, dpe):
        cols, w2 = ctx.saved_tensors
        dpe = dpe.contiguous().view(-1, w2.shape[0])
        M, Cout, K = dpe.shape[0], w2.shape[0], w2.shape[1]
        es = ctx.model._tables()
        gw, gb = es.grad_views[0], es.grad_views[1]          # patch_w, patch_b slots (zeroed by the engine's backward, which ran just before,
        w, b = ctx.params                                     # unless gradients are being accumulated over micro-steps)
        accumulate = w.grad is not None and w.grad.data_ptr() == gw.data_ptr()
        ops.gemm(ops.operand(dpe, mn_major=True), ops.operand(cols, mn_major=True), gw.view(Cout, K), Cout, K, M, flags=ops.EPI_ATOMIC)
        ops.colsum_(dpe, gb)
        if accumulate:
            return None, None, None, None, None
        off_w, off_b = es.grad_offs[0], es.grad_offs[1]      # fresh view objects: AccumulateGrad adopts a gradient it solely owns
        return (None, None, es.grad_arena[off_w:off_w + w.numel()].view(w.shape), es.grad_arena[off_b:off_b + b.numel()].view(b.shape), None)


class _TokenGateFn(torch.autograd.Function):
    """mask = straight-through Gumbel top-k of Linear(C,1)(pe [* patch gate]); the gradient reaches pe, the patch gate and `gumbel.*`."""

    @staticmethod
    def forward(ctx, feat, v, c1, pe, patch_scale, gate_w, gate_b, noise, tau, k):
        B, np_, _ = pe.shape
        mask, ysoft, ls, _ = ops.token_gate_fwd(feat, v, c1, gate_b, patch_scale, noise, tau, k, B, np_)
        ctx.save_for_backward(ysoft, ls, pe, gate_w, patch_scale)
        ctx.tau, ctx.has_bias = tau, gate_b is not None
        return mask

    @staticmethod
    def backward(ctx, dmask):
        ysoft, ls, pe, gate_w, patch_scale = ctx.saved_tensors
        ds = ops.token_gate_bwd(dmask.contiguous(), ysoft, ls, ctx.tau)
        dpe = torch.zeros_like(pe)
        dwg = torch.zeros(gate_w.numel(), device=pe.device)
        dbg = torch.zeros(1, device=pe.device) if ctx.has_bias else None
        dps = torch.zeros_like(patch_scale) if patch_scale is not None else None
        ops.token_gate_apply_(ds, pe.contiguous(), gate_w.reshape(-1).contiguous(), patch_scale, dpe, dwg, dbg, dps)
        return None, None, None, dpe, dps, dwg.view_as(gate_w), dbg, None, None, None


def gumbel_noise_like(B, np_, device):
    """the draw of the reference's gumbel_softmax (:39-40): -log(Exponential(1)), same generator, same shape"""
    return -torch.empty(B, np_, device=device).exponential_().log()


def token_gate_from_tokens(model, pe, patch_scale, tau, k):
    """mask for token embeddings the caller already holds (T2T front end): scores = pe . gumbel.weight + bias in fp32"""
    B, np_, _ = pe.shape
    noise = gumbel_noise_like(B, np_, pe.device)
    pe = pe.contiguous()
    return _TokenGateFn.apply(pe.detach(), model.gumbel.weight.detach().reshape(-1).contiguous(), None, pe, patch_scale, model.gumbel.weight, model.gumbel.bias,
                              noise, float(tau), int(k))


def token_gate_mask(model, x, patch_scale, tau, k):
    """Returns (pe, token_mask): the raw patch embeddings are handed on to the engine as `pe_in`, so the patch conv runs once per forward
    (and its weight gradient once per backward) although both the scores and the token stream consume it."""
    pw, pb = model.patch_embed.proj.weight, model.patch_embed.proj.bias
    pe = _PatchEmbedFn.apply(model, x, pw, pb, int(model.patch_embed.patch_size[0]))
    cols, model._tg_cols = model._tg_cols, None
    B, np_, _ = pe.shape
    v, c1 = ops.token_gate_fold(pw.detach(), pb.detach(), model.gumbel.weight.detach().reshape(-1).contiguous())
    noise = gumbel_noise_like(B, np_, pe.device)
    token_mask = _TokenGateFn.apply(cols, v, c1, pe, patch_scale, model.gumbel.weight, model.gumbel.bias, noise, float(tau), int(k))
    return pe, token_mask
