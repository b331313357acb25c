"""Token (patch) slimming gate, mode 2 of `--enable_patch_gating` (reference models/model_distilled.py:446-456):

    scores = Linear(C,1)(patch_embed(x) [* patch gate])            -> [B, 196]
    mask   = straight-through top-k of softmax((log_softmax(scores) + Gumbel) / tau)
    mask[:, 0] = 1                                                  (hits patch 0, as in the reference)

The [B,196] score / top-k arithmetic stays in torch (it is B*196 numbers and it must consume the
reference's random stream); the patch embedding it reads is produced by the same im2col + tcgen05
GEMM the engine uses, through its own small autograd node so the scores' gradient reaches
`patch_embed.proj` and `gumbel`.  The returned mask is applied inside the engine's token-assembly kernel, and the embeddings themselves
enter the engine as `pe_in` (the engine then skips its own im2col + patch GEMM, and returns d_pe to this node's backward).
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


def token_gate_mask(model, x, patch_scale, tau, k):
    """Returns (pe, token_mask): the raw patch embeddings are handed on to the engine as `pe_in`, so the patch conv runs once per forward
    (and its weight gradient once per backward) although both the scores and the token stream consume it."""
    from .model_distilled import gumbel_softmax
    pe = _PatchEmbedFn.apply(model, x, model.patch_embed.proj.weight, model.patch_embed.proj.bias, int(model.patch_embed.patch_size[0]))
    scored = pe if patch_scale is None else pe * patch_scale.view(1, -1, 1)
    B = pe.shape[0]
    token_scores = F.linear(scored, model.gumbel.weight, model.gumbel.bias).reshape(B, -1)
    token_mask = gumbel_softmax(F.log_softmax(token_scores, dim=-1), k=k, tau=tau, hard=True)
    token_mask = token_mask.clone()
    token_mask[:, 0] = 1.
    return pe, token_mask.contiguous()
