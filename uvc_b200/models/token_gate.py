"""Token (patch) slimming gate, mode 2 of `--enable_patch_gating` (reference models/model_distilled.py:446-456):

    scores = Linear(C,1)(patch_embed(x) [* patch gate])            -> [B, 196]
    mask   = straight-through top-k of softmax((log_softmax(scores) + Gumbel) / tau)
    mask[:, 0] = 1                                                  (hits patch 0, as in the reference)

The [B,196] score / top-k arithmetic stays in torch (it is B*196 numbers and it must consume the
reference's random stream); the patch embedding it reads is produced by the same im2col + tcgen05
GEMM the engine uses, through its own small autograd node so the scores' gradient reaches
`patch_embed.proj` and `gumbel`.  The returned mask is applied inside the engine's token-assembly kernel.
"""
import torch
import torch.nn.functional as F

from .. import ops


class _PatchEmbedFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, patch):
        cols = ops.im2col16(x.contiguous(), patch)
        w2 = w.reshape(w.shape[0], -1)
        pe = ops.linear(cols, w2, b)
        ctx.save_for_backward(cols, w2)
        ctx.wshape = w.shape
        return pe.view(x.shape[0], -1, w.shape[0])

    @staticmethod
    def backward(ctx, dpe):
        cols, w2 = ctx.saved_tensors
        dpe = dpe.contiguous().view(-1, w2.shape[0])
        M, Cout, K = dpe.shape[0], w2.shape[0], w2.shape[1]
        dw = torch.zeros_like(w2)
        ops.gemm(ops.operand(dpe, mn_major=True), ops.operand(cols, mn_major=True), dw, Cout, K, M)
        db = torch.zeros(Cout, device=dpe.device)
        ops.colsum_(dpe, db)
        return None, dw.view(ctx.wshape), db, None


def token_gate_mask(model, x, patch_scale, tau, k):
    from .model_distilled import gumbel_softmax
    pe = _PatchEmbedFn.apply(x, model.patch_embed.proj.weight, model.patch_embed.proj.bias, int(model.patch_embed.patch_size[0]))
    if patch_scale is not None:
        pe = pe * patch_scale.view(1, -1, 1)
    B = pe.shape[0]
    token_scores = F.linear(pe, model.gumbel.weight, model.gumbel.bias).reshape(B, -1)
    token_mask = gumbel_softmax(F.log_softmax(token_scores, dim=-1), k=k, tau=tau, hard=True)
    token_mask = token_mask.clone()
    token_mask[:, 0] = 1.
    return token_mask.contiguous()
