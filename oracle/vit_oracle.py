"""ORACLE (test infrastructure only) — plain-PyTorch CPU restatement of the reference's hot path.

Only `tests/`, `__graft_entry__.smoke()`, `oracle/gen_golden.py` and `bench.py`'s reference / cpu_baseline
legs may import this file; the product (`uvc_b200/`) never does.

Every function restates, op for op, a span of the reference (paths under /root/reference/UVC) and cites
it.  The restatement is PINNED against the unmodified reference modules imported in place
(`oracle/ref_shim.py`): `tests/test_oracle_vs_reference.py` (runs where /root/reference exists) checks
logits, losses, gradients and MAC lists bit-for-bit on CPU, and `tests/golden/*.pt` (written by
`oracle/gen_golden.py` FROM THE REFERENCE) pins it again on the GPU box where the reference is absent.

Everything works on a state dict with the reference's key names, in whatever dtype the tensors have
(fp32 for parity with the reference; the GPU tests also evaluate it in fp64 to measure which of
{reference fp32, CUDA TF32} is closer to exact arithmetic).
"""
import math

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------------------------- model forward
def patch_embed(sd, x, patch=16):
    """models/model_distilled.py:145-153 — Conv2d(k=stride=patch) -> flatten(2).transpose(1,2)."""
    y = F.conv2d(x, sd["patch_embed.proj.weight"], sd["patch_embed.proj.bias"], stride=patch)
    return y.flatten(2).transpose(1, 2)


def attention(sd, pre, x, num_heads):
    """models/model_distilled.py:168-191."""
    B, N, C = x.shape
    qkv = F.linear(x, sd[pre + "attn.qkv.weight"], sd.get(pre + "attn.qkv.bias"))
    qkv = qkv.reshape(B, N, 3, num_heads, C // num_heads).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    attn = (q @ k.transpose(-2, -1)) * ((C // num_heads) ** -0.5)
    attn = attn.softmax(dim=-1)
    x = (attn @ v).transpose(1, 2).reshape(B, N, C)
    return F.linear(x, sd[pre + "attn.proj.weight"], sd[pre + "attn.proj.bias"])


def mlp(sd, pre, x):
    """models/model_distilled.py:112-126 (nn.GELU() = erf form)."""
    x = F.linear(x, sd[pre + "mlp.fc1.weight"], sd[pre + "mlp.fc1.bias"])
    x = F.gelu(x)
    return F.linear(x, sd[pre + "mlp.fc2.weight"], sd[pre + "mlp.fc2.bias"])


def block(sd, i, x, num_heads, eps):
    """models/model_distilled.py:238-244 (enable_part_gating == 0 branch)."""
    pre = f"blocks.{i}."
    C = x.shape[-1]
    x = x + attention(sd, pre, F.layer_norm(x, (C,), sd[pre + "norm1.weight"], sd[pre + "norm1.bias"], eps), num_heads)
    x = x + mlp(sd, pre, F.layer_norm(x, (C,), sd[pre + "norm2.weight"], sd[pre + "norm2.bias"], eps))
    return x


def block_macs(B, N, C, H, Fh):
    """MAC bookkeeping of Attention.forward / Mlp.forward (models/model_distilled.py:115,121,177,182,185,189)."""
    d = C // H
    return [B * 3 * C * N * C, N * B * H * N * d, N * B * H * N * d, B * N * C * C, Fh * B * N * C, C * B * N * Fh]


def token_performer(sd, pre, x):
    """T2TViT/models/token_performer.py:31-69 in eval mode (dropouts are identities): LayerNorm -> kqv Linear ->
    positive random features exp(w^T x - |x|^2/2)/sqrt(m) for k and q -> linear attention normalised by D -> proj
    with v as the skip connection -> LayerNorm + 2-layer GELU MLP residual."""
    emb = sd[pre + "proj.weight"].shape[0]
    w = sd[pre + "w"]
    m = w.shape[0]
    x = F.layer_norm(x, (x.shape[-1],), sd[pre + "norm1.weight"], sd[pre + "norm1.bias"], 1e-5)
    k, q, v = torch.split(F.linear(x, sd[pre + "kqv.weight"], sd[pre + "kqv.bias"]), emb, dim=-1)

    def prm_exp(t):   # :31-43
        td = (t * t).sum(dim=-1, keepdim=True).repeat(1, 1, m) / 2
        return torch.exp(torch.einsum('bti,mi->btm', t, w) - td) / math.sqrt(m)

    kp, qp = prm_exp(k), prm_exp(q)
    D = torch.einsum('bti,bi->bt', qp, kp.sum(dim=1)).unsqueeze(dim=2)
    kptv = torch.einsum('bin,bim->bnm', v, kp)
    y = torch.einsum('bti,bni->btn', qp, kptv) / (D.repeat(1, 1, emb) + 1e-8)
    y = v + F.linear(y, sd[pre + "proj.weight"], sd[pre + "proj.bias"])
    h = F.layer_norm(y, (emb,), sd[pre + "norm2.weight"], sd[pre + "norm2.bias"], 1e-5)
    h = F.linear(F.gelu(F.linear(h, sd[pre + "mlp.0.weight"], sd[pre + "mlp.0.bias"])), sd[pre + "mlp.2.weight"], sd[pre + "mlp.2.bias"])
    return y + h


def token_performer_macs(B, T, dim, emb, m):
    """MAC bookkeeping of Token_performer (token_performer.py:54-68), B included as the reference does."""
    attn = B * (T * dim * 3 * emb + 2 * (T * emb + emb * T * emb) + T * m + T * emb * m + T * m * emb + T * emb * emb)
    return attn + B * (T * emb * emb + emb * emb * emb)


def t2t_tokens(sd, x):
    """T2T_module.forward, tokens_type='performer' (T2TViT/models/t2t_vit.py:83-105): Unfold 7x7/4 -> performer ->
    Unfold 3x3/2 -> performer -> Unfold 3x3/2 -> project.  Returns ([B, 196, C] tokens, macs) for 224x224 inputs."""
    pre = "tokens_to_token."
    B = x.shape[0]
    macs = 0
    t = F.unfold(x, kernel_size=7, stride=4, padding=2).transpose(1, 2)
    for name in ("attention1.", "attention2."):
        emb, m = sd[pre + name + "proj.weight"].shape[0], sd[pre + name + "w"].shape[0]
        macs += token_performer_macs(B, t.shape[1], t.shape[2], emb, m)
        t = token_performer(sd, pre + name, t)
        side = int(math.isqrt(t.shape[1]))
        t = t.transpose(1, 2).reshape(B, t.shape[2], side, side)
        t = F.unfold(t, kernel_size=3, stride=2, padding=1).transpose(1, 2)
    return F.linear(t, sd[pre + "project.weight"], sd[pre + "project.bias"]), macs


def token_gate(sd, pe, noise, tau, k, patch_scale=None):
    """Token slimming gate of models/model_distilled.py:446-456 with gumbel_softmax :36-63 / scatter :21-33, the Gumbel draw supplied:
    scores = Linear(C,1)(pe) -> log_softmax -> (+ noise) / tau -> softmax -> one-hot of the top k -> straight-through value; column 0 forced to 1.
    Returns (mask [B, np] as the reference multiplies it in, y_soft)."""
    x = pe if patch_scale is None else pe * patch_scale.view(1, -1, 1)
    B = x.shape[0]
    scores = F.linear(x, sd["gumbel.weight"], sd["gumbel.bias"]).reshape(B, -1)
    y_soft = ((F.log_softmax(scores, dim=-1) + noise) / tau).softmax(-1)
    index = y_soft.topk(k, dim=-1)[1]
    y_hard = torch.zeros_like(y_soft).scatter_(1, index, 1.0)
    mask = y_hard - y_soft.detach() + y_soft
    mask = mask.clone()
    mask[:, 0] = 1.0
    return mask, y_soft


def forward(sd, x, depth, num_heads, eps=1e-6, blend=None, skip=None, patch_scale=None, token_mask=None, enable_jumping=False, patch=16,
            tokens=None):
    """DistilledVisionTransformer.forward_features + forward, enable_dist == 0 (models/model_distilled.py:429-531).

    With `tokens` ([B, np, C], e.g. from t2t_tokens) the same body is T2T_ViT.forward_features + forward
    (T2TViT/models/t2t_vit.py:168-208; Block = transformer_block.py:42-112, the same arithmetic as block() with
    eps 1e-5 and no qkv bias); `blend` rows are then the softmax / Gumbel `distrib` of :183-187.

    blend: [L,2] tensor of (d0, d1) = the `distrib` of :480-493 (already sampled), or None
    skip:  list[bool] of hard-skipped blocks (:496-500), or None
    patch_scale: [196] multiplier (:434-444); token_mask: [B,196] multiplier (:446-456)
    returns logits [B, num_classes]
    """
    x = patch_embed(sd, x, patch) if tokens is None else tokens
    if patch_scale is not None:
        x = x * patch_scale.view(1, -1, 1)
    if token_mask is not None:
        x = x * token_mask.unsqueeze(-1)
    B = x.shape[0]
    x = torch.cat((sd["cls_token"].expand(B, -1, -1), x), dim=1)
    x = x + sd["pos_embed"]
    accum = 0
    for i in range(depth):
        if blend is not None:
            tmp = block(sd, i, x, num_heads, eps)
            x = blend[i, 1] * tmp + blend[i, 0] * x
        elif skip is None or not skip[i]:
            x = block(sd, i, x, num_heads, eps)
        accum = accum + x
    if enable_jumping:
        x = accum
    x = F.layer_norm(x, (x.shape[-1],), sd["norm.weight"], sd["norm.bias"], eps)
    return F.linear(x[:, 0], sd["head.weight"], sd["head.bias"])


# --------------------------------------------------------------------------------------------- losses
def soft_target_cross_entropy(logits, target):
    """timm.loss.SoftTargetCrossEntropy (un-vendored dependency; joint_train.py:940-942):
    mean_b sum_c -target * log_softmax(logits)."""
    return torch.sum(-target * F.log_softmax(logits, dim=-1), dim=-1).mean()


def distillation_loss(logits, teacher_logits, target, alpha, T, distillation_type="soft"):
    """utils/losses.py:25-65 with outputs_kd is outputs (enable_deit == 0, models/model_distilled.py:523-524).
    returns (loss, base, kd)."""
    base = soft_target_cross_entropy(logits, target)
    if distillation_type == "none":
        return base, base, torch.zeros_like(base)
    kd = F.kl_div(F.log_softmax(logits / T, dim=1), F.log_softmax(teacher_logits / T, dim=1), reduction="sum", log_target=True) \
        * (T * T) / logits.numel()
    return base * (1 - alpha) + kd * alpha, base, kd


# --------------------------------------------------------------------------------------------- mixup (timm.data.Mixup, batch mode)
def one_hot_smooth(y, num_classes, smoothing):
    """timm.data.mixup.one_hot / mixup_target: off = smoothing / K, on = 1 - smoothing + off."""
    off = smoothing / num_classes
    on = 1.0 - smoothing + off
    return torch.full((y.shape[0], num_classes), off, dtype=torch.float32, device=y.device).scatter_(1, y.view(-1, 1), on)


def mixup_target(y, num_classes, lam, smoothing):
    y1 = one_hot_smooth(y, num_classes, smoothing)
    y2 = one_hot_smooth(y.flip(0), num_classes, smoothing)
    return y1 * lam + y2 * (1.0 - lam)


# --------------------------------------------------------------------------------------------- optimiser
def clip_adamw_step(params, grads, ms, vs, step, lr, max_norm=1.0, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.05):
    """torch.nn.utils.clip_grad_norm_(., max_norm) + torch.optim.AdamW.step() (joint_train.py:271,428-429), single-tensor form."""
    total = torch.sqrt(sum((g.double() ** 2).sum() for g in grads)).float()
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    b1, b2 = betas
    bc1 = 1 - b1 ** step
    bc2 = 1 - b2 ** step
    for p, g, m, v in zip(params, grads, ms, vs):
        g = g * coef
        p.mul_(1 - lr * weight_decay)
        m.lerp_(g, 1 - b1)
        v.mul_(b2).addcmul_(g, g, value=1 - b2)
        denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
        p.addcdiv_(m, denom, value=-lr / bc1)
    return total
