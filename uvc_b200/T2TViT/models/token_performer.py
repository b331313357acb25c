"""Token_performer — the linear-attention layer of the tokens-to-token front end
(reference UVC/T2TViT/models/token_performer.py:8-69): same constructor, attribute names and state-dict keys.

Scope note (SURVEY.md §8 row a-T / §8f row 2): the front end is ~6 % of T2T-ViT-14's FLOPs and is the NEXT row to move
behind the C ABI; in this round it is plain torch device ops feeding the engine's `pe_in`.  The 14 backbone blocks —
where the time goes — run through uvc_vit_forward / uvc_vit_backward.
"""
import math

import torch
import torch.nn as nn


class Token_performer(nn.Module):
    def __init__(self, dim, in_dim, head_cnt=1, kernel_ratio=0.5, dp1=0.1, dp2=0.1):
        super().__init__()
        self.emb = in_dim * head_cnt
        self.kqv = nn.Linear(dim, 3 * self.emb)
        self.dp = nn.Dropout(dp1)
        self.proj = nn.Linear(self.emb, self.emb)
        self.head_cnt = head_cnt
        self.norm1 = nn.LayerNorm(dim)
        self.norm2 = nn.LayerNorm(self.emb)
        self.epsilon = 1e-8
        self.mlp = nn.Sequential(nn.Linear(self.emb, self.emb), nn.GELU(), nn.Linear(self.emb, self.emb), nn.Dropout(dp2))
        self.m = int(self.emb * kernel_ratio)
        self.w = nn.Parameter(nn.init.orthogonal_(torch.randn(self.m, self.emb)) * math.sqrt(self.m), requires_grad=False)

    def prm_exp(self, x):
        """Positive random features of the softmax kernel: exp(w^T x - |x|^2 / 2) / sqrt(m)   (:31-43)."""
        xd = (x * x).sum(dim=-1, keepdim=True) / 2
        return torch.exp(x.float() @ self.w.t() - xd) / math.sqrt(self.m)

    def single_attn(self, x):
        k, q, v = torch.split(self.kqv(x), self.emb, dim=-1)
        kp, qp = self.prm_exp(k), self.prm_exp(q)                       # [B, T, m]
        D = (qp @ kp.sum(dim=1).unsqueeze(-1))                           # [B, T, 1]
        kptv = v.float().transpose(1, 2) @ kp                            # [B, emb, m]
        y = (qp @ kptv.transpose(1, 2)) / (D + self.epsilon)             # [B, T, emb]
        y = v + self.dp(self.proj(y))                                    # v is the skip connection (:51)
        B, T, dim = x.shape
        macs = B * (T * dim * 3 * self.emb + 2 * (T * self.emb + self.emb * T * self.emb) + T * self.m + T * self.emb * self.m
                    + T * self.m * self.emb + T * self.emb * self.emb)   # :54-63
        return y, macs

    def forward(self, x):
        x, macs = self.single_attn(self.norm1(x))
        x = x + self.mlp(self.norm2(x))
        return x, macs + x.shape[0] * (x.shape[1] * x.shape[2] * self.emb + x.shape[2] * self.emb * self.emb)   # :67
