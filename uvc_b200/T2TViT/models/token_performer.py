"""Token_performer -- the linear-attention layer of the tokens-to-token front end
(reference UVC/T2TViT/models/token_performer.py:8-69): same constructor, attribute names and state-dict keys.

The module only HOLDS parameters: the whole front end (both soft splits, both Token_performers, the last soft split and the projection)
runs as one `uvc_t2t_forward` / `uvc_t2t_backward` call issued by `T2T_module` (uvc_b200/csrc/t2t_frontend.cu).  `macs()` restates the
reference's own MAC accounting (:54-63, :67).
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
        if self.emb != 64 or self.m != 32 or head_cnt != 1 or dp1 != dp2:
            raise NotImplementedError("the sm_100a front end implements the Token_performer T2T_module builds: emb 64, 32 random features, one head")

    def engine_tensors(self):
        """parameters in the order of `uvc_performer_tensors` (include/uvc_b200.h)"""
        return [self.norm1.weight, self.norm1.bias, self.kqv.weight, self.kqv.bias, self.w, self.proj.weight, self.proj.bias,
                self.norm2.weight, self.norm2.bias, self.mlp[0].weight, self.mlp[0].bias, self.mlp[2].weight, self.mlp[2].bias]

    def macs(self, B, T):
        """token_performer.py:54-63 (single_attn) + :67 (mlp), for an input [B, T, dim]"""
        dim = self.kqv.in_features
        m = B * (T * dim * 3 * self.emb + 2 * (T * self.emb + self.emb * T * self.emb) + T * self.m + T * self.emb * self.m
                 + T * self.m * self.emb + T * self.emb * self.emb)
        return m + B * (T * self.emb * self.emb + self.emb * self.emb * self.emb)

    def forward(self, *a, **k):
        raise RuntimeError("Token_performer is a parameter container: run the enclosing T2T_module (one uvc_t2t_forward call)")
