"""T2T-ViT with UVC block gates, backbone on the sm_100a engine (reference UVC/T2TViT/models/t2t_vit.py:46-250).

`T2T_ViT` keeps the reference's constructor arguments, attribute names and state-dict keys (`tokens_to_token.*`, `cls_token`,
the fixed sinusoid `pos_embed`, `blocks.N.*` without a qkv bias, `norm`, `head`, `block_skip_gating`), so a reference T2T checkpoint
loads unchanged.  The 14 `Block`s (transformer_block.py:42-112 — the DeiT block arithmetic with LayerNorm eps 1e-5, no qkv bias,
mlp_ratio 3), the token assembly, the final norm and the head run through `uvc_vit_forward` / `uvc_vit_backward`: the front end's
[B, 196, C] tokens enter as `pe_in`, and their gradient comes back as `d_pe` and continues through torch autograd into
`tokens_to_token`.

Deviations from the reference, which cannot execute T2T + UVC at HEAD (SURVEY.md §5 / §8 row a-T):
  * `forward` accepts the `(x, tau, number)` call joint_train.py:410,1012 makes (the reference's takes `x` only and raises);
  * the block-gate branch works (`F` is not imported in the reference file; `self.gumbel_hard` is never set there — here it is);
  * `enable_patch_gating` 1 / 2 apply the DeiT gate logic (models/model_distilled.py:434-456) to the tokens_to_token output.  The flags are
    set on the model AFTER construction (uvc_optimizer.py:125-129), so the token scorer `gumbel` = Linear(C, 1) always exists: the state dict
    is the reference's plus `gumbel.weight` / `gumbel.bias` (reference checkpoints load with strict=False, as joint_train.py:148 does).
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from ..._lib import VitDims, operand_f16_for
from ...models.model_distilled import DistilledVisionTransformer, _VitFunction, _engine_param_list, gumbel_softmax, hard_skip_list
from .token_performer import Token_performer


def get_sinusoid_encoding(n_position, d_hid):
    """Fixed sinusoid table [1, n_position, d_hid] (transformer_block.py:115-125): float64 angles pos / 10000^(2*(j//2)/d),
    sin on even columns, cos on odd ones, stored as float32."""
    pos = torch.arange(n_position, dtype=torch.float64).unsqueeze(1)
    j = torch.arange(d_hid, dtype=torch.float64).unsqueeze(0)
    ang = pos / torch.pow(torch.tensor(10000.0, dtype=torch.float64), 2 * torch.div(j, 2, rounding_mode="floor") / d_hid)
    even = (torch.arange(d_hid) % 2 == 0).unsqueeze(0)
    return torch.where(even, torch.sin(ang), torch.cos(ang)).float().unsqueeze(0)


class _T2TFrontFn(torch.autograd.Function):
    """One autograd node for the whole front end: forward = uvc_t2t_forward, backward = uvc_t2t_backward."""

    @staticmethod
    def forward(ctx, mod, x, *params):
        need_grad = any(ctx.needs_input_grad) and getattr(mod, "_outer_grad_enabled", True)      # see _VitFunction.forward
        tokens, p, seed = mod._engine_forward(x, save=need_grad)
        ctx.mod, ctx.p, ctx.seed = mod, p, seed
        ctx.save_for_backward(x)
        ctx.requires = [q.requires_grad for q in params]
        return tokens

    @staticmethod
    def backward(ctx, d_tokens):
        (x,) = ctx.saved_tensors
        grads = ctx.mod._engine_backward(x, d_tokens.contiguous().float(), ctx.p, ctx.seed)
        return (None, None, *[g if r else None for g, r in zip(grads, ctx.requires)])


class T2T_module(nn.Module):
    """Tokens-to-token front end, tokens_type='performer' (t2t_vit.py:46-105): three soft splits (Unfold 7x7/4, 3x3/2, 3x3/2) with a
    Token_performer after the first two, then a Linear to the embedding width -- on the device ONE call per pass into libuvc_sm100.so
    (uvc_t2t_forward / uvc_t2t_backward, csrc/t2t_frontend.cu); the sub-modules hold the parameters under the reference's names."""

    def __init__(self, img_size=224, tokens_type='performer', in_chans=3, embed_dim=768, token_dim=64):
        super().__init__()
        if tokens_type != 'performer':
            raise NotImplementedError(f"tokens_type={tokens_type!r}: UVC builds t2t_vit_14, which uses 'performer' (t2t_vit.py:248)")
        if token_dim != 64:
            raise NotImplementedError("token_dim: the sm_100a front end implements T2T_module's default of 64")
        self.soft_split0 = nn.Unfold(kernel_size=(7, 7), stride=(4, 4), padding=(2, 2))
        self.soft_split1 = nn.Unfold(kernel_size=(3, 3), stride=(2, 2), padding=(1, 1))
        self.soft_split2 = nn.Unfold(kernel_size=(3, 3), stride=(2, 2), padding=(1, 1))
        self.attention1 = Token_performer(dim=in_chans * 7 * 7, in_dim=token_dim, kernel_ratio=0.5)
        self.attention2 = Token_performer(dim=token_dim * 3 * 3, in_dim=token_dim, kernel_ratio=0.5)
        self.project = nn.Linear(token_dim * 3 * 3, embed_dim)
        self.num_patches = (img_size // 16) * (img_size // 16)
        self._ws = {}

    # ------------------------------------------------------------------ engine plumbing
    def _params(self):
        return self.attention1.engine_tensors() + self.attention2.engine_tensors() + [self.project.weight, self.project.bias]

    @staticmethod
    def _fill(tensors):
        from ..._lib import PerformerTensors, T2TTensors
        t = T2TTensors()
        for blk, chunk in ((t.attn1, tensors[:13]), (t.attn2, tensors[13:26])):
            for name, v in zip(PerformerTensors.NAMES, chunk):
                setattr(blk, name, None if v is None else v.data_ptr())
        t.project_w, t.project_b = tensors[26].data_ptr(), tensors[27].data_ptr()
        return t

    def _dims(self, x):
        from ..._lib import T2TDims
        d = T2TDims()
        d.B, d.in_chans, d.img, d.C = int(x.shape[0]), int(x.shape[1]), int(x.shape[2]), int(self.project.out_features)
        d.ln_eps = float(self.attention1.norm1.eps)
        return d

    def _workspace(self, dims, save):
        import ctypes as C_
        from ... import _lib
        key = (dims.B, dims.img, bool(save))
        ws = self._ws.get(key)
        dev = self.project.weight.device
        if ws is None or ws.device != dev:
            n = int(_lib.load().uvc_t2t_workspace_bytes(C_.byref(dims), 1 if save else 0))
            if n == 0:
                raise _lib.UvcError("uvc_t2t_workspace_bytes: " + _lib.load().uvc_last_error().decode())
            ws = torch.empty(n, dtype=torch.uint8, device=dev)
            self._ws = {k: v for k, v in self._ws.items() if k[2] != key[2]}
            self._ws[key] = ws
        return ws

    def _check(self, x, tensors):
        from ..._lib import UvcError
        if not x.is_cuda:
            raise UvcError("uvc_b200 runs on CUDA tensors only (no CPU fallback): move the model and the batch to a B200")
        if x.shape[2] != x.shape[3]:
            raise UvcError("the tokens-to-token front end takes square images")
        for t in tensors:
            if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
                raise UvcError("tokens_to_token parameters must be contiguous CUDA fp32 tensors")

    def _engine_forward(self, x, save):
        import ctypes as C_
        from ... import _lib
        from ..._lib import T2TForwardArgs
        params = [p.detach() for p in self._params()]
        self._check(x, params)
        a = T2TForwardArgs()
        a.dims = self._dims(x)
        a.w = self._fill(params)
        a.x = x.data_ptr()
        tokens = torch.empty(x.shape[0], self.num_patches, a.dims.C, device=x.device, dtype=torch.float32)
        a.tokens = tokens.data_ptr()
        a.save_for_backward = 1 if save else 0
        p = float(self.attention1.dp.p) if self.training else 0.0        # the reference's nn.Dropout(0.1) sites are active in train mode only
        seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if p > 0 else 0   # CPU generator: reproducible under torch.manual_seed, no device sync
        a.dropout_p, a.seed = p, seed
        ws = self._workspace(a.dims, save)
        a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel()
        _lib.check(_lib.load().uvc_t2t_forward(C_.byref(a), C_.c_void_p(torch.cuda.current_stream().cuda_stream)), "uvc_t2t_forward")
        return tokens, p, seed

    def _engine_backward(self, x, d_tokens, p, seed):
        import ctypes as C_
        from ... import _lib
        from ..._lib import T2TBackwardArgs
        params = [q.detach() for q in self._params()]
        grads = [torch.zeros_like(q) for q in params]
        a = T2TBackwardArgs()
        a.dims = self._dims(x)
        a.w, a.g = self._fill(params), self._fill(grads)
        a.x, a.d_tokens = x.data_ptr(), d_tokens.data_ptr()
        a.dropout_p, a.seed = p, seed
        a.grad_scale = float(getattr(self, "grad_scale", 0.0))
        ws = self._workspace(a.dims, True)
        a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel()
        _lib.check(_lib.load().uvc_t2t_backward(C_.byref(a), C_.c_void_p(torch.cuda.current_stream().cuda_stream)), "uvc_t2t_backward")
        return grads

    def forward(self, x):
        B = x.shape[0]
        x = x.contiguous().float()
        self._outer_grad_enabled = torch.is_grad_enabled()
        try:
            tokens = _T2TFrontFn.apply(self, x, *self._params())
        finally:
            self._outer_grad_enabled = True
        side = x.shape[2] // 4
        macs = self.attention1.macs(B, side * side) + self.attention2.macs(B, (side // 2) ** 2)
        return tokens, macs          # the reference does not count project's MACs either (:105)


class T2T_ViT(DistilledVisionTransformer):
    def __init__(self, img_size=224, tokens_type='performer', in_chans=3, num_classes=1000, embed_dim=768, depth=12, num_heads=12,
                 mlp_ratio=4., qkv_bias=False, qk_scale=None, drop_rate=0., attn_drop_rate=0., drop_path_rate=0., norm_layer=nn.LayerNorm,
                 token_dim=64, enable_block_gating=False, enable_jumping=False, enable_patch_gating=0, gumbel_hard=True, use_gumbel=False):
        if qk_scale is not None:
            raise NotImplementedError("qk_scale: only the default head_dim ** -0.5 is on the sm_100a path (t2t_vit_14() leaves it unset)")
        super().__init__(enable_dist=0, enable_jumping=enable_jumping, enable_block_gating=enable_block_gating, enable_patch_gating=enable_patch_gating,
                         gumbel_hard=gumbel_hard, use_gumbel=use_gumbel, img_size=img_size, patch_size=16, in_chans=in_chans,
                         num_classes=num_classes, embed_dim=embed_dim, depth=depth, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias,
                         drop_rate=drop_rate, attn_drop_rate=attn_drop_rate, drop_path_rate=drop_path_rate, norm_layer=norm_layer)
        # the DeiT patch conv is replaced by the tokens-to-token module; keep only its geometry (196 tokens on a 14x14 grid)
        self._img, self._in_chans = int(img_size), int(in_chans)
        del self.patch_embed
        self.tokens_to_token = T2T_module(img_size=img_size, tokens_type=tokens_type, in_chans=in_chans, embed_dim=embed_dim, token_dim=token_dim)
        self.num_patches = self.tokens_to_token.num_patches
        self.pos_embed = nn.Parameter(get_sinusoid_encoding(self.num_patches + 1, embed_dim), requires_grad=False)
        self.tokens_to_token.apply(self._init_weights)

    def _init_weights(self, m):      # t2t_vit.py:162-169
        if isinstance(m, nn.Linear):
            nn.init.trunc_normal_(m.weight, std=.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    @torch.jit.ignore
    def no_weight_decay(self):
        return {'cls_token'}

    def get_classifier(self):
        return self.head

    # ------------------------------------------------------------------ engine plumbing
    def _dims(self, B):
        d = VitDims()
        d.B, d.img, d.patch, d.in_chans = int(B), self._img, 16, self._in_chans     # geometry only: 196 + 1 tokens
        d.C, d.H, d.Fh, d.L = self.embed_dim, self.blocks[0].attn.num_heads, self.blocks[0].mlp.fc1.out_features, len(self.blocks)
        d.num_classes = self.num_classes
        d.ln_eps = float(self.norm.eps)
        d.operand_f16 = 1 if operand_f16_for(self, d.C // d.H, self.num_patches + 1, d.C, d.Fh) else 0
        return d

    def _macs_backbone(self, B, executed):
        C_, N = self.embed_dim, self.num_patches + 1
        H = self.blocks[0].attn.num_heads
        d = C_ // H
        Fh = self.blocks[0].mlp.fc1.out_features
        per_block = [B * 3 * C_ * N * C_, N * B * H * N * d, N * B * H * N * d, B * N * C_ * C_, Fh * B * N * C_, C_ * B * N * Fh]   # transformer_block.py:30,36,62,67,70,74
        return [list(per_block) if e else [] for e in executed]

    def _block_gates(self):
        """t2t_vit.py:181-194: with block gating the blend is a (Gumbel-)softmax over each gate row; otherwise hard skipping."""
        if self.enable_block_gating:
            if self.use_gumbel:
                return F.gumbel_softmax(self.block_skip_gating, tau=0.5, hard=self.gumbel_hard, eps=1e-10, dim=-1).contiguous(), None
            return F.softmax(self.block_skip_gating, dim=-1).contiguous(), None
        return None, hard_skip_list(self)

    def forward_logits(self, x, tau=-1, ratio=0.9):
        B = x.shape[0]
        if not x.is_cuda:
            from ..._lib import UvcError
            raise UvcError("uvc_b200 runs on CUDA tensors only (no CPU fallback): move the model and the batch to a B200")
        pe, macs_embed = self.tokens_to_token(x.float())
        np_ = self.num_patches
        patch_scale = token_mask = None
        if self.enable_patch_gating == 1:
            pg = torch.sigmoid(self.patch_gating).reshape(np_)
            if self.patch_hard:
                pg = (pg.detach() >= 0.5).float()
                pg[0] = 1
            patch_scale = pg.contiguous()
        if tau > 0 and self.enable_patch_gating == 2:
            from ...models.token_gate import token_gate_from_tokens
            token_mask = token_gate_from_tokens(self, pe, patch_scale, tau, int(ratio * np_))
        blend, skip = self._block_gates()
        params = [p for _, p in _engine_param_list(self)]
        self._outer_grad_enabled = torch.is_grad_enabled()
        try:
            logits = _VitFunction.apply(self, pe.contiguous(), blend, patch_scale, token_mask, skip, *params)
        finally:
            self._outer_grad_enabled = True
        executed = [True] * len(self.blocks) if skip is None else [not s for s in skip]
        return logits, (macs_embed, self._macs_backbone(B, executed))

    def forward(self, x, tau=-1, number=0.9):
        x, macs_list = self.forward_logits(x, tau, number)
        if self.training:
            return (x, x), macs_list
        return x, macs_list


def t2t_vit_14(pretrained=False, **kwargs):
    """t2t_vit.py:244-250."""
    if pretrained:
        raise NotImplementedError("pretrained=True needs the upstream checkpoint download; load a local checkpoint with load_state_dict")
    return T2T_ViT(tokens_type='performer', embed_dim=384, depth=14, num_heads=6, mlp_ratio=3., **kwargs)
