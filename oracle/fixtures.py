"""ORACLE (test infrastructure only) — deterministic synthetic weights / batches shared by the golden
generator (which feeds them to the UNMODIFIED reference) and by the tests (which feed them to the CUDA path).

Nothing here depends on a model constructor's RNG order: every tensor is drawn from its own
`torch.Generator` seeded by (seed, key), on CPU, so the same bytes are produced here and on the GPU box.
"""
import hashlib

import torch

MODEL_DIMS = {  # models/configs.py:34-53,112-165 of the reference
    "deit_tiny_patch16_224": dict(embed_dim=192, depth=12, num_heads=3),
    "deit_small_patch16_224": dict(embed_dim=384, depth=12, num_heads=6),
    "deit_base_patch16_224": dict(embed_dim=768, depth=12, num_heads=12),
    "t2t_vit_14": dict(embed_dim=384, depth=14, num_heads=6, mlp_ratio=3, t2t=True),   # T2TViT/models/t2t_vit.py:244-250
}


def _gen(seed, key):
    h = int.from_bytes(hashlib.sha256(f"{seed}:{key}".encode()).digest()[:7], "little")
    g = torch.Generator(device="cpu")
    g.manual_seed(h)
    return g


def param_shapes(embed_dim, depth, num_heads, mlp_ratio=4, num_classes=1000, patch=16, in_chans=3, img=224, t2t=False, token_dim=64):
    C, Fh = embed_dim, int(embed_dim * mlp_ratio)
    n = (img // patch) ** 2 + 1
    shapes = {"cls_token": (1, 1, C), "pos_embed": (1, n, C), "norm.weight": (C,), "norm.bias": (C,), "head.weight": (num_classes, C),
              "head.bias": (num_classes,), "block_skip_gating": (depth, 2)}
    if t2t:   # T2TViT/models/t2t_vit.py:63-73, token_performer.py:9-29
        for name, dim in (("attention1.", in_chans * 49), ("attention2.", token_dim * 9)):
            p = "tokens_to_token." + name
            e = token_dim
            shapes.update({p + "w": (e // 2, e), p + "kqv.weight": (3 * e, dim), p + "kqv.bias": (3 * e,), p + "proj.weight": (e, e),
                           p + "proj.bias": (e,), p + "norm1.weight": (dim,), p + "norm1.bias": (dim,), p + "norm2.weight": (e,),
                           p + "norm2.bias": (e,), p + "mlp.0.weight": (e, e), p + "mlp.0.bias": (e,), p + "mlp.2.weight": (e, e),
                           p + "mlp.2.bias": (e,)})
        shapes.update({"tokens_to_token.project.weight": (C, token_dim * 9), "tokens_to_token.project.bias": (C,)})
    else:
        shapes.update({"patch_embed.proj.weight": (C, in_chans, patch, patch), "patch_embed.proj.bias": (C,),
                       "gumbel.weight": (1, C), "gumbel.bias": (1,)})
    for i in range(depth):
        p = f"blocks.{i}."
        if not t2t:
            shapes[p + "attn.qkv.bias"] = (3 * C,)
        shapes.update({p + "norm1.weight": (C,), p + "norm1.bias": (C,), p + "attn.qkv.weight": (3 * C, C),
                       p + "attn.proj.weight": (C, C), p + "attn.proj.bias": (C,), p + "norm2.weight": (C,), p + "norm2.bias": (C,),
                       p + "mlp.fc1.weight": (Fh, C), p + "mlp.fc1.bias": (Fh,), p + "mlp.fc2.weight": (C, Fh), p + "mlp.fc2.bias": (C,),
                       p + "attn_skip_gating": (2,), p + "mlp_skip_gating": (2,)})
    return shapes


def sinusoid_table(n_position, d_hid):
    """The fixed position table of T2T-ViT (T2TViT/models/transformer_block.py:115-125): angle = pos / 10000^(2*(j//2)/d),
    sin on even columns, cos on odd columns, computed in float64 and stored as float32."""
    pos = torch.arange(n_position, dtype=torch.float64).unsqueeze(1)
    j = torch.arange(d_hid, dtype=torch.float64).unsqueeze(0)
    ang = pos / torch.pow(torch.tensor(10000.0, dtype=torch.float64), 2 * torch.div(j, 2, rounding_mode="floor") / d_hid)
    tab = torch.where((torch.arange(d_hid) % 2 == 0).unsqueeze(0), torch.sin(ang), torch.cos(ang))
    return tab.float().unsqueeze(0)


def make_state_dict(model_type="deit_tiny_patch16_224", depth=None, seed=0, wstd=0.02, num_classes=1000):
    """A "trained-looking" synthetic checkpoint: weights ~ N(0, wstd) (patch conv ~ N(0, 1/sqrt(768))), LayerNorm
    gains near 1, small non-zero biases everywhere (so bias / beta paths are exercised), gates at the reference's
    init [-1, 1] (models/model_distilled.py:416)."""
    dims = dict(MODEL_DIMS[model_type])
    if depth is not None:
        dims["depth"] = depth
    sd = {}
    for k, shp in param_shapes(num_classes=num_classes, **dims).items():
        g = _gen(seed, k)
        if k.endswith("skip_gating"):
            t = torch.tensor([-1.0, 1.0]).expand(shp).contiguous()
        elif "norm" in k and k.endswith("weight"):
            t = 1.0 + 0.1 * torch.randn(shp, generator=g)
        elif k.endswith("bias"):
            t = 0.02 * torch.randn(shp, generator=g)
        elif k == "patch_embed.proj.weight":
            t = torch.randn(shp, generator=g) * (1.0 / 768 ** 0.5)
        elif k == "pos_embed" and dims.get("t2t"):
            t = sinusoid_table(shp[1], shp[2])
        elif k.endswith("attention1.w") or k.endswith("attention2.w"):
            t = torch.randn(shp, generator=g) * 0.7          # entries of an orthogonal [m, emb] matrix times sqrt(m) have this scale
        elif k.startswith("tokens_to_token."):
            t = torch.randn(shp, generator=g) * 0.05
        else:
            t = torch.randn(shp, generator=g) * wstd
        sd[k] = t
    return sd, dims


def make_batch(B, seed=730, num_classes=1000, img=224):
    """images ~ N(0,1) (ImageNet-normalised statistics), integer labels."""
    x = torch.randn(B, 3, img, img, generator=_gen(seed, "x"))
    y = torch.randint(0, num_classes, (B,), generator=_gen(seed, "y"))
    return x, y


def soft_targets(B, seed=730, num_classes=1000, lam=0.7, smoothing=0.1):
    """A mixup-style soft target matrix (rows sum to 1)."""
    from .vit_oracle import mixup_target
    _, y = make_batch(B, seed, num_classes, img=16)
    return mixup_target(y, num_classes, lam, smoothing)


def checksum(t):
    t = t.detach().double().flatten()
    return [float(t.sum()), float((t * t).sum()), float(t[:: max(1, t.numel() // 97)].sum())]
