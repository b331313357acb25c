"""Import the UNMODIFIED reference modules in place, on CPU (TEST INFRASTRUCTURE ONLY).

This file is part of the oracle: only `tests/`, `oracle/gen_golden.py`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline leg may import it.  It never
copies reference source; it only makes `/root/reference/UVC` importable under
torch 2.x without `timm` / `apex` / `ml_collections`:

  * a stub `timm` package with the five symbols `models/model_distilled.py:7-10` pulls in;
  * `Tensor.cuda()` / `Module.cuda()` turned into no-ops, because `.cuda()` is hard-coded at
    `models/model_distilled.py:29,40,480,483` and `uvc_utils.py:162,166,180-212,257,263,446-451`;
  * `get_uvc_layers` re-stated here (`joint_train.py:530-564` cannot be imported: its
    top-level imports need apex / timm.data / ml_collections).

`/root/reference` only exists in the build container, never on the GPU box, so
everything that must travel is written to `tests/golden/` by `oracle/gen_golden.py`.
"""
import os
import sys
import types

import torch
import torch.nn as nn

REF_ROOT = os.environ.get("UVC_REFERENCE_ROOT", "/root/reference")
REF_UVC = os.path.join(REF_ROOT, "UVC")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_UVC, "uvc_utils.py"))


def _install_timm_stub():
    if "timm" in sys.modules and not getattr(sys.modules["timm"], "_uvc_stub", False):
        return  # a real timm is present; use it
    def _cfg(url="", **kw):
        return {"url": url, **kw}

    def register_model(fn):
        return fn

    class DropPath(nn.Module):
        def __init__(self, p=0.0):
            super().__init__()
            assert p == 0.0, "oracle shim only supports drop_path == 0 (every shipped config)"

        def forward(self, x):
            return x

    def to_2tuple(x):
        return tuple(x) if isinstance(x, (tuple, list)) else (x, x)

    timm = types.ModuleType("timm"); timm._uvc_stub = True
    models = types.ModuleType("timm.models")
    vt = types.ModuleType("timm.models.vision_transformer"); vt._cfg = _cfg
    reg = types.ModuleType("timm.models.registry"); reg.register_model = register_model
    layers = types.ModuleType("timm.models.layers")
    layers.trunc_normal_ = nn.init.trunc_normal_
    layers.DropPath = DropPath
    layers.to_2tuple = to_2tuple
    helpers = types.ModuleType("timm.models.layers.helpers"); helpers.to_2tuple = to_2tuple
    mhelpers = types.ModuleType("timm.models.helpers"); mhelpers.load_pretrained = lambda *a, **k: None
    timm.models = models
    models.vision_transformer, models.registry, models.layers, models.helpers = vt, reg, layers, mhelpers
    layers.helpers = helpers
    for name, mod in [("timm", timm), ("timm.models", models), ("timm.models.vision_transformer", vt),
                      ("timm.models.registry", reg), ("timm.models.layers", layers),
                      ("timm.models.layers.helpers", helpers), ("timm.models.helpers", mhelpers)]:
        sys.modules[name] = mod


_loaded = {}


def load():
    """Return a namespace with the reference's own modules (imported from REF_UVC)."""
    if _loaded:
        return _loaded["ns"]
    if not available():
        raise RuntimeError(f"reference checkout not found under {REF_UVC}")
    _install_timm_stub()
    # .cuda() -> no-op so the reference's hard-coded device moves stay on CPU
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
        nn.Module.cuda = lambda self, *a, **k: self
    if REF_UVC not in sys.path:
        sys.path.insert(0, REF_UVC)
    import importlib
    ns = types.SimpleNamespace()
    ns.model_distilled = importlib.import_module("models.model_distilled")
    ns.uvc_utils = importlib.import_module("uvc_utils")
    ns.uvc_optimizer = importlib.import_module("uvc_optimizer")
    ns.losses = importlib.import_module("utils.losses")
    ns.scheduler = importlib.import_module("utils.scheduler")
    ns.get_uvc_layers = get_uvc_layers
    _loaded["ns"] = ns
    return ns


def get_uvc_layers(model):
    """Re-statement of joint_train.py:530-564 (module scan by name)."""
    layer_names = {None: None}
    uvc_layers = {"W1": [], "W2": [], "W3": []}
    for name, m in model.named_modules():
        if not hasattr(m, "in_features"):
            continue
        if "attn.proj" in name:
            layer_names[m] = name; uvc_layers["W1"].append(m); m.uvc_s = 0
        elif "mlp.fc2" in name:
            layer_names[m] = name; uvc_layers["W3"].append(m); m.uvc_s = 0
        if "mlp.fc1" in name:
            layer_names[m] = name; uvc_layers["W2"].append(m); m.uvc_s = 0
    d = {"s_dict": {}, "r_dict": {}}
    for i, m in enumerate(uvc_layers["W1"]):
        d["s_dict"][m] = [i, 0]; d["r_dict"][m] = i
    for i, m in enumerate(uvc_layers["W3"]):
        d["s_dict"][m] = [i, 1]
    return layer_names, uvc_layers, d


MODEL_DIMS = {  # models/configs.py:34-53,112-165
    "deit_tiny_patch16_224": dict(embed_dim=192, depth=12, num_heads=3),
    "deit_small_patch16_224": dict(embed_dim=384, depth=12, num_heads=6),
    "deit_base_patch16_224": dict(embed_dim=768, depth=12, num_heads=12),
}


def make_ref_model(ns, model_type="deit_tiny_patch16_224", depth=None, gumbel_hard=False, **kw):
    """Build the reference model exactly as joint_train.py:135-140 does."""
    from functools import partial
    dims = dict(MODEL_DIMS[model_type])
    if depth is not None:
        dims["depth"] = depth
    m = ns.model_distilled.DistilledVisionTransformer(
        enable_dist=0, patch_size=16, mlp_ratio=4, qkv_bias=True,
        norm_layer=partial(nn.LayerNorm, eps=1e-6), drop_rate=0, gumbel_hard=gumbel_hard, **dims, **kw)
    return m


def register_masks(model):
    """joint_train.py:169-171."""
    for _, p in model.named_modules():
        if hasattr(p, "weight"):
            p.register_buffer("mask", torch.ones_like(p.weight))


def default_args(**over):
    """argparse defaults of joint_train.py:684-879 that the ADMM path reads, with the
    shipped run_uvc_train.sh values."""
    a = types.SimpleNamespace(
        head_size=64, num_heads=3, flops_with_mhsa=1, use_gumbel=1, enable_block_gating=1,
        enable_part_gating=0, enable_patch_gating=0, enable_jumping=0, eps=0.1, eps_decay=0.92,
        enable_warmup=1, soptim="sgd", roptim="sgd", slr=0.02, rlr=0.02, glr=0.1, ylr=1e-4, plr=1e-4,
        zlr_schedule_list=[1, 5, 9, 13, 17], budget=0.5, sl2wd=0.0, gating_weight=5e-4,
        z_grad_clip=0.5, gating_interval=50)
    for k, v in over.items():
        setattr(a, k, v)
    return a
