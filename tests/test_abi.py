"""CPU checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/uvc_b200.h
declares, and the ctypes struct layouts match the C structs.  No compute calls (no GPU here)."""
import ctypes
import os
import re

from uvc_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "uvc_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"UVC_API\s+[\w\s\*]+?\b(uvc_\w+)\s*\(", text)))


def test_header_symbols_are_exported():
    lib = _lib.load()
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/uvc_b200.h but not exported"
    assert sorted(_lib.EXPORTS) == syms, "uvc_b200._lib.EXPORTS is out of sync with the header"


def test_version_and_struct_sizes():
    lib = _lib.load()
    assert lib.uvc_version() == 9
    for name, st in _lib._ABI_STRUCTS.items():
        assert lib.uvc_abi_sizeof(name.encode()) == ctypes.sizeof(st), name
    assert lib.uvc_abi_sizeof(b"no_such_struct") == -1


def test_errors_do_not_cross_the_boundary():
    lib = _lib.load()
    rc = lib.uvc_gemm_tf32(None, None)
    assert rc == -2 and b"NULL" in lib.uvc_last_error()
    d = _lib.VitDims()          # all zero -> bad shape, reported through the return value of the size query
    assert lib.uvc_vit_workspace_bytes(ctypes.byref(d), 1) == 0
    assert lib.uvc_layernorm_fwd(None, 0, None, None, 1e-6, None, 0, None, None, 1, 4, 0, None) == -2


def test_workspace_size_query_runs_on_cpu():
    lib = _lib.load()
    d = _lib.VitDims()
    d.B, d.img, d.patch, d.in_chans, d.C, d.H, d.Fh, d.L, d.num_classes, d.ln_eps = 128, 224, 16, 3, 384, 6, 1536, 12, 1000, 1e-6
    train = lib.uvc_vit_workspace_bytes(ctypes.byref(d), 1)
    infer = lib.uvc_vit_workspace_bytes(ctypes.byref(d), 0)
    assert 0 < infer < train < 40 * 2 ** 30
    d.operand_f16 = 1                      # fp16 operand storage: the saved activations shrink by about a third
    train16 = lib.uvc_vit_workspace_bytes(ctypes.byref(d), 1)
    infer16 = lib.uvc_vit_workspace_bytes(ctypes.byref(d), 0)
    assert 0 < infer16 < train16 < 0.8 * train
    d.H = 4                                # head dim 96: the fp16 mode is refused through the size query's 0
    assert lib.uvc_vit_workspace_bytes(ctypes.byref(d), 1) == 0 and b"fp16 operand storage" in lib.uvc_last_error()


def test_no_cpu_fallback():
    import pytest
    import torch
    from uvc_b200.models import deit_tiny_patch16_224
    from uvc_b200.models.model_distilled import DistilledVisionTransformer
    from functools import partial
    m = DistilledVisionTransformer(enable_dist=0, patch_size=16, embed_dim=192, depth=1, num_heads=3, mlp_ratio=4, qkv_bias=True,
                                   norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), drop_rate=0)
    with pytest.raises(Exception) as e:
        m(torch.zeros(1, 3, 224, 224))
    assert "CUDA" in str(e.value) or "cuda" in str(e.value)


def test_epilogue_flag_values_match_the_header():
    """The UVC_EPI_* / UVC_GEMM_* enum of include/uvc_b200.h and the constants the ctypes binding passes must agree bit for bit."""
    text = open(os.path.join(ROOT, "include", "uvc_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    vals = {k: int(v) for k, v in re.findall(r"\b(UVC_(?:EPI|GEMM)_\w+)\s*=\s*(\d+)", text)}
    want = {"UVC_EPI_BIAS": _lib.EPI_BIAS, "UVC_EPI_GELU": _lib.EPI_GELU, "UVC_EPI_GELU_BWD": _lib.EPI_GELU_BWD, "UVC_EPI_RESIDUAL": _lib.EPI_RESIDUAL,
            "UVC_EPI_ATOMIC": _lib.EPI_ATOMIC, "UVC_EPI_ROUND_TF32": _lib.EPI_ROUND_TF32, "UVC_EPI_COLSUM": _lib.EPI_COLSUM,
            "UVC_GEMM_F16": _lib.GEMM_F16, "UVC_EPI_AUX_F16": _lib.EPI_AUX_F16, "UVC_EPI_BLEND": _lib.EPI_BLEND}
    assert vals == want
    assert len(set(want.values())) == len(want) and all(v & (v - 1) == 0 for v in want.values())      # distinct single bits


def test_deferred_host_values_resolve_lazily():
    """uvc_optimizer(lazy=True) hands back Deferred handles; they must resolve once, through float() / np.asarray() / .tolist() / .size."""
    import numpy as np
    from uvc_b200.uvc_optimizer import Deferred

    class Snap:
        calls = 0

        def numpy(self):
            Snap.calls += 1
            return np.arange(7, dtype=np.float32)
    s = Snap()
    cur, arr = Deferred(s, lambda h: float(h[0] + 0.5)), Deferred(s, lambda h: h[1:7].reshape(3, 2).copy())
    assert Snap.calls == 0                                    # nothing fetched until first use
    assert float(cur) == 0.5 and arr.size == 6 and arr.shape == (3, 2)
    assert np.asarray(arr).tolist() == arr.tolist() == [[1.0, 2.0], [3.0, 4.0], [5.0, 6.0]]
    assert f"{float(cur) * 100:.1f}" == "50.0"
