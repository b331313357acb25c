"""ctypes binding of libuvc_sm100.so (include/uvc_b200.h).

There is NO fallback: if the shared library cannot be loaded the import of any op fails loudly.
ctypes (not pybind) keeps the boundary a plain C ABI with no coupling to torch's C++ ABI.
"""
import ctypes as C
import os

from . import build as _build

_HERE = os.path.dirname(os.path.abspath(__file__))


class UvcError(RuntimeError):
    pass


class Operand(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("ld", C.c_int64), ("bs1", C.c_int64), ("bs2", C.c_int64),
                ("mn_major", C.c_int32), ("_pad", C.c_int32)]


class GemmArgs(C.Structure):
    _fields_ = [("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32), ("nb1", C.c_int32), ("nb2", C.c_int32),
                ("splits", C.c_int32), ("A", Operand), ("B", Operand),
                ("D", C.c_void_p), ("ldd", C.c_int64), ("d_bs1", C.c_int64), ("d_bs2", C.c_int64),
                ("bias", C.c_void_p),
                ("R", C.c_void_p), ("ldr", C.c_int64), ("r_bs1", C.c_int64), ("r_bs2", C.c_int64),
                ("aux", C.c_void_p), ("ldaux", C.c_int64), ("aux_bs1", C.c_int64), ("aux_bs2", C.c_int64),
                ("alpha", C.c_float), ("beta", C.c_float), ("alpha_dev", C.c_void_p), ("beta_dev", C.c_void_p),
                ("flags", C.c_int32), ("colsum_scale", C.c_float), ("colsum", C.c_void_p), ("R2", C.c_void_p), ("ldr2", C.c_int64), ("D2", C.c_void_p), ("ldd2", C.c_int64), ("blend_dev", C.c_void_p),
                ("alpha_dev2", C.c_void_p), ("colsum_scale_dev", C.c_void_p),
                ("D16", C.c_void_p), ("ldd16", C.c_int64)]


_F = C.c_void_p   # device float*


class BlockTensors(C.Structure):
    NAMES = ("norm1_w", "norm1_b", "qkv_w", "qkv_b", "proj_w", "proj_b", "norm2_w", "norm2_b", "fc1_w", "fc1_b", "fc2_w", "fc2_b")
    _fields_ = [(n, _F) for n in NAMES]


class VitTensors(C.Structure):
    NAMES = ("patch_w", "patch_b", "cls_token", "pos_embed", "norm_w", "norm_b", "head_w", "head_b")
    _fields_ = [(n, _F) for n in NAMES] + [("blocks", C.POINTER(BlockTensors))]


class VitDims(C.Structure):
    _fields_ = [("B", C.c_int32), ("img", C.c_int32), ("patch", C.c_int32), ("in_chans", C.c_int32),
                ("C", C.c_int32), ("H", C.c_int32), ("Fh", C.c_int32), ("L", C.c_int32), ("num_classes", C.c_int32),
                ("ln_eps", C.c_float), ("operand_f16", C.c_int32)]


UVC_MAX_DEPTH = 32


class VitLayout(C.Structure):
    """uvc_vit_layout: Stage-2 compaction (live heads / neurons per block)"""
    _fields_ = [("n_heads", C.c_int32 * UVC_MAX_DEPTH), ("n_neurons", C.c_int32 * UVC_MAX_DEPTH), ("head_idx", C.c_void_p), ("neuron_idx", C.c_void_p)]


class VitForwardArgs(C.Structure):
    _fields_ = [("dims", VitDims), ("w", VitTensors), ("x", _F), ("blend", _F), ("skip_host", C.c_void_p),
                ("patch_scale", _F), ("token_mask", _F), ("save_for_backward", C.c_int32), ("enable_jumping", C.c_int32),
                ("logits", _F), ("pe_out", _F), ("workspace", C.c_void_p), ("workspace_bytes", C.c_uint64), ("pe_in", _F),
                ("layout", C.POINTER(VitLayout)), ("weights_converted", C.c_int32)]


class VitBackwardArgs(C.Structure):
    _fields_ = [("dims", VitDims), ("w", VitTensors), ("g", VitTensors), ("dlogits", _F), ("blend", _F),
                ("skip_host", C.c_void_p), ("patch_scale", _F), ("token_mask", _F), ("enable_jumping", C.c_int32),
                ("grad_scale", C.c_float), ("d_blend", _F), ("d_patch_scale", _F), ("d_token_mask", _F),
                ("workspace", C.c_void_p), ("workspace_bytes", C.c_uint64), ("d_pe", _F), ("layout", C.POINTER(VitLayout))]


class PerformerTensors(C.Structure):
    NAMES = ("norm1_w", "norm1_b", "kqv_w", "kqv_b", "w", "proj_w", "proj_b", "norm2_w", "norm2_b", "mlp0_w", "mlp0_b", "mlp2_w", "mlp2_b")
    _fields_ = [(n, _F) for n in NAMES]


class T2TTensors(C.Structure):
    _fields_ = [("attn1", PerformerTensors), ("attn2", PerformerTensors), ("project_w", _F), ("project_b", _F)]


class T2TDims(C.Structure):
    _fields_ = [("B", C.c_int32), ("img", C.c_int32), ("in_chans", C.c_int32), ("C", C.c_int32), ("ln_eps", C.c_float)]


class T2TForwardArgs(C.Structure):
    _fields_ = [("dims", T2TDims), ("w", T2TTensors), ("x", _F), ("tokens", _F), ("save_for_backward", C.c_int32), ("dropout_p", C.c_float),
                ("seed", C.c_uint64), ("workspace", C.c_void_p), ("workspace_bytes", C.c_uint64)]


class T2TBackwardArgs(C.Structure):
    _fields_ = [("dims", T2TDims), ("w", T2TTensors), ("g", T2TTensors), ("x", _F), ("d_tokens", _F), ("dropout_p", C.c_float), ("seed", C.c_uint64),
                ("grad_scale", C.c_float), ("workspace", C.c_void_p), ("workspace_bytes", C.c_uint64)]


class AdmmArgs(C.Structure):
    _PP = C.POINTER(C.c_void_p)
    _fields_ = [("L", C.c_int32), ("H", C.c_int32), ("d", C.c_int32), ("Fh", C.c_int32),
                ("w1", _PP), ("w3", _PP), ("m1", _PP), ("m3", _PP), ("m2", _PP),
                ("c1", _F), ("c2", _F), ("c3", _F), ("rank1", _F), ("rank2", _F), ("rank3", _F),
                ("s", _F), ("r", _F), ("y", _F), ("p", _F), ("z", _F),
                ("gate", _F), ("gate_grad", _F), ("noise", _F), ("macs", _F),
                ("embed_macs", C.c_double), ("full_flops", C.c_double), ("lr", C.c_double),
                ("budget", C.c_float), ("z_grad_clip", C.c_float), ("slr", C.c_float), ("rlr", C.c_float), ("ylr", C.c_float),
                ("plr", C.c_float), ("zlr", C.c_float), ("sl2wd", C.c_float), ("gating_weight", C.c_float), ("eps", C.c_float),
                ("use_gumbel", C.c_int32), ("gumbel_hard", C.c_int32), ("warmup", C.c_int32), ("gate_mult", C.c_float),
                ("gate_grad_acc", _F), ("out", _F)]


_ABI_STRUCTS = {"uvc_admm_args": AdmmArgs, "uvc_operand": Operand, "uvc_gemm_args": GemmArgs, "uvc_block_tensors": BlockTensors,
                "uvc_vit_tensors": VitTensors, "uvc_vit_dims": VitDims, "uvc_vit_layout": VitLayout, "uvc_vit_forward_args": VitForwardArgs,
                "uvc_vit_backward_args": VitBackwardArgs, "uvc_t2t_tensors": T2TTensors, "uvc_t2t_forward_args": T2TForwardArgs,
                "uvc_t2t_backward_args": T2TBackwardArgs}

# every symbol include/uvc_b200.h declares (checked by tests/test_abi.py)
EXPORTS = [
    "uvc_version", "uvc_last_error", "uvc_abi_sizeof", "uvc_launch_count", "uvc_gemm_profile", "uvc_gemm_profile_read", "uvc_gemm_profile_read_kind", "uvc_gemm_tf32",
    "uvc_layernorm_fwd", "uvc_layernorm_bwd", "uvc_layernorm_bwd_cs", "uvc_softmax_fwd", "uvc_softmax_bwd", "uvc_colsum", "uvc_blend_fwd",
    "uvc_blend_dots", "uvc_im2col16", "uvc_assemble_tokens", "uvc_assemble_tokens_bwd", "uvc_round_tf32", "uvc_scale_add",
    "uvc_attn_ldp", "uvc_attention_fwd", "uvc_attention_bwd", "uvc_attention_fwd_lse", "uvc_attention_bwd_fused", "uvc_distill_loss", "uvc_sqnorm_accum", "uvc_clip_adamw", "uvc_sqnorm_accum_flags", "uvc_clip_adamw_flags",
    "uvc_vit_workspace_bytes", "uvc_vit_forward", "uvc_vit_backward",
    "uvc_layernorm_fwd_f16", "uvc_layernorm_bwd_f16", "uvc_cvt_f16", "uvc_attention_fwd_f16", "uvc_attention_bwd_f16",
    "uvc_token_gate_fold", "uvc_token_gate_fwd", "uvc_token_gate_bwd", "uvc_token_gate_apply",
    "uvc_mixup",
    "uvc_t2t_workspace_bytes", "uvc_t2t_forward", "uvc_t2t_backward",
    "uvc_admm_scores", "uvc_admm_prox", "uvc_admm_masks", "uvc_admm_primal", "uvc_admm_dual", "uvc_admm_resource",
]

EPI_BIAS, EPI_GELU, EPI_GELU_BWD, EPI_RESIDUAL, EPI_ATOMIC, EPI_ROUND_TF32, EPI_COLSUM, GEMM_F16, EPI_AUX_F16, EPI_BLEND = 1, 2, 4, 8, 16, 32, 64, 128, 256, 512

_lib = None

# Precision mode of the engine (uvc_vit_dims.operand_f16).  Default: fp16 operand storage wherever the kernels support it; UVC_PRECISION=tf32 (or
# model.operand_f16 = False) selects the fp32-storage / TF32 path of round 1.  The fp16 gradient operands carry a power-of-two loss scale that
# the engine picks on the device from max|dlogits| (or model.grad_scale if set > 0) and removes again wherever it produces an fp32 gradient.


def operand_f16_for(model, d, ntok, C_, Fh):
    want = getattr(model, "operand_f16", None)
    if want is None:
        want = os.environ.get("UVC_PRECISION", "f16").lower() not in ("tf32", "fp32", "0")
    return bool(want) and d == 64 and ntok <= 208 and C_ % 8 == 0 and Fh % 8 == 0


def lib_path():
    return _build.LIB_PATH


def load():
    """Load (building first if the in-tree .so is missing/stale and nvcc exists). Raises on failure."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("UVC_LIB_PATH") or _build.LIB_PATH       # UVC_LIB_PATH: bring-up only (an alternative build of the same sources)
    if path == _build.LIB_PATH and _build.is_stale():
        try:
            _build.build_library()
        except Exception as e:  # no nvcc on this box and no prebuilt library
            if not os.path.isfile(path):
                raise UvcError(f"libuvc_sm100.so is missing and could not be built: {e}") from e
    try:
        lib = C.CDLL(path)
    except OSError as e:
        raise UvcError(f"cannot load {path}: {e}") from e
    lib.uvc_version.restype = C.c_int
    lib.uvc_last_error.restype = C.c_char_p
    lib.uvc_abi_sizeof.restype = C.c_int
    lib.uvc_abi_sizeof.argtypes = [C.c_char_p]
    for name, st in _ABI_STRUCTS.items():
        n = lib.uvc_abi_sizeof(name.encode())
        if n != C.sizeof(st):
            raise UvcError(f"ABI mismatch for {name}: library {n} bytes, binding {C.sizeof(st)} bytes")
    i32, i64, f32, vp = C.c_int32, C.c_int64, C.c_float, C.c_void_p
    protos = {
        "uvc_gemm_tf32": [C.POINTER(GemmArgs), vp],
        "uvc_layernorm_fwd": [vp, i64, vp, vp, f32, vp, i64, vp, vp, i32, i32, i32, vp],
        "uvc_layernorm_bwd": [vp, i64, vp, i64, vp, vp, vp, vp, vp, vp, vp, i64, vp, vp, i32, i32, vp],
        "uvc_layernorm_bwd_cs": [vp, i64, vp, i64, vp, vp, vp, vp, vp, vp, vp, i64, vp, vp, vp, vp, i32, i32, vp],
        "uvc_layernorm_fwd_f16": [vp, i64, vp, vp, f32, vp, i64, vp, vp, i32, i32, vp],
        "uvc_layernorm_bwd_f16": [vp, i64, f32, vp, i64, vp, vp, vp, vp, vp, vp, vp, vp, f32, i64, vp, vp, vp, vp, i32, i32, vp],
        "uvc_cvt_f16": [vp, vp, vp, i32, i32, vp],
        "uvc_attention_fwd_f16": [vp, vp, vp, i32, i32, i32, i32, f32, vp],
        "uvc_attention_bwd_f16": [vp, vp, vp, vp, vp, vp, vp, f32, i32, i32, i32, i32, f32, vp],
        "uvc_token_gate_fold": [vp, vp, vp, i32, i32, vp, vp, vp],
        "uvc_token_gate_fwd": [vp, i64, i32, vp, vp, vp, vp, vp, f32, i32, i32, i32, vp, vp, vp, vp, vp],
        "uvc_token_gate_bwd": [vp, vp, vp, f32, i32, i32, vp, vp],
        "uvc_token_gate_apply": [vp, vp, vp, vp, i32, i32, i32, vp, vp, vp, vp, vp],
        "uvc_softmax_fwd": [vp, i64, i64, i32, i32, vp],
        "uvc_softmax_bwd": [vp, vp, i64, i64, i32, f32, i32, vp],
        "uvc_colsum": [vp, i64, i32, i32, vp, vp, vp],
        "uvc_blend_fwd": [vp, vp, vp, vp, i64, vp],
        "uvc_blend_dots": [vp, vp, vp, vp, i64, vp],
        "uvc_im2col16": [vp, vp, i32, i32, i32, i32, i32, vp],
        "uvc_round_tf32": [vp, vp, i64, vp],
        "uvc_assemble_tokens": [vp, vp, vp, vp, vp, vp, i32, i32, i32, vp],
        "uvc_assemble_tokens_bwd": [vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, vp],
        "uvc_scale_add": [vp, vp, vp, f32, i64, vp],
        "uvc_attention_fwd": [vp, vp, vp, i32, i32, i32, i32, f32, vp],
        "uvc_attention_bwd": [vp, vp, vp, vp, vp, i32, i32, i32, i32, f32, vp],
        "uvc_attention_fwd_lse": [vp, vp, vp, i32, i32, i32, i32, f32, vp],
        "uvc_attention_bwd_fused": [vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, f32, vp],
        "uvc_distill_loss": [vp, vp, vp, i32, i32, f32, f32, f32, vp, vp, vp],
        "uvc_mixup": [vp, vp, vp, i32, i32, i32, i32, i32, f32, f32, i32, i32, i32, i32, i32, vp],
        "uvc_sqnorm_accum": [vp, i64, vp, vp],
        "uvc_sqnorm_accum_flags": [vp, vp, i64, vp, vp],
        "uvc_clip_adamw_flags": [vp, vp, vp, vp, vp, i64, vp, f32, f32, f32, f32, f32, f32, i32, vp],
        "uvc_clip_adamw": [vp, vp, vp, vp, vp, i64, vp, f32, f32, f32, f32, f32, f32, i32, vp],
        "uvc_vit_forward": [C.POINTER(VitForwardArgs), vp],
        "uvc_vit_backward": [C.POINTER(VitBackwardArgs), vp],
        "uvc_t2t_forward": [C.POINTER(T2TForwardArgs), vp],
        "uvc_t2t_backward": [C.POINTER(T2TBackwardArgs), vp],
    }
    for name in ("scores", "prox", "masks", "primal", "dual", "resource"):
        protos["uvc_admm_" + name] = [C.POINTER(AdmmArgs), vp]
    for name, argtypes in protos.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = C.c_int
    lib.uvc_launch_count.argtypes = []; lib.uvc_launch_count.restype = C.c_longlong
    lib.uvc_gemm_profile.argtypes = [C.c_int]; lib.uvc_gemm_profile.restype = C.c_int
    lib.uvc_gemm_profile_read.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_longlong)]
    lib.uvc_gemm_profile_read.restype = C.c_int
    lib.uvc_gemm_profile_read_kind.argtypes = [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_longlong)]
    lib.uvc_gemm_profile_read_kind.restype = C.c_int
    lib.uvc_attn_ldp.argtypes = [i32]; lib.uvc_attn_ldp.restype = i32
    lib.uvc_vit_workspace_bytes.argtypes = [C.POINTER(VitDims), i32]; lib.uvc_vit_workspace_bytes.restype = C.c_uint64
    lib.uvc_t2t_workspace_bytes.argtypes = [C.POINTER(T2TDims), i32]; lib.uvc_t2t_workspace_bytes.restype = C.c_uint64
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().uvc_last_error().decode(errors="replace")
        raise UvcError(f"{what} failed (status {rc}): {msg}")
