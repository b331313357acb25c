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
                ("flags", C.c_int32), ("_pad", C.c_int32)]


_ABI_STRUCTS = {"uvc_operand": Operand, "uvc_gemm_args": GemmArgs}

EPI_BIAS, EPI_GELU, EPI_GELU_BWD, EPI_RESIDUAL, EPI_ATOMIC = 1, 2, 4, 8, 16

_lib = None


def lib_path():
    return _build.LIB_PATH


def load():
    """Load (building first if the in-tree .so is missing/stale and nvcc exists). Raises on failure."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB_PATH
    if _build.is_stale():
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
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().uvc_last_error().decode(errors="replace")
        raise UvcError(f"{what} failed (status {rc}): {msg}")
