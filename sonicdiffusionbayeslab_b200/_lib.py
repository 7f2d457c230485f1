"""ctypes binding of ``libsonic.so`` (the C ABI declared in ``include/sonic.h``).

There is deliberately no fallback: if the shared library is missing, or a call fails, an
exception is raised -- the product path never silently degrades to PyTorch or CPU code.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SONIC_LIB") or os.path.join(_HERE, "libsonic.so")      # SONIC_LIB: A/B kernel experiments


class SonicError(RuntimeError):
    pass


class GemmArgs(C.Structure):
    _fields_ = [
        ("a0", C.c_void_p), ("c0", C.c_int32), ("ld0", C.c_int32),
        ("a1", C.c_void_p), ("c1", C.c_int32), ("ld1", C.c_int32),
        ("n_img", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
        ("w", C.c_void_p), ("N", C.c_int32), ("taps", C.c_int32),
        ("bias", C.c_void_p), ("row_bias", C.c_void_p),
        ("residual", C.c_void_p), ("ld_res", C.c_int32),
        ("out", C.c_void_p), ("ld_out", C.c_int32),
        ("epilogue", C.c_int32), ("block_n", C.c_int32),
        ("gn_partial", C.c_void_p),
        ("ln_stats_out", C.c_void_p), ("row_scale", C.c_void_p),
        ("stride", C.c_int32), ("upsample", C.c_int32),
    ]


_lib = None


def lib() -> C.CDLL:
    """Load libsonic.so once.  torch is imported first so both share one CUDA runtime."""
    global _lib
    if _lib is None:
        import torch  # noqa: F401  (loads libcudart / creates the primary context owner)

        if not os.path.exists(LIB_PATH):
            raise SonicError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). There is no fallback path.")
        _lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        _lib.sonic_last_error.restype = C.c_char_p
        _lib.sonic_version.restype = C.c_char_p
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise SonicError(f"{what} failed (rc={rc}): {lib().sonic_last_error().decode(errors='replace')}")


def stream_ptr() -> C.c_void_p:
    import torch

    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())
