"""ctypes binding of libdinov3_b200.so (the C ABI declared in include/dinov3_b200.h).

There is deliberately no fallback: if the shared library is missing, or the process has no sm_100 GPU, every
entry point raises.  PyTorch tensors are only containers for device memory here (``tensor.data_ptr()``).
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import torch

_LIB = None
_INIT_DEVICE = None

PKG_ROOT = Path(__file__).resolve().parent.parent  # .../dinov3-jax_b200
LIB_PATH = PKG_ROOT / "libdinov3_b200.so"

# epilogue flags (include/dinov3_b200.h)
EP_BIAS, EP_GELU, EP_STORE_PRE, EP_MUL_DGELU, EP_GAMMA, EP_RESID, EP_OUT_F32, EP_ACCUM = 1, 2, 4, 8, 16, 32, 64, 128


class NativeError(RuntimeError):
    pass


class GemmEpilogue(C.Structure):
    _fields_ = [
        ("bias", C.c_void_p), ("gamma", C.c_void_p), ("resid", C.c_void_p), ("aux_in", C.c_void_p),
        ("aux_out", C.c_void_p), ("out", C.c_void_p),
        ("ld_out", C.c_int), ("ld_aux", C.c_int), ("ld_resid", C.c_int), ("flags", C.c_int), ("alpha", C.c_float),
    ]


def lib() -> C.CDLL:
    """Load the shared library (no GPU needed for loading / symbol checks)."""
    global _LIB
    if _LIB is None:
        path = os.environ.get("DINOV3_B200_LIB", str(LIB_PATH))
        if not os.path.exists(path):
            raise NativeError(
                f"{path} not found: build it with `python __graft_entry__.py` (nvcc, sm_100a). "
                "There is no CPU / PyTorch fallback for the training hot path.")
        _LIB = C.CDLL(path)
        _LIB.d3_last_error.restype = C.c_char_p
        _LIB.d3_launch_count.restype = C.c_longlong
    return _LIB


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        raise NativeError(f"{what} failed ({rc}): {lib().d3_last_error().decode()}")


def init(device: int | None = None) -> C.CDLL:
    """Bind the library to a CUDA device; raises if it is not a Blackwell (sm_100) part."""
    global _INIT_DEVICE
    l = lib()
    if not torch.cuda.is_available():
        raise NativeError("no CUDA device: the dinov3 B200 engine has no CPU fallback")
    if device is None:
        device = torch.cuda.current_device()
    if _INIT_DEVICE != device:
        check(l.d3_init(int(device)), "d3_init")
        _INIT_DEVICE = device
    return l


def stream_ptr() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t: torch.Tensor | None) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


def launch_count() -> int:
    return int(lib().d3_launch_count())


def reset_launch_count() -> None:
    lib().d3_reset_launch_count()
