"""ctypes binding of libfedcola_b200.so (the C ABI declared in include/fedcola_b200.h).

There is NO fallback: if the library is missing or a call fails, a RuntimeError is raised."""
import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfedcola_b200.so")
_lib = None
_lock = threading.Lock()

c_int, c_ll, c_f, c_vp, c_d = ctypes.c_int, ctypes.c_longlong, ctypes.c_float, ctypes.c_void_p, ctypes.c_double


def lib():
    """Load (building first if sources changed and nvcc exists) and return the CDLL."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            from . import build as _build
            _build.build()
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"fedcola_b200: {LIB_PATH} is missing and could not be built; "
                               "there is no CPU fallback for the hot path")
        L = ctypes.CDLL(LIB_PATH)
        L.fc_last_error.restype = ctypes.c_char_p
        if L.fc_abi_version() != 1:
            raise RuntimeError("fedcola_b200: ABI version mismatch")
        _lib = L
        return _lib


def check(code, what=""):
    if code != 0:
        msg = lib().fc_last_error().decode(errors="replace")
        raise RuntimeError(f"fedcola_b200.{what} failed ({code}): {msg}")


def ptr(t):
    """Device/host address of a torch tensor (or None -> NULL)."""
    return c_vp(0) if t is None else c_vp(t.data_ptr())


def stream_ptr(device=None):
    import torch
    return c_vp(torch.cuda.current_stream(device).cuda_stream)


def require_cuda(t, name="tensor"):
    if not t.is_cuda:
        raise RuntimeError(f"fedcola_b200: {name} must live on a CUDA device (no CPU fallback)")
