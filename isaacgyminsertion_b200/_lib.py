"""ctypes binding of libigi_b200.so (the C-ABI in include/igi_b200.h).

The product path has no CPU fallback: if the shared library is missing or a
tensor is not a CUDA tensor the call raises.  Build with
`python -m isaacgyminsertion_b200.build` (or `__graft_entry__.build()`).
"""
import ctypes
import os
import re

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libigi_b200.so")
HEADER_PATH = os.path.join(_HERE, "..", "include", "igi_b200.h")

_lib = None

_c = ctypes
_P = _c.c_void_p


def declared_symbols(header=HEADER_PATH):
    """Names of every function include/igi_b200.h declares."""
    text = open(header).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(igi_[a-z0-9_]+)\s*\(", text)))


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not built: run `python -m isaacgyminsertion_b200.build` "
            "(there is no CPU fallback for the observation hot path)")
    lib = ctypes.CDLL(LIB_PATH)
    lib.igi_version.restype = _c.c_int
    lib.igi_last_error.restype = _c.c_char_p
    for name in declared_symbols():
        fn = getattr(lib, name)  # raises AttributeError if a declared symbol is not exported
        if name in ("igi_launch_count", "igi_rms_scratch_bytes"):
            fn.restype = _c.c_longlong
        elif name not in ("igi_last_error",):
            fn.restype = _c.c_int
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().igi_last_error().decode(errors="replace")
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")


class on_device:
    """`with on_device(dev):` makes `dev` the current CUDA device for the launches inside when it is not already
    (an engine may live on a device other than the caller's current one); a no-op otherwise."""

    def __init__(self, device):
        d = torch.device(device) if not isinstance(device, torch.device) else device
        self.idx = d.index
        self.prev = None

    def __enter__(self):
        if self.idx is not None:
            cur = torch.cuda.current_device()
            if cur != self.idx:
                self.prev = cur
                torch.cuda.set_device(self.idx)
        return self

    def __exit__(self, *exc):
        if self.prev is not None:
            torch.cuda.set_device(self.prev)
        return False


def stream_ptr(device=None):
    return _P(torch.cuda.current_stream(device).cuda_stream)


def dptr(t, dtype=None, name="tensor"):
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return _P(0)
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"{name}: expected a CUDA tensor (no CPU fallback)")
    if dtype is not None and t.dtype != dtype:
        raise RuntimeError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise RuntimeError(f"{name}: tensor must be contiguous")
    return _P(t.data_ptr())


def carr(ctype, values):
    return (ctype * len(values))(*values)
