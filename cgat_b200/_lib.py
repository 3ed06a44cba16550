"""ctypes binding of libcgat_b200.so (the C ABI declared in include/cgat_b200.h).

There is no CPU fallback: if the library is missing, or a kernel is asked to run on non-CUDA
tensors, the call raises.  Build the library with `python -m cgat_b200.build` (or
`__graft_entry__.build()`); it lives in-tree next to this file.
"""
from __future__ import annotations

import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcgat_b200.so")
_lib = None

_P, _I64, _I32, _F32, _SZ = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_float, ctypes.c_size_t

# name -> (restype, argtypes); must list every symbol declared in include/cgat_b200.h
SIGNATURES = {
    "cgat_abi_version": (ctypes.c_int, []),
    "cgat_last_error": (ctypes.c_char_p, []),
    "cgat_launch_count": (_I64, []),
    "cgat_csr_workspace_bytes": (_SZ, [_I64, _I64]),
    "cgat_csr_build": (ctypes.c_int, [_P, _P, _I64, _I64, _P, _P, _P, _P, _P, _P, _SZ, _P]),
    "cgat_segment_ptr": (ctypes.c_int, [_P, _I64, _I64, _P, _P, _P]),
    "cgat_seg_softmax_fwd": (ctypes.c_int, [_P, _P, _P, _P, _I64, _I32, _I32, _I32, _F32, _P, _P, _P, _P]),
    "cgat_seg_softmax_bwd": (ctypes.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _I64, _I32, _I32, _I32, _F32,
                                            _P, _P, _P]),
}


class CgatLibraryError(RuntimeError):
    pass


def load():
    """Load the shared library (once). Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CgatLibraryError(
            f"{LIB_PATH} not found: the CUDA library is required (no CPU fallback). "
            "Build it with `python -m cgat_b200.build`.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def ptr(t):
    """Device pointer of a tensor (None -> NULL). Enforces CUDA + contiguous."""
    if t is None:
        return None
    if not t.is_cuda:
        raise CgatLibraryError("cgat_b200 kernels need CUDA tensors (there is no CPU path)")
    if not t.is_contiguous():
        raise CgatLibraryError("cgat_b200 kernels need contiguous tensors")
    return t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


def call(name, *args):
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise CgatLibraryError(f"{name} failed: {lib.cgat_last_error().decode()}")


def launch_count():
    return int(load().cgat_launch_count())
