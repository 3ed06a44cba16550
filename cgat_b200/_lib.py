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
# CGAT_B200_LIB=trap selects the debug build whose mbarrier waits are bounded (python -m cgat_b200.build --trap-barriers)
_VARIANT = os.environ.get("CGAT_B200_LIB", "")
LIB_PATH = os.path.join(_HERE, f"libcgat_b200_{_VARIANT}.so" if _VARIANT else "libcgat_b200.so")
_lib = None

_P, _I64, _I32, _F32, _SZ = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_float, ctypes.c_size_t

ABI_VERSION = 7  # CGAT_B200_ABI_VERSION in include/cgat_b200.h

# name -> (restype, argtypes); must list every symbol declared in include/cgat_b200.h
SIGNATURES = {
    "cgat_abi_version": (ctypes.c_int, []),
    "cgat_last_error": (ctypes.c_char_p, []),
    "cgat_launch_count": (_I64, []),
    "cgat_csr_workspace_bytes": (_SZ, [_I64, _I64]),
    "cgat_csr_build": (ctypes.c_int, [_P, _P, _I64, _I64, _P, _P, _P, _P, _P, _I32, _P, _SZ, _P]),
    "cgat_status_flags": (ctypes.c_int, [_P, _I32]),
    "cgat_segment_ptr": (ctypes.c_int, [_P, _I64, _I64, _P, _P, _P]),
    "cgat_seg_softmax_fwd": (ctypes.c_int, [_P, _P, _P, _P, _I64, _I32, _I32, _I32, _F32, _P, _P, _P, _P]),
    "cgat_seg_softmax_bwd": (ctypes.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _I64, _I32, _I32, _I32, _F32,
                                            _P, _P, _P]),
    "cgat_gemm3x_nt": (ctypes.c_int, [_P, _I64, _P, _I64, _P, _P, _I64, _I64, _I64, _I64, _I32, _P]),
    "cgat_gemm3x_nt_res": (ctypes.c_int, [_P, _I64, _P, _P, _P, _I64, _I64, _I64, _I64, _I32, _P]),
    "cgat_gemm3x_nt_splitk": (ctypes.c_int, [_P, _I64, _P, _I64, _P, _I64, _I64, _I64, _I64, _I64, _I32, _P]),
    "cgat_gemm3x_tn": (ctypes.c_int, [_P, _I64, _P, _I64, _P, _I64, _I64, _I64, _I64, _I64, _I32, _P]),
    "cgat_packed_floats": (_I64, [_I64, _I64]),
    "cgat_pack_kmajor": (ctypes.c_int, [_P, _I64, _I64, _I64, _I32, _P, _P]),
    "cgat_packed_floats_f16": (_I64, [_I64, _I64]),
    "cgat_pack_kmajor_f16": (ctypes.c_int, [_P, _I64, _I64, _I64, _I32, _P, _P]),
    "cgat_pack_kmajor_f16s": (ctypes.c_int, [_P, _I64, _I64, _I64, _I32, _F32, _F32, _P, _P]),
    "cgat_edge_attn_fwd_f16": (ctypes.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I64, _I64, _I32,
                                              _I32, _I32, _F32, _P]),
    "cgat_edge_attn_bwd_prep_f16": (ctypes.c_int, [_P] * 19 + [_I64, _I64, _I32, _I32, _I32, _F32, _P]),
    "cgat_edge_attn_wgrad_f16_splits": (_I32, [_I32, _I32, _I32]),
    "cgat_edge_attn_wgrad_f16": (ctypes.c_int, [_P] * 9 + [_I64, _I32, _I32, _I32, _P]),
    "cgat_hyper_rowdot_fwd_f16": (ctypes.c_int, [_P, _P, _P, _P, _P, _P, _P, _I64, _I32, _P]),
    "cgat_hyper_rowscale_f16": (ctypes.c_int, [_P, _P, _P, _P, _P, _I64, _I32, _P]),
    "cgat_hyper_rowdot_fwd": (ctypes.c_int, [_P, _P, _P, _P, _P, _P, _P, _I64, _I32, _P]),
    "cgat_hyper_trunk_packed_floats": (_I64, [_I32, _I32]),
    "cgat_hyper_trunk_pack": (ctypes.c_int, [_P, _P, _I32, _I32, _I32, _P, _P]),
    "cgat_hyper_trunk_fwd": (ctypes.c_int, [_P, _P, _P, _P, _P, _I64, _I32, _I32, _P]),
    "cgat_hyper_trunk_bwd": (ctypes.c_int, [_P, _P, _P, _P, _P, _P, _I64, _I32, _I32, _P]),
    "cgat_gemm3x_tn_batched": (ctypes.c_int, [_P, _P, _I32, _I64, _I64, _P, _P, _I64, _I64, _I64, _I32, _P]),
    "cgat_hyper_rowscale_parts": (_I32, [_I64, _I32]),
    "cgat_hyper_rowscale_parts_f16": (_I32, [_I64, _I32]),
    "cgat_hyper_rowscale": (ctypes.c_int, [_P, _P, _P, _P, _P, _I64, _I32, _P]),
    "cgat_hyper_rowscale_f16_amax": (ctypes.c_int, [_P, _P, _P, _P, _P, _P, _I64, _I32, _P]),
    "cgat_hyper_wgrad_f16": (ctypes.c_int, [_P, _P, _P, _P, _P, _P, _I64, _I32, _P]),
    "cgat_hyper_wgrad_splits": (_I32, [_I64]),
    "cgat_hyper_wgrad": (ctypes.c_int, [_P, _P, _P, _P, _I64, _I32, _P]),
    "cgat_edge_attn_fwd": (ctypes.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I64, _I64, _I32,
                                          _I32, _I32, _F32, _P]),
    "cgat_edge_attn_bwd_prep": (ctypes.c_int, [_P] * 19 + [_I64, _I64, _I32, _I32, _I32, _F32, _P]),
    "cgat_edge_attn_grid": (_I32, [_I64]),
    "cgat_edge_attn_dgrad_grid": (_I32, [_I64]),
    "cgat_edge_attn_dgrad": (ctypes.c_int, [_P] * 10 + [_I64, _I32, _P, _I32, _P, _I64, _I64, _I32, _I32, _I32, _P]),
    "cgat_edge_attn_dgrad_f16": (ctypes.c_int, [_P] * 9 + [_I64, _I64, _I32, _I32, _I32, _P]),
    "cgat_edge_attn_reduce_chunks": (_I32, [_I64]),
    "cgat_edge_attn_reduce_groups": (_I32, [_I64]),
    "cgat_edge_attn_reduce_counters": (_I32, [_I64, _I32]),
    "cgat_edge_attn_reduce": (ctypes.c_int, [_P, _I64, _P, _P, _P, _P, _P, _I64, _I32, _I32, _P, _P, _P, _I32, _I64, _I32, _P]),
    "cgat_edge_attn_wgrad_splits": (_I32, [_I32]),
    "cgat_edge_attn_wgrad": (ctypes.c_int, [_P] * 8 + [_I64, _I32, _I32, _I32, _P]),
    "cgat_sum_parts": (ctypes.c_int, [_P, _I32, _I64, _P, _I64, _I32, _P]),
    "cgat_sum_parts_bias_act": (ctypes.c_int, [_P, _I32, _I64, _P, _P, _I64, _I64, _I32, _P]),
    "cgat_adamw_flat": (ctypes.c_int, [_P, _P, _P, _P, _I64, _P, _P, _F32, _F32, _F32, _F32, _F32, _P]),
    "cgat_l1_loss": (ctypes.c_int, [_P, _I64, _P, _I64, _P, _P, _I64, _I64, _I32, _P]),
    "cgat_collate_plan": (ctypes.c_int, [_P, _I64, _P, _P, _P, _P, _P, _P]),
    "cgat_collate_fill": (ctypes.c_int, [_P, _I64] + [_P] * 8 + [_I32, _I32] + [_P] * 13 + [_I64] * 6 + [_P]),
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
    if lib.cgat_abi_version() != ABI_VERSION:
        raise CgatLibraryError(f"{LIB_PATH} has ABI version {lib.cgat_abi_version()}, this package binds version "
                               f"{ABI_VERSION}: rebuild with `python -m cgat_b200.build --force`")
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


def ptr_array(tensors):
    """HOST array of device pointers (for the entry points that take `const float* const*`)."""
    return (ctypes.c_void_p * len(tensors))(*[ptr(t) for t in tensors])


def i64_array(values):
    return (ctypes.c_int64 * len(values))(*[int(v) for v in values])


_prof = None  # list of (key, start_event, end_event, work) while profiling


def call(name, *args, work=None):
    """Invoke one C-ABI entry point on the current stream.  `work` (optional) declares the
    algorithmic bytes / flops of this launch for bench.py's roofline:
    dict(key=..., bound='hbm'|'tensor', bytes=..., flops=..., note=...)."""
    lib = load()
    if _prof is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    rc = getattr(lib, name)(*args)
    if _prof is not None:
        e1.record()
        _prof.append((name, e0, e1, work or {}))
    if rc != 0:
        raise CgatLibraryError(f"{name} failed: {lib.cgat_last_error().decode()}")


def profile_begin():
    global _prof
    _prof = []


def profile_end():
    """-> {key: dict(ms, launches, bytes, flops, bound, note)} summed over the profiled launches."""
    global _prof
    rows, rec = {}, _prof
    _prof = None
    torch.cuda.synchronize()
    for name, e0, e1, work in rec:
        key = work.get("key", name)
        r = rows.setdefault(key, dict(ms=0.0, launches=0, bytes=0.0, flops=0.0, bound=work.get("bound", "hbm"),
                                      note=work.get("note", "")))
        r["ms"] += e0.elapsed_time(e1)
        r["launches"] += 1
        r["bytes"] += float(work.get("bytes", 0))
        r["flops"] += float(work.get("flops", 0))
    return rows


STATUS_BITS = {1: "an edge destination (edge_index[1]) is outside [0, num_nodes): the edge was dropped",
               2: "an edge source (edge_index[0]) is outside [0, num_nodes): clamped to 0",
               4: "a shell rank (edge_attr) is outside the rank-embedding table: clamped to 0",
               8: "a non-finite aggregate left the fused edge attention (fp16 operand overflow of a diverged network, "
                  "or NaN / Inf inputs); CGAT_B200_F16X3_EDGE=0 selects the tf32 kernels"}


def status_flags(reset=True):
    """Sticky validation flags raised by kernels since the last reset (include/cgat_b200.h: cgat_status_flags).
    Synchronises the device."""
    out = ctypes.c_uint32(0)
    rc = load().cgat_status_flags(ctypes.byref(out), int(reset))
    if rc != 0:
        raise CgatLibraryError(f"cgat_status_flags failed: {load().cgat_last_error().decode()}")
    return int(out.value)


def check_status(reset=True):
    """Raise if any kernel flagged invalid input / a non-finite result since the last call (the reference raises an
    IndexError from nn.Embedding / index_select at the same inputs).  Synchronises: call it after a step, not inside."""
    v = status_flags(reset)
    if v:
        raise CgatLibraryError("cgat_b200: " + "; ".join(msg for bit, msg in STATUS_BITS.items() if v & bit))


def launch_count():
    return int(load().cgat_launch_count())
