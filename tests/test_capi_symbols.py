"""-m "not gpu": the C-ABI library loads here (no GPU) and exports every symbol include/cgat_b200.h declares."""
import ctypes
import os
import re

from cgat_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "cgat_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cgat_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported():
    build.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 8
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/cgat_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == names, "ctypes binding table and header disagree"
    lib.cgat_abi_version.restype = ctypes.c_int
    assert lib.cgat_abi_version() >= 1


def test_no_cpu_fallback():
    """Kernel-backed ops refuse CPU tensors instead of silently computing something else."""
    import pytest
    import torch
    from cgat_b200 import graph
    ei = torch.zeros((2, 4), dtype=torch.int64)
    with pytest.raises(Exception):
        graph.build_edge_plan(ei, torch.ones(4, dtype=torch.int64), 2)
