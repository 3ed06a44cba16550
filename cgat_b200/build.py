"""In-tree nvcc build of libcgat_b200.so for sm_100a (no torch headers involved: plain C ABI).

    python -m cgat_b200.build [--force] [--trap-barriers] [--variant NAME -DMACRO[=V] ...]

Every .cu under csrc/ is compiled to an object file (in parallel, cached by a digest of the source, the
headers and the flags) and linked into one shared library next to this file.  `--trap-barriers` builds
libcgat_b200_trap.so with -DCGAT_MBAR_TRAP: every mbarrier wait is bounded and a wait that never completes
prints the barrier's name and traps (select it with CGAT_B200_LIB=trap; tests/test_gpu_gemm.py runs the
split-K sweep on it)."""
from __future__ import annotations

import glob
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(ROOT, "csrc")
OBJ = os.path.join(ROOT, "build")
LIB = os.path.join(ROOT, "libcgat_b200.so")
LIB_TRAP = os.path.join(ROOT, "libcgat_b200_trap.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17", "--expt-relaxed-constexpr",
    "-Xcompiler", "-fPIC", "-Xptxas", "-v",
]
TRAP_FLAGS = ["-DCGAT_MBAR_TRAP=200000"]   # x 20 us per poll = 4 s per wait before the trap


def _sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _headers_digest(flags):
    h = hashlib.sha256()
    files = sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + [
        os.path.join(ROOT, "..", "include", "cgat_b200.h"), os.path.abspath(__file__)]
    for f in files:
        with open(f, "rb") as fh:
            h.update(f.encode() + b"\0" + fh.read())
    h.update(" ".join(flags).encode())
    return h.hexdigest()


def _compile_one(src, flags, hdig, tag, verbose):
    name = os.path.splitext(os.path.basename(src))[0]
    obj = os.path.join(OBJ, f"{name}{tag}.o")
    stamp = obj + ".stamp"
    with open(src, "rb") as fh:
        dig = hashlib.sha256(hdig.encode() + fh.read()).hexdigest()
    if os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read().strip() == dig:
        return obj, "", False
    nvcc = os.environ.get("NVCC", "nvcc")
    res = subprocess.run([nvcc] + flags + ["-c", "-o", obj, src], capture_output=True, text=True)
    log = res.stdout + res.stderr
    if res.returncode != 0:
        sys.stderr.write(log)
        raise RuntimeError(f"nvcc failed on {src}")
    if verbose:
        sys.stderr.write(log)
    with open(stamp, "w") as fh:
        fh.write(dig)
    return obj, log, True


def build(force=False, verbose=False, trap=False, variant=None, defines=()):
    """Compile every .cu under csrc/ into one shared library. Returns the library path.
    `variant` + `defines` build libcgat_b200_<variant>.so with extra -D flags (experiments; CGAT_B200_LIB=<variant>)."""
    flags = NVCC_FLAGS + (TRAP_FLAGS if trap else []) + [f"-D{d}" for d in defines]
    lib, tag = (LIB_TRAP, "_trap") if trap else (LIB, "")
    if variant:
        lib, tag = os.path.join(ROOT, f"libcgat_b200_{variant}.so"), f"_{variant}"
    os.makedirs(OBJ, exist_ok=True)
    if force:
        for f in glob.glob(os.path.join(OBJ, f"*{tag}.o.stamp")):
            os.remove(f)
    hdig = _headers_digest(flags)
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        results = list(ex.map(lambda s: _compile_one(s, flags, hdig, tag, verbose), _sources()))
    objs = [r[0] for r in results]
    if not any(r[2] for r in results) and os.path.exists(lib) and not force:
        return lib
    nvcc = os.environ.get("NVCC", "nvcc")
    res = subprocess.run([nvcc, "-shared", "-o", lib] + objs + ["-lcudart"], capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError(f"linking {os.path.basename(lib)} failed")
    if not trap and not variant:
        # ptxas resource usage of the objects rebuilt in this call (registers / spills per kernel)
        with open(os.path.join(ROOT, "build_ptxas.log"), "a" if not force else "w") as fh:
            fh.write("".join(r[1] for r in results))
    return lib


if __name__ == "__main__":
    _variant = sys.argv[sys.argv.index("--variant") + 1] if "--variant" in sys.argv else None
    _defs = [a[2:] for a in sys.argv if a.startswith("-D")]
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, trap="--trap-barriers" in sys.argv,
                variant=_variant, defines=_defs))
