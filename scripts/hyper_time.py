"""Time the f16 hyper kernels alone (CUDA events, inputs > L2 not needed: weight stream 8 MB is L2-resident by design).
usage: python scripts/hyper_time.py [n_atoms] [f]   (library variant via CGAT_B200_LIB)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cgat_b200 import _lib, ops

n = int(sys.argv[1]) if len(sys.argv) > 1 else 5888
f = int(sys.argv[2]) if len(sys.argv) > 2 else 128
dev = "cuda"
torch.manual_seed(0)
w = torch.randn(f * f + f, f, device=dev) * 0.05
bias = torch.randn(f * f + f, device=dev) * 0.05
z = torch.randn(n, f, device=dev)
y = torch.randn(n, f, device=dev)
g = torch.randn(n, f, device=dev) * 1e-3
wp = ops.packed_kmajor(w, rows=f * f, f16=True)
lib = _lib.load()
parts = int(lib.cgat_hyper_rowscale_parts(n, f))
buf = torch.empty((parts, n, f), device=dev)
out = torch.empty(n, f, device=dev)
e = torch.zeros(n, f, device=dev)
def t(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
flops = 2.0 * n * f * f * f
ms = t(lambda: _lib.call("cgat_hyper_rowscale_f16", _lib.ptr(z), _lib.ptr(g), _lib.ptr(bias), _lib.ptr(wp), _lib.ptr(buf), n, f, _lib.stream()))
print(f"[{os.environ.get('CGAT_B200_LIB','')}] rowscale n={n} f={f}: {ms*1e3:.1f} us  {flops/ms/1e9:.1f} TFLOP/s")
ms = t(lambda: _lib.call("cgat_hyper_rowdot_fwd_f16", _lib.ptr(z), _lib.ptr(y), _lib.ptr(e), None, _lib.ptr(bias), _lib.ptr(wp), _lib.ptr(out), n, f, _lib.stream()))
print(f"[{os.environ.get('CGAT_B200_LIB','')}] rowdot   n={n} f={f}: {ms*1e3:.1f} us  {flops/ms/1e9:.1f} TFLOP/s")

splits = int(lib.cgat_hyper_wgrad_splits(n)) if hasattr(lib, "cgat_hyper_wgrad_splits") else 2
gam = g.abs().max().reshape(1).contiguous()
wout = torch.empty((splits, f * f, f), device=dev)
tail = torch.empty((splits, f, 2 * f), device=dev)
ms = t(lambda: _lib.call("cgat_hyper_wgrad_f16", _lib.ptr(g), _lib.ptr(y), _lib.ptr(z), _lib.ptr(gam), _lib.ptr(wout), _lib.ptr(tail), n, f, _lib.stream()))
print(f"[{os.environ.get('CGAT_B200_LIB','')}] wgrad    n={n} f={f}: {ms*1e3:.1f} us  {flops/ms/1e9:.1f} TFLOP/s  ({splits} splits)")
