"""Per-CTA timeline of one hyper f16 launch (needs the dbg64 library variant: CGAT_B200_LIB=dbg64)."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cgat_b200 import _lib, ops
n = int(sys.argv[1]) if len(sys.argv) > 1 else 5632
f = 128
dev = "cuda"; torch.manual_seed(0)
w = torch.randn(f * f + f, f, device=dev) * 0.05
bias = torch.randn(f * f + f, device=dev) * 0.05
z = torch.randn(n, f, device=dev); y = torch.randn(n, f, device=dev); g = torch.randn(n, f, device=dev) * 1e-3
wp = ops.packed_kmajor(w, rows=f * f, f16=True)
lib = _lib.load()
parts = int(lib.cgat_hyper_rowscale_parts(n, f))
buf = torch.empty((parts, n, f), device=dev); out = torch.empty(n, f, device=dev); e = torch.zeros(n, f, device=dev)
tl = np.zeros((160, 16), dtype=np.int64)
lib.cgat_debug_hyper_timeline.argtypes = [ctypes.c_void_p]
def show(name, fn):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); fn(); b.record(); torch.cuda.synchronize()
    lib.cgat_debug_hyper_timeline(tl.ctypes.data)
    t = tl[tl[:, 1] > 0]
    ghz = 1.965
    rel = lambda s: (t[:, s] - t[:, 0]) / ghz / 1e3   # us since CTA entry
    print(f"== {name} n={n}: event {a.elapsed_time(b)*1e3:.1f} us; CTAs {len(t)}; globaltimer span {(t[:,15].max()-t[:,1].min())/1e3:.1f} us, "
          f"start spread {(t[:,1].max()-t[:,1].min())/1e3:.2f} us")
    for s, lab in [(2, "prologue done"), (3, "first z tile staged"), (4, "first weight stage full (MMA issues)"),
                   (5, "first accumulator ready"), (6, "last accumulator drained"), (7, "tid0 before final sync"),
                   (8, "after final syncthreads"), (9, "exit")]:
        r = rel(s); print(f"  {lab:40s} min {r.min():7.2f}  med {np.median(r):7.2f}  max {r.max():7.2f} us")
    oi = t[:, 11]; print(f"  o-items per CTA min {oi.min()} med {np.median(oi)} max {oi.max()}; items {t[:,10].min()}..{t[:,10].max()}")
    busy = (t[:, 6] - t[:, 5]) / ghz / 1e3
    print(f"  steady state (first ready -> last drained) per o-item: med {np.median(busy / np.maximum(oi - 1, 1)):.3f} us = {np.median(busy / np.maximum(oi - 1, 1))*ghz*1e3:.0f} clk")
    m = t[:, 12] > 0
    if m.any():
        w8 = (t[m, 12] - t[m, 14]) / ghz / 1e3; st = (t[m, 13] - t[m, 12]) / ghz / 1e3
        print(f"  restage: {m.sum()} CTAs; wait for a_free med {np.median(w8):.2f} us, staging med {np.median(st):.2f} us")
show("rowscale", lambda: _lib.call("cgat_hyper_rowscale_f16", _lib.ptr(z), _lib.ptr(g), _lib.ptr(bias), _lib.ptr(wp), _lib.ptr(buf), n, f, _lib.stream()))
show("rowdot", lambda: _lib.call("cgat_hyper_rowdot_fwd_f16", _lib.ptr(z), _lib.ptr(y), _lib.ptr(e), None, _lib.ptr(bias), _lib.ptr(wp), _lib.ptr(out), n, f, _lib.stream()))
