"""Times cgat_gemm3x_nt against torch.matmul (fp32 SIMT and single-pass TF32) — scratch measurement."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cgat_b200 import _lib

def gemm(a, b, c):
    M, K = a.shape; N = b.shape[0]
    _lib.call("cgat_gemm3x_nt", a.data_ptr(), a.stride(0), b.data_ptr(), b.stride(0), None, c.data_ptr(), c.stride(0), M, N, K, 0, _lib.stream())

def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

for (M, N, K) in [(5504, 5120, 128), (5504, 16512, 128), (66048, 128, 256), (5504, 128, 5120), (8192, 8192, 512)]:
    a = torch.randn(M, K, device="cuda"); b = torch.randn(N, K, device="cuda"); c = torch.empty(M, N, device="cuda")
    t = timeit(lambda: gemm(a, b, c))
    torch.backends.cuda.matmul.allow_tf32 = False
    t32 = timeit(lambda: torch.matmul(a, b.t(), out=c))
    torch.backends.cuda.matmul.allow_tf32 = True
    ttf = timeit(lambda: torch.matmul(a, b.t(), out=c))
    torch.backends.cuda.matmul.allow_tf32 = False
    fl = 2.0 * M * N * K
    print(f"{M}x{N}x{K}: gemm3x {t:.3f} ms ({fl/t/1e9:.1f} TF fp32-equiv) | torch fp32 {t32:.3f} ms ({fl/t32/1e9:.1f} TF) | torch tf32 {ttf:.3f} ms ({fl/ttf/1e9:.1f} TF)")
