"""Scratch: relative L2 error of the fused hypernetwork-trunk kernels vs fp64, next to PyTorch's own fp32."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cgat_b200 import ops
from tests.test_gpu_kernels import _trunk_case

def run(n, n_j, dev, dt, fused):
    ops._TRUNK = fused
    f = 128
    h, layers, tails = _trunk_case(n, n_j, 100 + n)
    g = torch.Generator().manual_seed(n)
    gz = [torch.randn(n, f, generator=g) for _ in range(n_j)]
    ge = [torch.randn(n, f, generator=g) for _ in range(n_j)]
    hh = h.to(dev, dt).requires_grad_(True)
    ll = [[(w.to(dev, dt).requires_grad_(True), b.to(dev, dt).requires_grad_(True)) for w, b in l] for l in layers]
    tt = [(w.to(dev, dt), b.to(dev, dt), r) for w, b, r in tails]
    zs, es = ops.hyper_trunks(hh, ll, tt)
    if es[0] is None:
        es = [z @ w[r:].t() + b[r:] for z, (w, b, r) in zip(zs, tt)]
    loss = sum((z * a.to(dev, dt)).sum() + (e * c.to(dev, dt)).sum() for z, e, a, c in zip(zs, es, gz, ge))
    loss.backward()
    out = {"z0": zs[0], "z3": zs[-1], "e0": es[0], "g_h": hh.grad}
    for s in range(4):
        out[f"g_w{s}"] = ll[0][s][0].grad
        out[f"g_b{s}"] = ll[0][s][1].grad
    return {k: v.detach().double().cpu() for k, v in out.items()}

for n, n_j in [(700, 4)]:
    ref = run(n, n_j, "cpu", torch.float64, False)
    a = run(n, n_j, "cuda:0", torch.float32, True)
    b = run(n, n_j, "cuda:0", torch.float32, False)
    print(f"n={n}")
    for k in ref:
        ea = ((a[k] - ref[k]).norm() / ref[k].norm()).item()
        eb = ((b[k] - ref[k]).norm() / ref[k].norm()).item()
        sa = (((a[k] - ref[k]) * ref[k].sign()).mean() / ref[k].abs().mean()).item()
        sb = (((b[k] - ref[k]) * ref[k].sign()).mean() / ref[k].abs().mean()).item()
        print(f"   {k:6s} fused relL2 {ea:.2e} signed {sa:+.2e}   torch fp32 {eb:.2e} signed {sb:+.2e}")
