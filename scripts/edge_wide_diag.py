"""Diagnostic: relative errors of the fused edge attention (forward + every gradient) against fp64, next to the
library-fp32 formulation on the same device.  python scripts/edge_wide_diag.py F heads n_cry k"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cgat_b200 import graph, ops, synthetic
from cgat_b200.CGAT import MultiHeadNetwork
from oracle import cgat_oracle as O
f, heads, n_cry, k = (int(a) for a in sys.argv[1:5])
DEV = "cuda:0"
fe = 128
sb = synthetic.make_batch(n_cry, k, seed=100 + n_cry)
gidx = sb.graph
n = gidx.num_nodes
torch.manual_seed(n_cry)
width = 2 * f + fe
mh_a = MultiHeadNetwork(width, f, int(width / 1.5), heads)
mh_m = MultiHeadNetwork(width, f, int(width / 1.5), heads)
x = torch.randn(n, f) * 0.5
tab = torch.randn(k + 1, fe)
w = torch.randn(n, f, generator=torch.Generator().manual_seed(5))
sd = {}
for pre, mod in (("A.", mh_a), ("M.", mh_m)):
    for kk, v in mod.state_dict().items():
        sd[pre + kk] = v.double().requires_grad_(True)
xd, tabd = x.double().requires_grad_(True), tab.double().requires_grad_(True)
src, dst = gidx.edge_index
m = torch.cat([xd[dst], tabd[gidx.edge_attr], xd[src]], dim=1)
alpha = O.pyg_softmax(O.multi_head_network(sd, "A.", m, heads), dst, n)
ref = O.seg_sum(O.multi_head_network(sd, "M.", m, heads) * alpha, dst, n).mean(dim=1)
(ref * w.double()).sum().backward()
plan = graph.build_edge_plan(gidx.edge_index.to(DEV), gidx.edge_attr.to(DEV), n)
mh_a, mh_m = mh_a.to(DEV), mh_m.to(DEV)


def run(fused):
    ops._FUSED = fused
    for p in list(mh_a.parameters()) + list(mh_m.parameters()):
        p.grad = None
    xc, tabc = x.to(DEV).requires_grad_(True), tab.to(DEV).requires_grad_(True)
    out = ops.edge_attention(xc, tabc, plan, mh_a, mh_m, heads)
    (out * w.to(DEV)).sum().backward()
    res = {"out": (out.detach(), ref.detach()), "d_x": (xc.grad, xd.grad), "d_tab": (tabc.grad, tabd.grad)}
    for pre, mod in (("A.", mh_a), ("M.", mh_m)):
        for kk, p in mod.named_parameters():
            res["d_" + pre + kk] = (p.grad.clone(), sd[pre + kk].grad)
    return res


a, b = run(True), run(False)
print(f"F={f} heads={heads} n={n} E={gidx.edge_index.shape[1]} hd={int(width / 1.5)}")
for key in a:
    ea = (a[key][0].double().cpu() - a[key][1]).abs()
    eb = (b[key][0].double().cpu() - b[key][1]).abs()
    r = a[key][1]
    tol = 1e-4 + 1e-3 * r.abs()
    print(f"{key:22s} fused: max {ea.max():.2e} relL2 {ea.norm() / r.norm():.2e} bad {(ea > tol).sum().item():6d} | "
          f"library: max {eb.max():.2e} relL2 {eb.norm() / r.norm():.2e} bad {(eb > tol).sum().item():6d} | ref max {r.abs().max():.2e}")
