"""Per-kernel times of one fused edge-attention layer (forward + backward) at the bench size (500 crystals, 12 nbrs,
5 heads, F = 128, Hd = 256), eager, CUDA events around every C-ABI call.  Library variant via CGAT_B200_LIB."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cgat_b200 import _lib, graph, ops, synthetic
from cgat_b200.CGAT import MultiHeadNetwork
DEV = "cuda:0"
n_cry = int(sys.argv[1]) if len(sys.argv) > 1 else 500
reps = 5
f, fe, heads, k = 128, 128, 5, 12
sb = synthetic.make_batch(n_cry, k, seed=1)
gidx = sb.graph
n = gidx.num_nodes
torch.manual_seed(0)
width = 2 * f + fe
mh_a = MultiHeadNetwork(width, f, int(width / 1.5), heads).to(DEV)
mh_m = MultiHeadNetwork(width, f, int(width / 1.5), heads).to(DEV)
x = (torch.randn(n, f) * 0.5).to(DEV).requires_grad_(True)
tab = torch.randn(k + 1, fe).to(DEV).requires_grad_(True)
w = torch.randn(n, f).to(DEV)
plan = graph.build_edge_plan(gidx.edge_index.to(DEV), gidx.edge_attr.to(DEV), n)
def step():
    out = ops.edge_attention(x, tab, plan, mh_a, mh_m, heads)
    (out * w).sum().backward()
for _ in range(3): step()
torch.cuda.synchronize()
_lib.profile_begin()
for _ in range(reps): step()
rows = _lib.profile_end()
tag = os.environ.get("CGAT_B200_LIB", "")
print(f"[{tag}] atoms {n} edges {gidx.edge_index.shape[1]}")
for key, r in sorted(rows.items(), key=lambda kv: -kv[1]["ms"]):
    if r["ms"] / reps < 0.02: continue
    print(f"[{tag}]   {key:24s} {r['launches']//reps:3d} x {1e3*r['ms']/r['launches']:8.1f} us")
