"""Scratch: ONE train step of a 1-layer CGAtNet at the cfg2 batch size (same per-kernel shapes as the 5-layer bench
model), so that an `ncu --set full` capture of every kernel type stays short."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from cgat_b200 import distributed as cdist
wl = dict(bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg2_train"])
wl["net"] = dict(wl["net"], n_graph=int(sys.argv[2]) if len(sys.argv) > 2 else 1)
dev = torch.device("cuda:0")
torch.manual_seed(0)
model, kw = bench.build_net(wl)
model = model.to(dev)
sb = bench.make_pool(wl, 0, 1)[0].to(dev)
tg = bench.target_norm(sb, dev)
if wl["train"]:
    out = model(sb.graph, sb.roost)
    (out[:, :1] - tg).abs().mean().backward()
else:
    with torch.no_grad():
        model(sb.graph, sb.roost)
torch.cuda.synchronize()
print("done")
