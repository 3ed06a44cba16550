"""Scratch: N train (or inference) steps of a bench workload, for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from cgat_b200 import distributed as cdist
name = sys.argv[1] if len(sys.argv) > 1 else "cfg2_train"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
wl = bench.WORKLOADS[name]
dev = torch.device("cuda:0")
torch.manual_seed(0)
model, kw = bench.build_net(wl)
model = model.to(dev)
pool = [sb.to(dev) for sb in bench.make_pool(wl, 0, 2)]
tg = [bench.target_norm(sb, dev) for sb in pool]
if wl["train"]:
    sync = cdist.GradSync(model, 1)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4, fused=True)
for i in range(steps):
    sb = pool[i % 2]
    if wl["train"]:
        out = model(sb.graph, sb.roost)
        (out[:, :1] - tg[i % 2]).abs().mean().backward()
        opt.step(); sync.zero_grad()
    else:
        with torch.no_grad():
            model(sb.graph, sb.roost)
torch.cuda.synchronize()
print("done")
