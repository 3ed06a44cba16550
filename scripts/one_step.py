"""Scratch: N steady-state train (or inference) steps of a bench workload, eager, for ncu.  One warm-up step runs before
cudaProfilerStart, so `ncu --profile-from-start off` sees steady-state steps only (no optimizer-state initialisation)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from cgat_b200 import optim as coptim
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
    opt = coptim.FlatAdamW(model, lr=bench.LR, weight_decay=bench.WD)


def one(i):
    sb = pool[i % 2]
    if wl["train"]:
        out = model(sb.graph, sb.roost)
        coptim.l1_loss(out[:, :1], tg[i % 2]).backward()
        opt.sync.finish(); opt.step(); opt.zero_grad()
    else:
        with torch.no_grad():
            model(sb.graph, sb.roost)


one(0)
torch.cuda.synchronize()
torch.cuda.profiler.start()
for i in range(steps):
    one(i + 1)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
