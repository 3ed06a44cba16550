"""Scratch: torch.profiler breakdown of one cfg2 train step (which CUDA kernels dominate)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import bench
from cgat_b200 import distributed as cdist

wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg2_train"]
dev = torch.device("cuda:0")
torch.manual_seed(0)
model, kw = bench.build_net(wl)
model = model.to(dev)
pool = [sb.to(dev) for sb in bench.make_pool(wl, 0, 2)]
tg = [bench.target_norm(sb, dev) for sb in pool]
train = wl["train"]
if train:
    sync = cdist.GradSync(model, 1)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4, fused=True)
def step(i):
    sb = pool[i % 2]
    if not train:
        with torch.no_grad():
            return model(sb.graph, sb.roost)
    out = model(sb.graph, sb.roost)
    loss = (out[:, :1] - tg[i % 2]).abs().mean()
    loss.backward(); opt.step(); sync.zero_grad()
for i in range(3): step(i)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=False) as prof:
    for i in range(2): step(i)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70))
