"""Kernel times INSIDE the replayed whole-step CUDA graph (warm caches, back-to-back launches, boost clocks): the
situation the bench's `value` is measured in.  torch.profiler (CUPTI) attributes a duration to every kernel node.
usage: python scripts/profile_graph_step.py [workload] [steps]"""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import bench
from cgat_b200 import optim as coptim, batching, graphed

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2_train"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
wl = bench.WORKLOADS[name]
dev = torch.device("cuda:0")
torch.manual_seed(0)
model, kw = bench.build_net(wl)
model = model.to(dev)
pool = [batching.pad_batch(sb) for sb in bench.make_pool(wl, 0)]
dev_pool = [sb.to(dev) for sb in pool]
n_real = wl["crystals"]
tg = [bench.target_norm(sb, dev, n_real) for sb in pool]
train = wl["train"]
if train:
    opt = coptim.FlatAdamW(model, lr=bench.LR, weight_decay=bench.WD)
    runner = graphed.GraphedTrainStep(model, opt, coptim.l1_loss, opt.sync)
    run = lambda i: runner.step(dev_pool[i % len(pool)], tg[i % len(pool)])
else:
    runner = graphed.GraphedForward(model)
    run = lambda i: runner(dev_pool[i % len(pool)])
for i in range(len(pool) + 4):
    run(i)
torch.cuda.synchronize()
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record()
for i in range(steps): run(i)
t1.record(); torch.cuda.synchronize()
print(f"# {name}: {t0.elapsed_time(t1)/steps:.3f} ms/step unprofiled")
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for i in range(steps): run(i)
    torch.cuda.synchronize()
rows = collections.defaultdict(lambda: [0.0, 0])
first, last = None, None
for ev in prof.events():
    if ev.device_type.name != "CUDA": continue
    d = ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
    k = ev.name
    for pre in ("void ", "cgat::(anonymous namespace)::", "at::native::", "(anonymous namespace)::"):
        k = k.replace(pre, "")
    rows[k[:90]][0] += d; rows[k[:90]][1] += 1
    s, e = ev.time_range.start, ev.time_range.end
    first = s if first is None else min(first, s); last = e if last is None else max(last, e)
tot = sum(v[0] for v in rows.values())
print(f"# kernels: {tot/steps/1e3:.3f} ms/step busy; span {(last-first)/steps/1e3:.3f} ms/step; {sum(v[1] for v in rows.values())//steps} launches/step")
print(f"# {'ms/step':>8} {'share':>6} {'n/step':>6} {'avg us':>8}  kernel")
own = 0.0
for k, (d, n) in sorted(rows.items(), key=lambda kv: -kv[1][0])[:60]:
    print(f"  {d/steps/1e3:8.3f} {100*d/tot:5.1f}% {n/steps:6.1f} {d/n:8.1f}  {k}")
    if "_kernel" in k and ("cgat" in k or not k.startswith(("cutlass", "reduce", "vectorized", "elementwise"))): pass
