"""Scratch: host-side cost of a train step (cProfile over 4 steps after warm-up; GPU work stays asynchronous)."""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from cgat_b200 import distributed as cdist

wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg2_train"]
dev = torch.device("cuda:0")
torch.manual_seed(0)
model, kw = bench.build_net(wl)
model = model.to(dev)
pool = [sb.to(dev) for sb in bench.make_pool(wl, 0, 2)]
tg = [bench.target_norm(sb, dev) for sb in pool]
sync = cdist.GradSync(model, 1)
opt = torch.optim.AdamW(model.parameters(), lr=1e-4, fused=True)
def step(i):
    sb = pool[i % 2]
    out = model(sb.graph, sb.roost)
    loss = (out[:, :1] - tg[i % 2]).abs().mean()
    loss.backward(); opt.step(); sync.zero_grad()
for i in range(4): step(i)
torch.cuda.synchronize()
t = time.perf_counter()
for i in range(4): step(i)
t_cpu = time.perf_counter() - t
torch.cuda.synchronize()
t_all = time.perf_counter() - t
print(f"4 steps: host enqueue {t_cpu*250:.2f} ms/step, with sync {t_all*250:.2f} ms/step")
pr = cProfile.Profile()
pr.enable()
for i in range(4): step(i)
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(35)
