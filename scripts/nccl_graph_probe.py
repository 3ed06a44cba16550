"""Probe (2+ GPUs): which way of overlapping the gradient all-reduce with a CUDA-graph-captured backward works here?
  mode A  NCCL all-reduce launched from an autograd hook INSIDE the capture (collective nodes in the graph)
  mode B  graph records EXTERNAL events from the hook; all-reduces run eagerly on a side stream that waits on them
Prints PROBE_<mode>_OK / mismatch.  Run under torchrun with a timeout."""
import faulthandler
import os
import sys

import torch
import torch.distributed as dist

faulthandler.dump_traceback_later(50, exit=True)
mode = sys.argv[1]
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(0)
model = torch.nn.Sequential(*[torch.nn.Linear(512, 512) for _ in range(6)]).to(dev)
params = list(model.parameters())
flat = torch.zeros(sum(p.numel() for p in params), device=dev)
views, off = [], 0
for p in params:
    views.append(flat[off:off + p.numel()].view_as(p))
    off += p.numel()
layer_of = {p: i // 2 for i, p in enumerate(params)}
ranges = []
o = 0
for l in range(6):
    n = sum(p.numel() for p in params[2 * l:2 * l + 2])
    ranges.append((o, o + n))
    o += n
x = torch.randn(64, 512, device=dev) * (rank + 1)

state = dict(pending=[2] * 6, works=[], capture_mode=None)
events = [torch.cuda.Event(external=True) for _ in range(6)]


def hook(p):
    l = layer_of[p]
    state["pending"][l] -= 1
    if state["pending"][l] == 0:
        torch._foreach_copy_([views[2 * l], views[2 * l + 1]], [params[2 * l].grad, params[2 * l + 1].grad])
        lo, hi = ranges[l]
        if mode == "A":
            state["works"].append(dist.all_reduce(flat[lo:hi], async_op=True))
        else:
            events[l].record(torch.cuda.current_stream())


for p in params:
    p.register_post_accumulate_grad_hook(hook)


def fwd_bwd():
    state["pending"] = [2] * 6
    for p in params:
        p.grad = None
    model(x).square().mean().backward()
    if mode == "A":
        for w in state["works"]:
            w.wait()
        state["works"] = []


def reference():
    for p in params:
        p.grad = None
    # plain eager: no hooks effect needed — recompute and all-reduce everything
    state["pending"] = [99] * 6
    model(x).square().mean().backward()
    g = torch.cat([p.grad.reshape(-1) for p in params])
    dist.all_reduce(g)
    return g


ref = reference()
print(f"[rank {rank}] reference done", flush=True)
# eager warm-up of the mode (establishes NCCL connections for these sizes)
comm = torch.cuda.Stream()
if mode == "A":
    fwd_bwd()
else:
    fwd_bwd()
    for l in range(5, -1, -1):
        lo, hi = ranges[l]
        dist.all_reduce(flat[lo:hi])
torch.cuda.synchronize()
print(f"[rank {rank}] eager mode step done, err {(flat - ref).abs().max().item():.3e}", flush=True)

g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    fwd_bwd()   # warm-up on the side stream as the docs ask
    if mode == "B":
        for l in range(5, -1, -1):
            lo, hi = ranges[l]
            dist.all_reduce(flat[lo:hi])
torch.cuda.current_stream().wait_stream(s)
torch.cuda.synchronize()
print(f"[rank {rank}] capturing", flush=True)
with torch.cuda.graph(g, capture_error_mode="thread_local"):
    fwd_bwd()
print(f"[rank {rank}] captured", flush=True)
for it in range(3):
    flat.zero_()
    g.replay()
    if mode == "B":
        cur = torch.cuda.current_stream()
        for l in range(5, -1, -1):          # backward finishes layer 5 first
            comm.wait_event(events[l])
            with torch.cuda.stream(comm):
                lo, hi = ranges[l]
                dist.all_reduce(flat[lo:hi])
        cur.wait_stream(comm)
    torch.cuda.synchronize()
    err = (flat - ref).abs().max().item()
    print(f"[rank {rank}] replay {it}: err {err:.3e}", flush=True)
    assert err < 1e-5 * ref.abs().max().item() + 1e-7, err
dist.barrier()
if rank == 0:
    print(f"PROBE_{mode}_OK", flush=True)
dist.destroy_process_group()
