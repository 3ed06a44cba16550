"""world_size-2 gloo test (CPU) of the gradient exchange used for multi-GPU training."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cgat_b200 import distributed as cdist


class _Toy(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.graphs = torch.nn.ModuleList([torch.nn.ModuleDict({
            "Node": torch.nn.Linear(4, 4),
            "Edge": torch.nn.ModuleDict({"MH_A": torch.nn.Linear(2, 2), "Pooling_NN": torch.nn.Linear(3, 3)})})
            for _ in range(2)])
        self.out = torch.nn.Linear(4, 1)


def _worker(rank, world, port):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    model = _Toy()
    for overlap in (True, False):
        sync = cdist.GradSync(model, world, overlap=overlap)
        # bucket order = the order backward finishes them: [out] then graphs.1, graphs.0
        assert sync.names[0][0].startswith("out.") and sync.names[1][0].startswith("graphs.1.")
        assert all(off % 128 == 0 for off in sync.offsets.values())     # 512-byte aligned slices
        x = torch.full((3, 4), float(rank + 1))
        loss = model.out(model.graphs[0]["Node"](x)).sum() + model.graphs[0]["Edge"]["Pooling_NN"](torch.ones(3)).sum()
        loss.backward()          # overlap=True: hooks pack finished buckets and start their all-reduce (async)
        if not overlap:
            assert all(p.grad is None or p.grad.data_ptr() != sync.slice_of(p).data_ptr() for p in sync.params)
        # parameters that took no part in this step (graphs.1.Node here) enter the exchange as zeros
        local = {p: (p.grad.detach().clone() if p.grad is not None else torch.zeros_like(p)) for p in sync.params}
        if overlap:
            # what the hooks already exchanged must not be gathered again: recompute the local gradients
            model2 = _Toy()
            model2.load_state_dict(model.state_dict())
            l2 = (model2.out(model2.graphs[0]["Node"](x)).sum()
                  + model2.graphs[0]["Edge"]["Pooling_NN"](torch.ones(3)).sum())
            l2.backward()
            p2 = dict(model2.named_parameters())
            local = {p: (p2[n].grad if p2[n].grad is not None else torch.zeros_like(p))
                     for n, p in model.named_parameters() if p in sync.offsets}
        sync.all_reduce()
        for p in sync.params:
            gathered = [torch.zeros_like(local[p]) for _ in range(world)]
            dist.all_gather(gathered, local[p].contiguous())
            assert torch.allclose(sync.slice_of(p).view_as(p), sum(gathered) / world), overlap
            assert p.grad.data_ptr() == sync.slice_of(p).data_ptr()      # grads are views of the flat buffer
        sync.zero_grad()
        assert all(p.grad is None for p in sync.params)
        sync.remove_hooks()
    dist.destroy_process_group()


def test_grad_sync_gloo_world2():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port), nprocs=2, join=True)


def test_dead_parameters_excluded():
    names = [n for n, _ in cdist.live_parameters(_Toy())]
    assert not any(".Edge.MH_A." in n for n in names)
    assert not any(n.startswith("graphs.1.Edge.Pooling_NN.") for n in names)
    assert any(n.startswith("graphs.0.Edge.Pooling_NN.") for n in names)
