"""Whole-step CUDA graphs for CGAtNet: one captured graph per shape bucket, replayed from static buffers.

A default-config train step is ~230 launches of this library's kernels plus ~700 small library kernels; enqueueing
them from Python takes about as long as the GPU needs to run them (profiles/: host enqueue 27 ms vs 31 ms of GPU
time at 500 crystals).  The reference has the same structure through PyG / Lightning and no answer to it.  Here a
step is captured ONCE per bucket signature (cgat_b200/batching.py pads ragged batches to bucket boundaries with a
dummy crystal) — forward, loss, backward and AdamW in one graph — and every later step of that bucket is three
things: copy the batch into the graph's static input buffers, `cudaGraphLaunch`, read the loss.  On more than one
rank the graph ends after the gradients have been packed into the flat all-reduce buffer; the single NCCL
all-reduce and the fused AdamW launch follow eagerly (two host calls per step).

Rules the capture relies on (all checked by tests/test_gpu_model.py::test_graphed_step_matches_eager):
  * nothing in CGAtNet.forward / backward syncs with the host or has data-dependent shapes (GraphBatch.num_graphs is
    a host int; all plans are built by kernels);
  * the C ABI never allocates, copies from host memory or synchronises (include/cgat_b200.h), so every entry point is
    capturable; pointer arrays are passed by value in kernel parameters;
  * packed weight images are rebuilt INSIDE the graph (ops.invalidate_packed before capture), because replays
    update the weights without Python noticing;
  * the optimizer is created with capturable=True and has run once eagerly (its state exists) before any capture;
  * gradients are graph-private (p.grad = None before capture), so no accumulate kernels exist and zero_grad is free.
"""
from __future__ import annotations

import torch

from . import _lib, batching, ops


class _Slot:
    """One captured graph + its static inputs / outputs."""
    __slots__ = ("graph", "inputs", "target", "out", "loss", "launches")


def _static_like(sb, dev):
    """Device buffers with the shapes of a (padded) batch; the graph reads its inputs from these."""
    return sb.to(dev)


def _copy_batch(dst, src):
    """src (pinned host or device) -> the static buffers of a slot; non_blocking so that the copies queue behind
    each other on the current stream and the host runs ahead."""
    for d, s in zip(dst.graph.tensors(), src.graph.tensors()):
        d.copy_(s, non_blocking=True)
    for d, s in zip(dst.roost, src.roost):
        d.copy_(s, non_blocking=True)


class GraphedTrainStep:
    """step(padded_batch, target) -> loss (0-dim device tensor, valid until the next step of the same bucket).

    `loss_fn(pred_real, target_real)` sees the rows of the real crystals only.  `sync` is a distributed.GradSync (its
    all-reduce is captured with the rest).  The first `eager_steps` calls run eagerly (lazy library initialisation,
    optimizer state), later calls capture on the first use of a bucket and replay afterwards."""

    def __init__(self, model, optimizer, loss_fn, sync=None, eager_steps=1, graph_collectives=True):
        for grp in optimizer.param_groups:
            if not grp.get("capturable", False):
                raise ValueError("GraphedTrainStep needs an optimizer created with capturable=True")
        self.model, self.opt, self.loss_fn = model, optimizer, loss_fn
        self.sync = sync if sync is not None else getattr(optimizer, "sync", None)
        # several ranks: True = the bucketed NCCL all-reduces (launched from autograd hooks, overlapping backward) and
        # the optimizer are captured with the rest; False = round-1 behaviour, the graph ends after the gradients are
        # packed and one blocking all-reduce + the optimizer follow eagerly after every replay
        self.graph_collectives = graph_collectives
        self.eager_left = max(1, int(eager_steps))
        self.slots = {}
        self.pool = None
        self.replayed_launches = 0      # kernels of this library executed through graph replays
        self.captures = 0

    # -- the step itself, as Python: used eagerly and under capture
    def _multi(self):
        return self.sync is not None and self.sync.world > 1

    def _split(self):
        """Several ranks with the collectives kept out of the graph."""
        return self._multi() and not self.graph_collectives

    def _fwd_bwd(self, sb, target):
        out = self.model(sb.graph, sb.roost)
        n_real = sb.graph.num_graphs - 1
        loss = self.loss_fn(out[:n_real, :1], target[:n_real])
        if self._split():
            world, self.sync.world = self.sync.world, 1      # hooks copy the buckets but do not start collectives
            loss.backward()
            self.sync.pack()             # live gradients -> the flat all-reduce buffer; p.grad = views into it
            self.sync.world = world
        else:
            loss.backward()              # hooks pack each finished bucket and start its all-reduce at once
        return out, loss

    def _finish(self):
        if self.sync is not None:
            if self._split():
                self.sync.reduce()
            else:
                self.sync.finish()       # wait for the bucket all-reduces; a no-op copy flush on one rank
        self.opt.step()

    def _body(self, sb, target):
        out, loss = self._fwd_bwd(sb, target)
        self._finish()
        return out, loss

    def _drop_grads(self):
        if self.sync is not None:
            self.sync.zero_grad()
        for p in self.model.parameters():
            p.grad = None

    def step(self, sb, target):
        """sb: batching.pad_batch output (pinned host or device tensors); target: (C+1, 1) normalised targets."""
        dev = next(self.model.parameters()).device
        if self.eager_left > 0:
            self.eager_left -= 1
            self._drop_grads()
            out, loss = self._body(sb.to(dev, non_blocking=True), target.to(dev, non_blocking=True))
            return loss.detach()
        sig = batching.signature(sb)
        slot = self.slots.get(sig)
        if slot is None:
            slot = self._capture(sig, sb, target, dev)
        _copy_batch(slot.inputs, sb)
        slot.target.copy_(target, non_blocking=True)
        slot.graph.replay()
        self.replayed_launches += slot.launches
        if self._split():
            self._finish()               # NCCL all-reduce + optimizer outside the graph
        ops.invalidate_packed()          # the step changed the weights behind Python's back
        return slot.loss

    def _capture(self, sig, sb, target, dev):
        slot = _Slot()
        slot.inputs = _static_like(sb, dev)
        slot.target = target.to(dev).clone()
        torch.cuda.synchronize()
        self._drop_grads()
        ops.invalidate_packed()          # the graph must contain its own pack launches
        slot.graph = torch.cuda.CUDAGraph()
        before = _lib.launch_count()
        # several ranks: NCCL's watchdog thread polls CUDA events while we capture; thread-local mode keeps its calls
        # from invalidating the capture (launches from the autograd thread into the capturing stream are still recorded)
        mode = "thread_local" if self._multi() else "global"
        with torch.cuda.graph(slot.graph, pool=self.pool, capture_error_mode=mode):
            # the whole step (several ranks: incl. the bucketed NCCL all-reduces, which overlap backward as parallel
            # branches of the graph); with graph_collectives=False: forward + backward + gradient packing only
            out, loss = self._fwd_bwd(slot.inputs, slot.target) if self._split() else self._body(slot.inputs, slot.target)
            slot.out, slot.loss = out.detach(), loss.detach()
        slot.launches = _lib.launch_count() - before
        if self.pool is None:
            self.pool = slot.graph.pool()
        self.slots[sig] = slot
        self.captures += 1
        return slot


class GraphedForward:
    """forward(padded_batch) -> predictions of the real crystals (C, out) as a view of the graph's static output
    (valid until the next call on the same bucket).  Inference: weights are static, packed images are built by the
    eager warm-up call and reused by every graph."""

    def __init__(self, model, eager_calls=1):
        self.model = model
        self.eager_calls = max(1, int(eager_calls))
        self.eager_left = self.eager_calls
        self.slots = {}
        self.pool = None
        self.replayed_launches = 0
        self.captures = 0
        self.epoch = ops._pack_epoch

    @torch.no_grad()
    def __call__(self, sb):
        dev = next(self.model.parameters()).device
        n_real = sb.graph.num_graphs - 1
        if ops._pack_epoch != self.epoch:
            # the weights changed since the graphs were captured (an optimizer step or a training-graph replay): the
            # captured launches read packed operand images that have been replaced — drop them and start over
            self.slots, self.eager_left, self.epoch = {}, self.eager_calls, ops._pack_epoch
        if self.eager_left > 0:
            self.eager_left -= 1
            d = sb.to(dev, non_blocking=True)
            return self.model(d.graph, d.roost)[:n_real]
        sig = batching.signature(sb)
        slot = self.slots.get(sig)
        if slot is None:
            slot = _Slot()
            slot.inputs = _static_like(sb, dev)
            torch.cuda.synchronize()
            slot.graph = torch.cuda.CUDAGraph()
            before = _lib.launch_count()
            with torch.cuda.graph(slot.graph, pool=self.pool):
                slot.out = self.model(slot.inputs.graph, slot.inputs.roost)
            slot.launches = _lib.launch_count() - before
            if self.pool is None:
                self.pool = slot.graph.pool()
            self.slots[sig] = slot
            self.captures += 1
        _copy_batch(slot.inputs, sb)
        slot.graph.replay()
        self.replayed_launches += slot.launches
        return slot.out[:n_real]
