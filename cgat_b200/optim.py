"""Train-step glue on this library's own kernels: AdamW over flat buffers and the L1 loss.

The reference builds `torch.optim.AdamW(lr=1.25e-4, weight_decay=1e-6)` and `nn.L1Loss` in its Lightning module
(reference CGAT/lightning_module.py:328-344, :131-142, :237-240).  Here the live parameters are re-homed into ONE
flat fp32 buffer laid out exactly like distributed.GradSync's gradient buffer, so an optimizer step is a single
HBM-bound launch (cgat_adamw_flat: 16 B read + 12 B written per parameter) that reads the all-reduced gradients in
place — with the data-parallel average folded into the read — instead of a multi-tensor-apply over 331 tensors.
Learning rate and step count live on the device: a captured CUDA graph replays with their current values, and a
schedule (the reference's cyclical LR, CGAT/utils.py) just writes the scalar before the step.
"""
from __future__ import annotations

import torch

from . import _lib, ops
from .distributed import GradSync


class FlatAdamW:
    """AdamW (decoupled weight decay, torch.optim.AdamW arithmetic) for a CGAtNet.

    opt = FlatAdamW(model, lr=1.25e-4, weight_decay=1e-6, world_size=W)   # after model.to(device)
    loss.backward(); opt.sync.finish(); opt.step(); opt.zero_grad()

    Parameters that never receive gradients are left alone, like torch.optim.AdamW leaves
    parameters whose .grad is None.  The parameters keep their names / shapes (state_dict is unchanged); only their
    storage moves into `flat_p`."""

    def __init__(self, model, lr=1.25e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-6, world_size=1,
                 process_group=None, sync=None):
        self.sync = sync if sync is not None else GradSync(model, world_size, process_group, overlap=True, average=False)
        if self.sync.average:
            raise ValueError("FlatAdamW folds the 1/world average into its gradient read: pass GradSync(average=False)")
        s = self.sync
        dev = s.flat.device
        if dev.type != "cuda":
            raise _lib.CgatLibraryError("FlatAdamW runs cgat_adamw_flat on the GPU (there is no CPU path)")
        self.flat_p = torch.zeros_like(s.flat)
        with torch.no_grad():
            for p in s.params:
                view = self.flat_p[s.offsets[p]: s.offsets[p] + p.numel()].view_as(p)
                view.copy_(p.data)
                p.data = view
        self.exp_avg = torch.zeros_like(s.flat)
        self.exp_avg_sq = torch.zeros_like(s.flat)
        self.step_t = torch.zeros(1, dtype=torch.float32, device=dev)
        self.lr_t = torch.full((1,), float(lr), dtype=torch.float32, device=dev)
        self.betas, self.eps, self.weight_decay = betas, eps, weight_decay
        # torch.optim-shaped view of the hyper-parameters (graphed.GraphedTrainStep checks `capturable`)
        self.param_groups = [dict(params=s.params, lr=float(lr), betas=betas, eps=eps, weight_decay=weight_decay,
                                  capturable=True)]
        ops.invalidate_packed()   # every weight moved: packed operand caches are keyed on data_ptr

    def set_lr(self, lr):
        """New learning rate for the following steps (also for replays of an already captured graph)."""
        self.param_groups[0]["lr"] = float(lr)
        self.lr_t.fill_(float(lr))

    def step(self):
        s = self.sync
        _lib.call("cgat_adamw_flat", _lib.ptr(self.flat_p), _lib.ptr(s.flat), _lib.ptr(self.exp_avg),
                  _lib.ptr(self.exp_avg_sq), s.total, _lib.ptr(self.lr_t), _lib.ptr(self.step_t), self.betas[0],
                  self.betas[1], self.eps, self.weight_decay, 1.0 / s.world, _lib.stream(),
                  work=dict(key="adamw_flat", bound="hbm", bytes=28.0 * s.total,
                            note="16 B read + 12 B written per parameter"))
        ops.invalidate_packed()

    def zero_grad(self, set_to_none=True):
        self.sync.zero_grad()

    def state_dict(self):
        return dict(step=self.step_t.clone(), lr=self.lr_t.clone(), exp_avg=self.exp_avg.clone(),
                    exp_avg_sq=self.exp_avg_sq.clone())

    def load_state_dict(self, sd):
        self.step_t.copy_(sd["step"]); self.lr_t.copy_(sd["lr"])
        self.exp_avg.copy_(sd["exp_avg"]); self.exp_avg_sq.copy_(sd["exp_avg_sq"])


class _L1Loss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, out, target, n_real):
        if out.dim() != 2 or out.stride(1) != 1:
            raise ValueError("l1_loss: out must be (rows, cols) with contiguous rows")
        target = target.reshape(-1).contiguous()
        loss = torch.empty(1, dtype=torch.float32, device=out.device)
        grad = torch.empty(out.shape, dtype=torch.float32, device=out.device) if out.requires_grad else None
        _lib.call("cgat_l1_loss", out.data_ptr(), out.stride(0), _lib.ptr(target), n_real, _lib.ptr(loss),
                  _lib.ptr(grad), out.shape[1], out.shape[0], out.shape[1], _lib.stream(),
                  work=dict(key="l1_loss", bound="hbm", bytes=4.0 * out.shape[0] * (2 + out.shape[1])))
        ctx.save_for_backward(grad)
        return loss.view(())

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return grad * g, None, None


def l1_loss(out, target, n_real=None):
    """mean |out[:n_real, 0] - target[:n_real]| — nn.L1Loss on the first prediction column against the normalised
    target (reference CGAT/lightning_module.py:206-210, 237-240), loss and gradient from one launch; rows beyond
    n_real (the padding crystal of batching.pad_batch) get zero gradient."""
    n_real = out.shape[0] if n_real is None else int(n_real)
    return _L1Loss.apply(out, target, n_real)
