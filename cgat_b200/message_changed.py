"""Dense blocks of CGAtNet, parameter-compatible with the reference's CGAT/message_changed.py
(same attribute names, shapes and default initialisers, so reference state_dicts load strictly).

  SimpleNetwork   — LeakyReLU(0.01) MLP                 (reference message_changed.py:36-66)
  Rezero          — learnable residual gate, init 0     (reference message_changed.py:69-78)
  ResidualNetwork — ReLU MLP with (optionally ReZero-gated) skip connections
                                                        (reference message_changed.py:81-138)
These are plain GEMM + epilogue chains over C (crystals) or K+1 (shell ranks) rows — 0.2 % of the
reference's forward time — and run on the library GEMM path.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops

LEAKY_SLOPE = 0.01  # nn.LeakyReLU() default; the reference never passes its negative_slope argument


class SimpleNetwork(nn.Module):
    def __init__(self, input_dim, output_dim, hidden_layer_dims):
        super().__init__()
        widths = [input_dim, *hidden_layer_dims]
        self.fcs = nn.ModuleList(nn.Linear(a, b) for a, b in zip(widths[:-1], widths[1:]))
        self.acts = nn.ModuleList(nn.LeakyReLU() for _ in widths[1:])  # kept for module-tree parity
        self.fc_out = nn.Linear(widths[-1], output_dim)

    def forward(self, fea):
        if fea.dim() != 2:
            for layer in self.fcs:
                fea = F.leaky_relu(layer(fea), LEAKY_SLOPE)
            return self.fc_out(fea)
        for layer in self.fcs:
            fea = ops.linear_act(fea, layer.weight, layer.bias, 1)
        return ops.linear_act(fea, self.fc_out.weight, self.fc_out.bias, 0)

    def __repr__(self):
        return type(self).__name__


class Rezero(nn.Module):
    def __init__(self):
        super().__init__()
        self.alpha = nn.Parameter(torch.zeros(1))

    def forward(self, x):
        return self.alpha * x

    def __repr__(self):
        return type(self).__name__


class ResidualNetwork(nn.Module):
    def __init__(self, input_dim, output_dim, hidden_layer_dims, if_rezero=False):
        super().__init__()
        widths = [input_dim, *hidden_layer_dims]
        pairs = list(zip(widths[:-1], widths[1:]))
        self.fcs = nn.ModuleList(nn.Linear(a, b) for a, b in pairs)
        self.res_fcs = nn.ModuleList(nn.Linear(a, b, bias=False) if a != b else nn.Identity() for a, b in pairs)
        self.acts = nn.ModuleList(nn.ReLU() for _ in pairs)
        self.fc_out = nn.Linear(widths[-1], output_dim)
        self.if_rezero = if_rezero
        if if_rezero:
            self.rezeros = nn.ModuleList(Rezero() for _ in pairs)

    def forward(self, fea, *, last_layer=True):
        for i, (fc, skip) in enumerate(zip(self.fcs, self.res_fcs)):
            branch = ops.linear_act(fea, fc.weight, fc.bias, 3) if fea.dim() == 2 else F.relu(fc(fea))
            if self.if_rezero:
                branch = self.rezeros[i](branch)
            fea = branch + skip(fea)
        return self.fc_out(fea) if last_layer else fea

    def __repr__(self):
        return type(self).__name__
