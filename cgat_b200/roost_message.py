"""Roost composition branch of CGAtNet, parameter-compatible with the reference's
CGAT/roost_message.py (Roost, MessageLayer, WeightedAttention, WeightedMeanPooling, featuriser,
collate_batch keep their names and call signatures).

The soft-attention pooling of WeightedAttention (reference roost_message.py:302-317) runs on the
segmented-softmax kernel (cgat_seg_softmax_fwd/bwd): the reference's scatter_max + two scatter_add
calls become one deterministic pass over contiguous segments.
"""
from __future__ import annotations

import json

import numpy as np
import torch
import torch.nn as nn

from . import ops
from .graph import SegmentPlan, build_segment_plan
from .message_changed import SimpleNetwork, ResidualNetwork  # noqa: F401  (re-exported like the reference)


class Featuriser:
    """Element-symbol -> feature vector table (reference roost_message.py:33-57)."""

    def __init__(self, allowed_types):
        self.allowed_types = set(allowed_types)
        self._embedding = {}

    def get_fea(self, key):
        assert key in self.allowed_types, f"{key} is not an allowed atom type"
        return self._embedding[key]

    def load_state_dict(self, state_dict):
        self._embedding = state_dict
        self.allowed_types = set(state_dict.keys())

    def get_state_dict(self):
        return self._embedding

    def embedding_size(self):
        return len(next(iter(self._embedding.values())))


class LoadFeaturiser(Featuriser):
    """JSON dict of element -> list[float] (reference roost_message.py:60-84)."""

    def __init__(self, embedding_file):
        with open(embedding_file) as fh:
            table = json.load(fh)
        super().__init__(table.keys())
        self._embedding = {k: np.array(v, dtype=float) for k, v in table.items()}


class WeightedAttention(nn.Module):
    """alpha_t = w_t**pow * exp(g_t - segmax g) / (segsum(w**pow * exp(g - segmax)) + 1e-13);
    out_s = sum_t alpha_t * message(fea_t)        (reference roost_message.py:302-317).

    `index` must be sorted (it is by construction: reference data.py:90-96, roost_message.py:445-453).
    Unlike the reference, the output is sized by `size` (default index[-1]+1), so a trailing element
    without pairs does not shrink it."""

    EPS = 1e-13

    def __init__(self, gate_nn, message_nn, num_heads=1):
        super().__init__()
        self.gate_nn = gate_nn
        self.message_nn = message_nn
        self.pow = nn.Parameter(torch.randn(1))

    def forward(self, fea, index, weights, plan: SegmentPlan | None = None, size=None):
        if plan is None:
            size = int(index[-1]) + 1 if size is None else size
            plan = build_segment_plan(index, size)
        gate = self.gate_nn(fea)                                   # (n, 1)
        msg = self.message_nn(fea)                                 # (n, F)
        u = (weights ** self.pow).reshape(-1)                      # (n,)
        out = ops.seg_softmax(gate.view(-1, 1, 1), msg.unsqueeze(1), plan, u=u, eps=self.EPS)
        return out.squeeze(1)

    def __repr__(self):
        return f"{type(self).__name__}(gate_nn={self.gate_nn})"


class WeightedMeanPooling(nn.Module):
    """Weighted mean over segments (reference roost_message.py:267-280; unused by CGAtNet)."""

    def forward(self, fea, index, weights):
        size = int(index[-1]) + 1
        num = torch.zeros((size, fea.shape[1]), dtype=fea.dtype, device=fea.device).index_add_(0, index, weights * fea)
        cnt = torch.zeros(size, dtype=fea.dtype, device=fea.device).index_add_(
            0, index, torch.ones_like(index, dtype=fea.dtype)).clamp_(min=1)
        return num / cnt.unsqueeze(1)

    def __repr__(self):
        return type(self).__name__


class MessageLayer(nn.Module):
    """One Roost message-passing step over the complete digraph of a crystal's distinct elements
    (reference roost_message.py:88-157)."""

    def __init__(self, fea_len, num_heads=1):
        super().__init__()
        self.pooling = nn.ModuleList(
            WeightedAttention(gate_nn=SimpleNetwork(2 * fea_len, 1, [256]),
                              message_nn=SimpleNetwork(2 * fea_len, fea_len, [256]))
            for _ in range(num_heads))

    def forward(self, elem_weights, elem_in_fea, self_fea_idx, nbr_fea_idx, plan: SegmentPlan | None = None):
        if plan is None:
            plan = build_segment_plan(self_fea_idx, elem_in_fea.shape[0])
        nbr_w = elem_weights[nbr_fea_idx, :]
        pair = torch.cat([elem_in_fea[self_fea_idx, :], elem_in_fea[nbr_fea_idx, :]], dim=1)
        heads = [att(pair, self_fea_idx, nbr_w, plan) for att in self.pooling]
        pooled = heads[0] if len(heads) == 1 else torch.stack(heads).mean(dim=0)
        return pooled + elem_in_fea

    def __repr__(self):
        return type(self).__name__


class Roost(nn.Module):
    """Composition-only descriptor per crystal (reference roost_message.py:160-264)."""

    def __init__(self, orig_elem_fea_len, elem_fea_len, n_graph):
        super().__init__()
        self.embedding = nn.Linear(orig_elem_fea_len, elem_fea_len - 1)
        self.graphs = nn.ModuleList(MessageLayer(elem_fea_len, 1) for _ in range(n_graph))
        self.cry_pool = nn.ModuleList(
            [WeightedAttention(gate_nn=SimpleNetwork(elem_fea_len, 1, [256]), message_nn=nn.Identity())])

    def forward(self, elem_weights, orig_elem_fea, self_fea_idx, nbr_fea_idx, crystal_elem_idx, n_crystals=None):
        n_elem = orig_elem_fea.shape[0]
        if n_crystals is None:
            n_crystals = int(crystal_elem_idx[-1]) + 1
        pair_plan = build_segment_plan(self_fea_idx, n_elem)
        cry_plan = build_segment_plan(crystal_elem_idx, n_crystals)
        fea = torch.cat([self.embedding(orig_elem_fea), elem_weights], dim=1)
        for layer in self.graphs:
            fea = layer(elem_weights, fea, self_fea_idx, nbr_fea_idx, pair_plan)
        heads = [att(fea, crystal_elem_idx, elem_weights, cry_plan) for att in self.cry_pool]
        return heads[0] if len(heads) == 1 else torch.stack(heads).mean(dim=0)

    def __repr__(self):
        return type(self).__name__


def collate_batch(dataset_list):
    """Concatenate per-crystal Roost tuples with node offsets (reference roost_message.py:400-458).
    Returns (weights (Nc,1), fea (Nc,D), self_idx (Mc,), nbr_idx (Mc,), crystal_idx (Nc,))."""
    w, f, si, ni, ci = [], [], [], [], []
    base = 0
    for i, (weights, fea, self_idx, nbr_idx) in enumerate(dataset_list):
        n_i = fea.shape[0]
        w.append(weights)
        f.append(fea)
        si.append(self_idx + base)
        ni.append(nbr_idx + base)
        ci.append(torch.full((n_i,), i, dtype=torch.long))
        base += n_i
    return (torch.cat(w, dim=0).view(-1, 1), torch.cat(f, dim=0), torch.cat(si, dim=0), torch.cat(ni, dim=0),
            torch.cat(ci))
