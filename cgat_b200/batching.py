"""Shape bucketing of collated batches, so that a whole train / inference step can be replayed as a CUDA graph.

The reference collates ragged batches (reference CGAT/lightning_module.py:199-202, CGAT/roost_message.py:400-458):
atoms N, Roost elements Nc and Roost pairs Mc differ from batch to batch, and a CUDA graph bakes sizes in.
`pad_batch` appends ONE dummy crystal that absorbs the difference to the next bucket boundary:

  * dummy atoms: zero features, `max_nbr` self-loop edges each (rank 1), crystal id C;
  * dummy Roost elements: zero features, equal weights, crystal id C; dummy pairs among them (sorted `self_idx`);
  * target 0 for the dummy crystal; the training loss is taken over the first C rows only.

Crystals are independent units (no edge, softmax segment, Roost graph or pool crosses a crystal),
so the real crystals' predictions are unchanged, and because the dummy crystal's prediction carries zero loss
weight and all of its activations are finite, every parameter gradient is unchanged too
(tests/test_host_logic.py::test_padding_is_invisible, tests/test_gpu_model.py::test_graphed_step_matches_eager).
A few buckets then cover every batch of a training run, one captured graph per bucket (cgat_b200/graphed.py).
"""
from __future__ import annotations

import numpy as np
import torch

from .synthetic import GraphBatch, SyntheticBatch

# bucket granularity (atoms, Roost elements, Roost pairs): at the default batch of 500 crystals (N ~ 5.5 k,
# Nc ~ 1.75 k, Mc ~ 4.7 k) this pads < 3 % and a handful of buckets cover the batch-to-batch spread
DEFAULT_BUCKETS = (256, 128, 512)


def _round_up(n, m):
    return (n + m - 1) // m * m


def padded_sizes(n_atoms, n_elems, n_pairs, buckets=DEFAULT_BUCKETS):
    """(N_pad, Nc_pad, Mc_pad): at least one dummy atom and one dummy element are always added, so dummy edges and
    dummy pairs have something to attach to."""
    return (_round_up(n_atoms + 1, buckets[0]), _round_up(n_elems + 1, buckets[1]), _round_up(n_pairs, buckets[2]))


def pad_batch(sb: SyntheticBatch, buckets=DEFAULT_BUCKETS) -> SyntheticBatch:
    """A batch with one extra (dummy) crystal and bucketed sizes; `num_graphs` = C + 1.  Host (CPU) tensors in,
    host tensors out — this is collation-time work."""
    g = sb.graph
    weights, fea, self_idx, nbr_idx, cry_idx = sb.roost
    N, E = g.x.shape[0], g.edge_index.shape[1]
    if N == 0 or E % N:
        raise ValueError("pad_batch needs the reference layout: every atom has the same number of edges")
    K = E // N
    C = g.num_graphs
    Nc, Mc = fea.shape[0], self_idx.shape[0]
    n_pad, nc_pad, mc_pad = padded_sizes(N, Nc, Mc, buckets)
    a, a2, b = n_pad - N, nc_pad - Nc, mc_pad - Mc

    dummy_atoms = torch.arange(N, n_pad, dtype=torch.int64)
    loops = dummy_atoms.repeat_interleave(K)
    x = torch.cat([g.x, g.x.new_zeros((a, g.x.shape[1]))])
    edge_index = torch.cat([g.edge_index, torch.stack([loops, loops])], dim=1)
    edge_attr = torch.cat([g.edge_attr, torch.ones(a * K, dtype=torch.int64)])
    batch = torch.cat([g.batch, torch.full((a,), C, dtype=torch.int64)])
    y = None if g.y is None else torch.cat([g.y, g.y.new_zeros(1)])
    gp = GraphBatch(x, edge_index, edge_attr, batch, y, num_graphs=C + 1)

    w_p = torch.cat([weights, weights.new_full((a2, 1), 1.0 / a2)])
    f_p = torch.cat([fea, fea.new_zeros((a2, fea.shape[1]))])
    pair = torch.arange(b, dtype=torch.int64)
    s_p = torch.cat([self_idx, Nc + (pair * a2) // max(b, 1)])          # sorted, spread over the dummy elements
    n_p = torch.cat([nbr_idx, Nc + (pair + 1) % a2])
    c_p = torch.cat([cry_idx, torch.full((a2,), C, dtype=torch.int64)])
    n_atoms = np.concatenate([sb.n_atoms, np.array([a], dtype=sb.n_atoms.dtype)])
    return SyntheticBatch(gp, (w_p, f_p, s_p, n_p, c_p), n_atoms)


def signature(sb: SyntheticBatch):
    """What a captured graph is keyed on: every size a kernel launch bakes in."""
    g = sb.graph
    return (g.x.shape[0], g.edge_index.shape[1], g.num_graphs, sb.roost[1].shape[0], sb.roost[2].shape[0])


def collate_graphs(graphs) -> GraphBatch:
    """What torch_geometric's `Batch.from_data_list` does for the attributes CGAtNet.forward reads (reference
    CGAT/lightning_module.py:199-200, CGAT/data.py:139-144): concatenate x / edge_attr / y, shift every crystal's
    edge_index by its node offset, and emit the sorted `batch` vector.  `graphs`: objects with x (n,D), edge_index
    (2,e) int64, edge_attr (e,) int64 and optionally y."""
    graphs = list(graphs)
    sizes = torch.tensor([g.x.shape[0] for g in graphs], dtype=torch.int64)
    offsets = torch.cumsum(sizes, 0) - sizes
    x = torch.cat([g.x for g in graphs])
    edge_index = torch.cat([g.edge_index + off for g, off in zip(graphs, offsets.tolist())], dim=1)
    edge_attr = torch.cat([g.edge_attr for g in graphs])
    batch = torch.repeat_interleave(torch.arange(len(graphs), dtype=torch.int64), sizes)
    ys = [getattr(g, "y", None) for g in graphs]
    y = None if any(v is None for v in ys) else torch.cat([v.reshape(-1) for v in ys])
    return GraphBatch(x, edge_index, edge_attr, batch, y, num_graphs=len(graphs))


def collate(samples) -> SyntheticBatch:
    """Per-crystal samples -> one collated batch in the layout CGAtNet.forward takes.  `samples`: iterable of
    (graph, (weights, fea, self_idx, nbr_idx)) as the reference's CompositionData.__getitem__ yields them
    (CGAT/data.py:139-144); the Roost part goes through roost_message.collate_batch."""
    from .roost_message import collate_batch
    samples = list(samples)
    graph = collate_graphs(g for g, _ in samples)
    roost = collate_batch([r for _, r in samples])
    n_atoms = np.array([g.x.shape[0] for g, _ in samples], dtype=np.int64)
    return SyntheticBatch(graph, roost, n_atoms)
