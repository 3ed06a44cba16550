"""Seeded synthetic crystal batches in the reference's collated layout.

The layout is the boundary contract of `CGAtNet.forward(batch, roost)`:
  * graph part — what `Batch.from_data_list` yields from `CompositionData.__getitem__`
    (reference CGAT/data.py:139-144, CGAT/lightning_module.py:199-200):
      x (N,200) f32, edge_index (2,E) i64 with row 0 = source (each atom repeated K times),
      row 1 = neighbour inside the same crystal, edge_attr (E,) i64 shell rank in [1,K],
      batch (N,) i64 sorted crystal id, y (C,) f32.
  * Roost part — what `collate_batch` yields (reference CGAT/roost_message.py:400-458) from the
    per-crystal complete digraph over distinct elements (reference CGAT/data.py:81-103):
      weights (Nc,1) f32, fea (Nc,200) f32, self_idx (Mc,) i64, nbr_idx (Mc,) i64, crystal_idx (Nc,) i64.

Everything is generated with vectorised numpy so that screening-sized pools (cfg3) are cheap.
"""
from __future__ import annotations

import os
from dataclasses import dataclass

import numpy as np
import torch

_DATA = os.path.join(os.path.dirname(__file__), "data", "matscholar_f32.npz")
_cache = {}


def matscholar():
    """(elements[103], fea[103,200] f32) — the reference's only in-tree fixture
    (embeddings/matscholar-embedding.json), stored as float32."""
    if "m" not in _cache:
        z = np.load(_DATA)
        _cache["m"] = (z["elements"], z["fea"].astype(np.float32))
    return _cache["m"]


class GraphBatch:
    """Duck-type of torch_geometric's Batch: the attributes CGAtNet.forward reads
    (reference CGAT/CGAT.py:566-571)."""

    def __init__(self, x, edge_index, edge_attr, batch, y=None, num_graphs=None):
        self.x, self.edge_index, self.edge_attr, self.batch, self.y = x, edge_index, edge_attr, batch, y
        # a host-side int like torch_geometric's Batch.num_graphs (known at collation time): reading it from the
        # device tensor would be a device->host sync in every forward (the reference has one at CGAT/CGAT.py:52)
        if num_graphs is None:
            num_graphs = int(batch[-1]) + 1 if batch.numel() else 0
        self.num_graphs = num_graphs

    @property
    def num_nodes(self):
        return self.x.shape[0]

    def to(self, device, non_blocking=False):
        f = lambda t: None if t is None else t.to(device, non_blocking=non_blocking)
        return GraphBatch(f(self.x), f(self.edge_index), f(self.edge_attr), f(self.batch), f(self.y), self.num_graphs)

    def pin_memory(self):
        f = lambda t: None if t is None else t.pin_memory()
        return GraphBatch(f(self.x), f(self.edge_index), f(self.edge_attr), f(self.batch), f(self.y), self.num_graphs)

    def tensors(self):
        return [t for t in (self.x, self.edge_index, self.edge_attr, self.batch, self.y) if t is not None]


@dataclass
class SyntheticBatch:
    graph: GraphBatch
    roost: tuple  # (weights, fea, self_idx, nbr_idx, crystal_idx)
    n_atoms: np.ndarray  # per crystal

    @property
    def num_crystals(self):
        return len(self.n_atoms)

    def to(self, device, non_blocking=False):
        return SyntheticBatch(self.graph.to(device, non_blocking),
                              tuple(t.to(device, non_blocking=non_blocking) for t in self.roost), self.n_atoms)

    def pin_memory(self):
        return SyntheticBatch(self.graph.pin_memory(), tuple(t.pin_memory() for t in self.roost), self.n_atoms)

    def nbytes(self):
        return sum(t.numel() * t.element_size() for t in list(self.graph.tensors()) + list(self.roost))


def _ragged_arange(counts):
    """concatenate(arange(c) for c in counts) without a Python loop."""
    counts = np.asarray(counts, dtype=np.int64)
    total = int(counts.sum())
    starts = np.cumsum(counts) - counts
    return np.arange(total, dtype=np.int64) - np.repeat(starts, counts)


def make_batch(n_crystals, max_nbr=12, seed=0, atoms_lo=2, atoms_hi=20, dtype=torch.float32):
    """SURVEY.md §8d generator. cfg1-4: atoms U{2..20}; cfg5: atoms U{200..256}, max_nbr=24."""
    rng = np.random.default_rng(seed)
    elems, table = matscholar()
    C, K = int(n_crystals), int(max_nbr)
    n_c = rng.integers(atoms_lo, atoms_hi + 1, size=C).astype(np.int64)
    k_c = rng.integers(2, np.minimum(5, n_c) + 1).astype(np.int64)  # >=2 distinct elements always
    N = int(n_c.sum())
    atom_off = np.cumsum(n_c) - n_c
    crystal_of_atom = np.repeat(np.arange(C, dtype=np.int64), n_c)
    local = _ragged_arange(n_c)

    # distinct elements per crystal: first k_c entries of a random permutation of the 103 keys
    perm = np.argsort(rng.random((C, len(elems))), axis=1)[:, :5]
    k_atom = k_c[crystal_of_atom]
    slot = np.where(local < k_atom, local, (rng.random(N) * k_atom).astype(np.int64))
    slot = np.minimum(slot, k_atom - 1)
    atom_elem = perm[crystal_of_atom, slot]
    x = table[atom_elem]

    # neighbours: K i.i.d. uniform picks inside the crystal (self loops / duplicates allowed)
    E = N * K
    src = np.repeat(np.arange(N, dtype=np.int64), K)
    nbr_local = (rng.random(E) * np.repeat(n_c[crystal_of_atom], K)).astype(np.int64)
    nbr_local = np.minimum(nbr_local, np.repeat(n_c[crystal_of_atom], K) - 1)
    dst = np.repeat(atom_off[crystal_of_atom], K) + nbr_local
    # shell rank: starts at 1, increments with p=0.4 along the K slots (non-decreasing, <= K)
    inc = (rng.random((N, K)) < 0.4).astype(np.int64)
    inc[:, 0] = 0
    rank = (1 + np.cumsum(inc, axis=1)).reshape(-1)

    # Roost tuple exactly as reference CGAT/data.py:81-103 + collate_batch
    counts = np.zeros((C, 5), dtype=np.int64)
    np.add.at(counts, (crystal_of_atom, slot), 1)
    Nc = int(k_c.sum())
    r_cry = np.repeat(np.arange(C, dtype=np.int64), k_c)
    r_local = _ragged_arange(k_c)
    r_off = np.cumsum(k_c) - k_c
    weights = (counts[r_cry, r_local] / n_c[r_cry]).astype(np.float32).reshape(-1, 1)
    r_fea = table[perm[r_cry, r_local]]
    # complete digraph: for element i, neighbours = all j != i in increasing order
    deg = np.repeat(k_c - 1, k_c)
    self_idx = np.repeat(np.arange(Nc, dtype=np.int64), deg)
    j = _ragged_arange(deg)
    i_local = np.repeat(r_local, deg)
    j = j + (j >= i_local)
    nbr_idx = np.repeat(r_off[r_cry], deg) + j

    y = (rng.standard_normal(C) * n_c).astype(np.float32)

    g = GraphBatch(torch.from_numpy(x).to(dtype),
                   torch.from_numpy(np.stack([src, dst])),
                   torch.from_numpy(rank),
                   torch.from_numpy(crystal_of_atom),
                   torch.from_numpy(y))
    roost = (torch.from_numpy(weights).to(dtype), torch.from_numpy(r_fea).to(dtype),
             torch.from_numpy(self_idx), torch.from_numpy(nbr_idx), torch.from_numpy(r_cry))
    return SyntheticBatch(g, roost, n_c)


def split_batch(sb: SyntheticBatch, lo: int, hi: int) -> SyntheticBatch:
    """Crystals [lo,hi) of a batch as an independent batch (indices re-based). Used by the
    batch-split equivalence test and by the multi-GPU sharder: no edge, softmax segment or Roost
    graph crosses a crystal boundary (reference CGAT/data.py:140; roost_message.py:445-446)."""
    g, r = sb.graph, sb.roost
    n_c = sb.n_atoms
    a0, a1 = int(n_c[:lo].sum()), int(n_c[:hi].sum())
    K = g.edge_index.shape[1] // g.x.shape[0]
    e0, e1 = a0 * K, a1 * K
    gb = GraphBatch(g.x[a0:a1], g.edge_index[:, e0:e1] - a0, g.edge_attr[e0:e1], g.batch[a0:a1] - lo,
                    None if g.y is None else g.y[lo:hi])
    cidx = r[4]
    m = (cidx >= lo) & (cidx < hi)
    n0 = int((cidx < lo).sum())
    n1 = n0 + int(m.sum())
    em = (r[2] >= n0) & (r[2] < n1)
    roost = (r[0][n0:n1], r[1][n0:n1], r[2][em] - n0, r[3][em] - n0, cidx[n0:n1] - lo)
    return SyntheticBatch(gb, roost, n_c[lo:hi])


def shard_bounds(n_atoms, max_nbr, world):
    """Contiguous crystal ranges per rank balanced by edge count."""
    edges = np.asarray(n_atoms, dtype=np.int64) * max_nbr
    cum = np.cumsum(edges)
    total = int(cum[-1]) if len(cum) else 0
    cuts = [0]
    for r in range(1, world):
        cuts.append(int(np.searchsorted(cum, total * r / world, side="left")) + 0)
    cuts.append(len(edges))
    cuts = np.maximum.accumulate(np.array(cuts))
    return [(int(cuts[i]), int(cuts[i + 1])) for i in range(world)]
