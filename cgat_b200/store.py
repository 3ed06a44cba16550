"""Crystal store resident in HBM + device-side collation.

The reference assembles every batch in Python on the host: `CompositionData.__getitem__` per crystal (reference
CGAT/data.py:61-144), `Batch.from_data_list` and `collate_batch` (CGAT/lightning_module.py:199-202,
CGAT/roost_message.py:400-458), then copies eleven tensors to the GPU.  With 180 GB of HBM a whole screening set
fits on the device (1 M crystals of ~11 atoms x 200 features = 8.8 GB), so here the per-crystal samples are packed ONCE
into ragged arrays, kept on the GPU, and a batch is assembled by two kernel launches from the list of selected crystal
ids (cgat_collate_plan / cgat_collate_fill) — bucket padding included (batching.pad_batch's dummy crystal), so the
result feeds graphed.GraphedForward / GraphedTrainStep directly.  Bit-exact against the host path
(tests/test_gpu_kernels.py::test_device_collation_bit_exact).
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib, batching
from .synthetic import GraphBatch, SyntheticBatch


class CrystalStore:
    """Packed ragged arrays of a set of crystals.

    x (A, D) f32; nbr / rank (A, K) int32 — neighbour index LOCAL to the crystal and shell rank (the reference's
    `nbr_fea_idx` / `nbr_fea` after the `[:, :max_nbr]` slice, CGAT/data.py:115-120); atom_ptr (C+1) int64;
    comp_w (M,) / comp_fea (M, D): weights and features of each crystal's distinct elements (CGAT/data.py:81-103);
    comp_ptr (C+1); y (C,).  Host copies of the two pointer arrays stay in numpy: batch sizes are known without a sync."""

    def __init__(self, x, nbr, rank, atom_ptr, comp_w, comp_fea, comp_ptr, y):
        self.x, self.nbr, self.rank, self.atom_ptr = x, nbr, rank, atom_ptr
        self.comp_w, self.comp_fea, self.comp_ptr, self.y = comp_w, comp_fea, comp_ptr, y
        self.atom_ptr_host = atom_ptr.cpu().numpy()
        self.comp_ptr_host = comp_ptr.cpu().numpy()
        self.max_nbr = nbr.shape[1]

    @property
    def num_crystals(self):
        return self.atom_ptr_host.shape[0] - 1

    @classmethod
    def from_batch(cls, sb: SyntheticBatch):
        """Pack a collated batch (e.g. synthetic.make_batch of the whole data set) crystal by crystal."""
        g = sb.graph
        n_c = torch.as_tensor(np.asarray(sb.n_atoms), dtype=torch.int64)
        atom_ptr = torch.cat([torch.zeros(1, dtype=torch.int64), torch.cumsum(n_c, 0)])
        N = g.x.shape[0]
        K = g.edge_index.shape[1] // N
        off = torch.repeat_interleave(atom_ptr[:-1], n_c)                       # node offset of each atom's crystal
        nbr = (g.edge_index[1].view(N, K) - off.view(-1, 1)).to(torch.int32)
        rank = g.edge_attr.view(N, K).to(torch.int32)
        w, fea, _, _, cidx = sb.roost
        m_c = torch.bincount(cidx, minlength=n_c.shape[0])
        comp_ptr = torch.cat([torch.zeros(1, dtype=torch.int64), torch.cumsum(m_c, 0)])
        y = g.y if g.y is not None else torch.zeros(n_c.shape[0])
        return cls(g.x.contiguous(), nbr.contiguous(), rank.contiguous(), atom_ptr, w.reshape(-1).contiguous(),
                   fea.contiguous(), comp_ptr, y.to(torch.float32).contiguous())

    @classmethod
    def from_samples(cls, samples):
        """Per-crystal samples as the reference's CompositionData.__getitem__ yields them:
        (graph with x / edge_index / edge_attr / y, (weights, fea, self_idx, nbr_idx))."""
        return cls.from_batch(batching.collate(samples))

    def to(self, device):
        mv = lambda t: t.to(device)
        out = CrystalStore.__new__(CrystalStore)
        out.x, out.nbr, out.rank, out.atom_ptr = mv(self.x), mv(self.nbr), mv(self.rank), mv(self.atom_ptr)
        out.comp_w, out.comp_fea, out.comp_ptr, out.y = mv(self.comp_w), mv(self.comp_fea), mv(self.comp_ptr), mv(self.y)
        out.atom_ptr_host, out.comp_ptr_host, out.max_nbr = self.atom_ptr_host, self.comp_ptr_host, self.max_nbr
        return out

    def nbytes(self):
        return sum(t.numel() * t.element_size() for t in (self.x, self.nbr, self.rank, self.atom_ptr, self.comp_w,
                                                         self.comp_fea, self.comp_ptr, self.y))

    # ------------------------------------------------------------------------------------------
    def sizes(self, sel_host):
        """(N, Nc, Mc) of the batch made of crystals `sel_host` (numpy / list of ids) — host arithmetic only."""
        sel = np.asarray(sel_host, dtype=np.int64)
        n = self.atom_ptr_host[sel + 1] - self.atom_ptr_host[sel]
        m = self.comp_ptr_host[sel + 1] - self.comp_ptr_host[sel]
        return int(n.sum()), int(m.sum()), int((m * (m - 1)).sum()), n

    def collate(self, sel_host, sel_dev=None, buckets=batching.DEFAULT_BUCKETS) -> SyntheticBatch:
        """The collated, bucket-padded batch of crystals `sel_host` (ids into the store), built on the device.
        `sel_dev`: the same ids as an int64 device tensor if the caller already has it there (else it is copied).
        Identical, bit for bit, to batching.pad_batch(batching.collate(samples[sel])) moved to the device."""
        if not self.x.is_cuda:
            raise _lib.CgatLibraryError("CrystalStore.collate runs on the GPU: move the store with .to('cuda') first")
        dev = self.x.device
        n_atoms, n_comp, n_pairs, n_c = self.sizes(sel_host)
        B = len(sel_host)
        if sel_dev is None:
            sel_dev = torch.as_tensor(np.asarray(sel_host, dtype=np.int64)).to(dev, non_blocking=True)
        n_pad, nc_pad, mc_pad = batching.padded_sizes(n_atoms, n_comp, n_pairs, buckets)
        D, K = self.x.shape[1], self.max_nbr
        i64 = dict(dtype=torch.int64, device=dev)
        f32 = dict(dtype=torch.float32, device=dev)
        plan = torch.empty((3, B + 1), **i64)
        st = _lib.stream()
        _lib.call("cgat_collate_plan", _lib.ptr(sel_dev), B, _lib.ptr(self.atom_ptr), _lib.ptr(self.comp_ptr),
                  _lib.ptr(plan[0]), _lib.ptr(plan[1]), _lib.ptr(plan[2]), st,
                  work=dict(key="collate_plan", bound="hbm", bytes=8.0 * 5 * B))
        x = torch.empty((n_pad, D), **f32)
        edge_index = torch.empty((2, n_pad * K), **i64)
        edge_attr = torch.empty(n_pad * K, **i64)
        batch = torch.empty(n_pad, **i64)
        y = torch.empty(B + 1, **f32)
        w = torch.empty((nc_pad, 1), **f32)
        fea = torch.empty((nc_pad, D), **f32)
        self_idx, nbr_idx = torch.empty(mc_pad, **i64), torch.empty(mc_pad, **i64)
        cry_idx = torch.empty(nc_pad, **i64)
        nbytes = 4.0 * (2 * n_pad * D + 2 * nc_pad * D) + 8.0 * (3 * n_pad * K + n_pad + 2 * mc_pad + nc_pad) + 8.0 * n_pad * K
        _lib.call("cgat_collate_fill", _lib.ptr(sel_dev), B, _lib.ptr(self.x), _lib.ptr(self.nbr), _lib.ptr(self.rank),
                  _lib.ptr(self.atom_ptr), _lib.ptr(self.comp_w), _lib.ptr(self.comp_fea), _lib.ptr(self.comp_ptr),
                  _lib.ptr(self.y), D, K, _lib.ptr(plan[0]), _lib.ptr(plan[1]), _lib.ptr(plan[2]), _lib.ptr(x),
                  _lib.ptr(edge_index), _lib.ptr(edge_attr), _lib.ptr(batch), _lib.ptr(y), _lib.ptr(w), _lib.ptr(fea),
                  _lib.ptr(self_idx), _lib.ptr(nbr_idx), _lib.ptr(cry_idx), n_atoms, n_pad, n_comp, nc_pad, n_pairs,
                  mc_pad, st, work=dict(key="collate_fill", bound="hbm", bytes=nbytes,
                                        note="gathers the selected crystals' rows once, writes the int64 batch layout"))
        gb = GraphBatch(x, edge_index, edge_attr, batch, y, num_graphs=B + 1)
        n_atoms_arr = np.concatenate([n_c, np.array([n_pad - n_atoms], dtype=n_c.dtype)])
        return SyntheticBatch(gb, (w, fea, self_idx, nbr_idx, cry_idx), n_atoms_arr)
