"""Per-batch integer structure, built on the device by cgat_csr_build /
cgat_segment_ptr and reused by every layer, forward and backward.

The reference has no such object: PyG's propagate re-derives gathers from `edge_index` and
torch_scatter reduces with atomics on every call (reference CGAT/CGAT.py:313-326).  Grouping the
edge list by destination once makes every softmax segment a contiguous range.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch

from . import _lib


@dataclass
class SourceOrder:
    """The same edges grouped by SOURCE atom (backward only: dL/dP src-block and dL/dT are sums over the
    out-edges of an atom / the edges of a rank)."""
    rowptr: torch.Tensor    # (N+1,) int32
    seg: torch.Tensor       # (E,) int32  source atom of each src-sorted edge
    row: torch.Tensor       # (E,) int32  position of that edge in the destination-sorted order
    rank: torch.Tensor      # (E,) int32


@dataclass
class EdgePlan:
    n_nodes: int
    n_edges: int
    perm: torch.Tensor      # (E,) int32  stable argsort of destinations
    rowptr: torch.Tensor    # (N+1,) int32
    src: torch.Tensor       # (E,) int32  source atom of each dst-sorted edge
    dst: torch.Tensor       # (E,) int32  destination atom (= segment id)
    rank: torch.Tensor      # (E,) int32  shell rank of each dst-sorted edge
    edge_index: torch.Tensor = None
    edge_attr: torch.Tensor = None
    _by_source: SourceOrder = None

    def by_source(self) -> SourceOrder:
        if self._by_source is None:
            flipped = _csr(self.edge_index.flip(0).contiguous(), self.edge_attr, self.n_nodes)
            perm_s, rowptr_s, _, seg_s, rank_s = flipped
            inv = torch.empty_like(self.perm)
            inv[self.perm.long()] = torch.arange(self.n_edges, dtype=torch.int32, device=self.perm.device)
            self._by_source = SourceOrder(rowptr_s, seg_s, inv[perm_s.long()].contiguous(), rank_s)
        return self._by_source


@dataclass
class SegmentPlan:
    n_rows: int
    n_seg: int
    ptr: torch.Tensor       # (n_seg+1,) int32
    index: torch.Tensor     # (n_rows,) int32


def _csr(edge_index, edge_attr, n_nodes, n_ranks=0):
    dev = edge_index.device
    E = edge_index.shape[1]
    lib = _lib.load()
    i32 = dict(dtype=torch.int32, device=dev)
    perm, src, dst, rank = (torch.empty(E, **i32) for _ in range(4))
    rowptr = torch.empty(n_nodes + 1, **i32)
    ws_bytes = int(lib.cgat_csr_workspace_bytes(E, n_nodes))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    _lib.call("cgat_csr_build", _lib.ptr(edge_index), _lib.ptr(edge_attr), E, n_nodes, _lib.ptr(perm),
              _lib.ptr(rowptr), _lib.ptr(src), _lib.ptr(dst), _lib.ptr(rank), int(n_ranks), _lib.ptr(ws), ws_bytes,
              _lib.stream())
    return perm, rowptr, src, dst, rank


def build_edge_plan(edge_index: torch.Tensor, edge_attr: torch.Tensor, n_nodes: int, n_ranks: int = 0) -> EdgePlan:
    """edge_index (2,E) int64 [source; destination] (reference CGAT/data.py:140), edge_attr (E,) int64.
    n_ranks: rows of the shell-rank embedding table (0: unknown, the kernels' limit of 32 is used); out-of-range node ids
    / ranks raise a sticky device flag (_lib.check_status) instead of gathering out of bounds."""
    if edge_index.dtype != torch.int64 or edge_attr.dtype != torch.int64:
        raise TypeError("edge_index / edge_attr must be int64 (the reference's layout)")
    edge_index = edge_index.contiguous()
    edge_attr = edge_attr.contiguous()
    perm, rowptr, src, dst, rank = _csr(edge_index, edge_attr, n_nodes, n_ranks)
    return EdgePlan(n_nodes, edge_index.shape[1], perm, rowptr, src, dst, rank, edge_index, edge_attr)


def build_segment_plan(index: torch.Tensor, n_seg: int) -> SegmentPlan:
    """ptr of a SORTED int64 segment index (batch.batch, Roost self_fea_idx / crystal_elem_idx)."""
    if index.dtype != torch.int64:
        raise TypeError("segment index must be int64")
    index = index.contiguous()
    n = index.shape[0]
    ptr = torch.empty(n_seg + 1, dtype=torch.int32, device=index.device)
    idx32 = torch.empty(n, dtype=torch.int32, device=index.device)
    _lib.call("cgat_segment_ptr", _lib.ptr(index), n, n_seg, _lib.ptr(ptr), _lib.ptr(idx32), _lib.stream())
    return SegmentPlan(n, n_seg, ptr, idx32)
