"""TEST INFRASTRUCTURE ONLY — torch-on-CPU stand-ins for the kernel-backed ops, so the HOST logic of
cgat_b200 (module wiring, table form of the edge update, per-atom/per-rank split of the first MLP
layer, plans) can be checked against the reference goldens without a GPU.  Installed by the
`emulated_kernels` fixture via monkeypatch; the product never imports this."""
import torch

from cgat_b200 import graph, ops
from oracle import cgat_oracle as O


def build_edge_plan(edge_index, edge_attr, n_nodes, n_ranks=0):
    perm, rowptr = O.csr_by_destination(edge_index, n_nodes)
    i32 = lambda t: t.to(torch.int32)
    return graph.EdgePlan(n_nodes, edge_index.shape[1], i32(perm), i32(rowptr), i32(edge_index[0][perm]),
                          i32(edge_index[1][perm]), i32(edge_attr[perm]))


def build_segment_plan(index, n_seg):
    return graph.SegmentPlan(index.shape[0], n_seg, O.segment_ptr(index, n_seg).to(torch.int32),
                             index.to(torch.int32))


def seg_softmax(gate, value, plan=None, *, ptr=None, seg_of_row=None, n_seg=None, u=None, eps=1e-16):
    if plan is not None:
        seg_of_row, n_seg = plan.index, plan.n_seg
    idx = seg_of_row.long()
    mx = O.seg_max(gate.detach(), idx, n_seg)
    e = (gate - mx[idx]).exp()
    if u is not None:
        e = e * u.view(-1, 1, 1)
    alpha = e / (O.seg_sum(e, idx, n_seg)[idx] + eps)
    return O.seg_sum(alpha * value, idx, n_seg)


def install(monkeypatch):
    from cgat_b200 import CGAT, roost_message
    for mod in (graph, CGAT, roost_message):
        if hasattr(mod, "build_edge_plan"):
            monkeypatch.setattr(mod, "build_edge_plan", build_edge_plan)
        if hasattr(mod, "build_segment_plan"):
            monkeypatch.setattr(mod, "build_segment_plan", build_segment_plan)
    monkeypatch.setattr(ops, "seg_softmax", seg_softmax)

    def hyper_linear(z, weight, bias, y, out_ch, e=None):
        in_ch = y.shape[1]
        p = torch.addmm(bias, z, weight.t())
        w = p[:, : in_ch * out_ch].view(-1, out_ch, in_ch)
        return torch.baddbmm(p[:, in_ch * out_ch:].unsqueeze(2), w, y.unsqueeze(2)).squeeze(2)
    monkeypatch.setattr(ops, "hyper_linear", hyper_linear)
