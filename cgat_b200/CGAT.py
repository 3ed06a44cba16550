"""CGAtNet — drop-in for the reference's CGAT/CGAT.py (`importlib.import_module(version).CGAtNet`,
reference CGAT/lightning_module.py:165-176): same constructor, forward signature, parameter names
and shapes (reference state_dicts load with strict=True), new compute underneath.

What changes relative to the reference:
  * edges are grouped by destination once per batch (graph.EdgePlan) instead of gathered /
    scattered with atomics on every call;
  * with no_hyper=True the edge embedding is a function of the integer shell rank only
    (reference CGAT.py:224-225, :583-584), so it is carried as a (K+1, F_e) table and never
    materialised per edge;
  * the first layer of the gate / message MLPs is linear in cat[x_i, e_ij, x_j] (reference
    CGAT.py:320-322), so it is evaluated per ATOM (x W_i^T, x W_j^T) and per RANK (e W_e^T + b) and
    only summed per edge — E/N = max_nbr times fewer FLOPs than the per-edge contraction, identical
    up to fp32 reassociation;
  * the dead attention of GATConvEdges (reference CGAT.py:209-223, overwritten at :224-225) is not
    computed; its parameters exist and receive no gradient, as in the reference;
  * segmented softmax / weighted sums are deterministic and atomic-free (ops.seg_softmax).
"""
from __future__ import annotations

import itertools

import torch
import torch.nn as nn

from . import ops
from .graph import EdgePlan, SegmentPlan, build_edge_plan, build_segment_plan
from .Hypernetworksmp import H_Net, H_Net_0
from .message_changed import ResidualNetwork, SimpleNetwork
from .roost_message import Roost


class MultiHeadNetwork(nn.Module):
    """`nb_heads` independent two-layer MLPs stored, like the reference (CGAT.py:65-112), as grouped
    Conv1d(k=1) parameters: fc_in.weight (H*Hd, In, 1), fc_out.weight (H*Out, Hd, 1).
    LeakyReLU slope is the default 0.01 (reference CGAT.py:95)."""

    def __init__(self, input_dim, output_dim, hidden_layer_dim, nb_heads, view=True):
        super().__init__()
        self.input_dim, self.output_dim = input_dim, output_dim
        self.hidden_dim, self.nb_heads = hidden_layer_dim, nb_heads
        self.fc_in = nn.Conv1d(input_dim * nb_heads, hidden_layer_dim * nb_heads, kernel_size=1, groups=nb_heads)
        self.acts = nn.LeakyReLU()
        self.fc_out = nn.Conv1d(hidden_layer_dim * nb_heads, output_dim * nb_heads, kernel_size=1, groups=nb_heads)
        self.view = view

    def w_in(self):
        return self.fc_in.weight.squeeze(-1)                       # (H*Hd, In)

    def w_out(self):
        return self.fc_out.weight.squeeze(-1).view(self.nb_heads, self.output_dim, self.hidden_dim)

    def forward(self, fea):
        """(n, In) -> (n, H, Out); all heads share the input (no `.repeat` copy, cf. reference :105)."""
        fea = fea.reshape(-1, self.input_dim)
        return ops.multi_head_mlp(fea, self.w_in(), self.fc_in.bias, self.w_out(), self.fc_out.bias,
                                  self.nb_heads)

    def __repr__(self):
        return type(self).__name__


class MHAttention(nn.Module):
    """Crystal-level attention pool (reference CGAT.py:14-62): per atom a message MLP on x and a
    gate MLP on [x ; roost_vec[crystal]], softmax over the atoms of each crystal, weighted sum
    -> (C, heads*out_channels), head-major."""

    def __init__(self, in_channels, out_channels, heads=1, vector_attention=False):
        super().__init__()
        self.heads, self.out_channels = heads, out_channels
        gate_out = out_channels if vector_attention else 1
        self.MH_A = MultiHeadNetwork(2 * in_channels, gate_out, in_channels, heads, view=False)
        self.MH_M = MultiHeadNetwork(in_channels, out_channels, in_channels, heads)

    def forward(self, fea, cry_fea, index, size=None, plan: SegmentPlan | None = None):
        if plan is None:
            size = int(index[-1]) + 1 if size is None else size    # the reference's host sync (:52)
            plan = build_segment_plan(index, size)
        msg = self.MH_M(fea)
        gate = self.MH_A(torch.cat([fea, cry_fea[index]], dim=1))
        out = ops.seg_softmax(gate, msg, plan, eps=1e-16)
        return out.reshape(plan.n_seg, self.heads * self.out_channels)


class GATConvEdges(nn.Module):
    """Edge update (reference CGAT.py:115-230).  With no_hyper=True only `Pooling_NN(edge_attr)`
    is live (:224-225); MH_A / MH_M are kept as parameters (state_dict parity) and never run."""

    def __init__(self, in_channels, out_channels, nbr_channels, heads=1, concat=True, negative_slope=0.2,
                 dropout=0, bias=True, vector_attention=False, first=False, no_hyper=True, **kwargs):
        super().__init__()
        self.in_channels, self.out_channels, self.nbr_channels = in_channels, out_channels, nbr_channels
        self.heads, self.vector_attention, self.first, self.no_hyper = heads, vector_attention, first, no_hyper
        width = 2 * in_channels + nbr_channels
        hidden = int(width / 1.5)
        self.MH_A = MultiHeadNetwork(width, out_channels if vector_attention else 1, hidden, heads)
        self.MH_M = MultiHeadNetwork(width, out_channels, hidden, heads)
        if no_hyper:
            self.Pooling_NN = SimpleNetwork(out_channels, out_channels, [out_channels])
        else:                                                       # reference CGAT.py:187-204
            net = H_Net_0 if first else H_Net
            self.Pooling_NN = net(out_channels, 3, out_channels, out_channels, 2, out_channels, out_channels)

    def forward(self, x, edge_index, edge_attr, x_0, size=None):
        """no_hyper=True: edge_attr may be per-edge (E, F_e) or the per-rank table (K+1, F_e) — the live update only
        looks at edge_attr itself (reference CGAT.py:224-225).  no_hyper=False (reference :209-223, :226-229): the
        message MLP output weighted by a softmax over HEADS of the gate MLP, averaged over heads, drives a per-edge
        hypernetwork update (H_Net_0(edge_attr, aggr) in the first layer, H_Net(edge_attr_0, edge_attr, aggr) after);
        edge_attr is then the per-edge (E, F_e) tensor in the ORIGINAL edge order and x_0 = edge_attr_0."""
        if self.no_hyper:
            return self.Pooling_NN(edge_attr)
        m = torch.cat([x.index_select(0, edge_index[0]), edge_attr, x.index_select(0, edge_index[1])], dim=1)
        alpha = self.MH_A(m).exp()
        alpha = alpha / alpha.sum(dim=1, keepdim=True)
        aggr = (self.MH_M(m) * alpha).mean(dim=1)
        if self.first:
            return self.Pooling_NN(edge_attr, aggr)
        return self.Pooling_NN(x_0, edge_attr, aggr)


class GATConvNodes(nn.Module):
    """Node attention layer (reference CGAT.py:233-340): gate and message MLPs over
    cat[x_i, e_ij, x_j], softmax over the in-edges of each destination atom per (head, channel),
    weighted sum, mean over heads, hypernetwork update."""

    def __init__(self, in_channels, out_channels, nbr_channels, heads=1, concat=False, negative_slope=0.2,
                 dropout=0, bias=True, final=False, vector_attention=False, first=False, **kwargs):
        super().__init__()
        self.in_channels, self.out_channels, self.nbr_channels = in_channels, out_channels, nbr_channels
        self.heads, self.final, self.first = heads, final, first
        self.vector_attention = vector_attention
        width = 2 * in_channels + nbr_channels
        hidden = int(width / 1.5)
        self.MH_A = MultiHeadNetwork(width, out_channels if vector_attention else 1, hidden, heads)
        self.MH_M = MultiHeadNetwork(width, out_channels, hidden, heads)
        if not final:
            net = H_Net_0 if first else H_Net
            self.Pooling_NN = net(out_channels, 3, out_channels, out_channels, 2, out_channels, out_channels)

    def aggregate(self, x, edge_table, plan: EdgePlan, edge_ids=None):
        """(N, F): mean over heads of the attention-weighted messages arriving at each atom
        (reference message() + scatter-add + update()'s head-mean, CGAT.py:319-329).  edge_table: the per-rank
        table (K+1, F_e), or — with edge_ids = plan.perm — per-edge features (E, F_e) in the original edge order."""
        return ops.edge_attention(x, edge_table, plan, self.MH_A, self.MH_M, self.heads, edge_ids=edge_ids)

    def forward(self, x, edge_table, plan: EdgePlan, x_0, edge_ids=None):
        aggr = self.aggregate(x, edge_table, plan, edge_ids)
        if self.final:
            return aggr
        if self.first:
            return self.Pooling_NN(x, aggr)                        # reference :330-331
        return self.Pooling_NN(x_0, x, aggr)                       # reference :332-333

    def __repr__(self):
        return f"{type(self).__name__}({self.in_channels}, {self.out_channels}, heads={self.heads})"


class CGAtNet(nn.Module):
    """Crystal graph attention network (reference CGAT.py:343-613)."""

    def __init__(self, orig_elem_fea_len, elem_fea_len, n_graph, nbr_embedding_size=128, neighbor_number=12,
                 mean_pooling=True, rezero=False, msg_heads=3, update_edges=False, vector_attention=False,
                 global_vector_attention=False, n_graph_roost=3, no_hyper=True):
        super().__init__()
        if not update_edges:
            raise NotImplementedError(
                "update_edges=False is broken in the reference itself (misaligned positional arguments at "
                "CGAT.py:408-413 raise a shape error); only update_edges=True is supported")
        self.mean_pooling, self.update_edges, self.no_hyper = mean_pooling, update_edges, no_hyper
        self.embedding = nn.Linear(orig_elem_fea_len, elem_fea_len, bias=False)
        self.nbr_embedding = nn.Embedding(neighbor_number + 1, nbr_embedding_size)
        self.graphs = nn.ModuleList(
            nn.ModuleDict({
                "Node": GATConvNodes(elem_fea_len, elem_fea_len, nbr_embedding_size, msg_heads, concat=True,
                                     vector_attention=vector_attention, first=(i == 0)),
                "Edge": GATConvEdges(elem_fea_len, nbr_embedding_size, nbr_embedding_size, msg_heads,
                                     concat=True, vector_attention=vector_attention, first=(i == 0),
                                     no_hyper=no_hyper),
            }) for i in range(n_graph))
        self.roost = Roost(orig_elem_fea_len, elem_fea_len, n_graph_roost)
        self.cry_pool = MHAttention(elem_fea_len, elem_fea_len, heads=msg_heads,
                                    vector_attention=global_vector_attention)
        self.msg_heads, self.elem_fea_len = msg_heads, elem_fea_len
        out_in = elem_fea_len if mean_pooling else elem_fea_len * msg_heads
        self.output_nn = ResidualNetwork(out_in, 2, [1024, 1024, 512, 512, 256, 256, 128], if_rezero=rezero)

    def forward(self, batch, roost, *, last_layer=True, return_graph_embedding=False):
        """batch: object with .x (N,orig) f32, .edge_index (2,E) i64, .edge_attr (E,) i64, .batch (N,) i64
        (optionally .num_graphs); roost: iterable of (weights, fea, self_idx, nbr_idx, crystal_idx),
        consumed once (reference lightning_module.py:202 passes a generator)."""
        weights, r_fea, self_idx, nbr_idx, r_cry = tuple(roost)
        n_atoms = batch.x.shape[0]
        n_cry = getattr(batch, "num_graphs", None)
        if n_cry is None:
            n_cry = int(batch.batch[-1]) + 1
        plan = build_edge_plan(batch.edge_index, batch.edge_attr, n_atoms, self.nbr_embedding.num_embeddings)
        cry_plan = build_segment_plan(batch.batch, n_cry)

        x = self.embedding(batch.x)                                # reference :570
        x_0 = x
        last = len(self.graphs) - 1
        if self.no_hyper:
            edge_table = self.nbr_embedding.weight                 # rank r -> e(r): reference :569
            for i, layer in enumerate(self.graphs):                # reference :580-585
                node_update = layer["Node"](x, edge_table, plan, x_0)
                if i < last:  # the last layer's edge update is never read (its params get no gradient)
                    edge_table = edge_table + layer["Edge"](x, None, edge_table, None)
                x = x + node_update
        else:
            # per-edge hypernetwork variant: the edge embedding depends on node data, so it is a real (E, F_e) tensor
            # (original edge order); the node attention reads it through plan.perm
            edge_attr = self.nbr_embedding(batch.edge_attr)        # reference :569
            edge_attr_0 = edge_attr
            for i, layer in enumerate(self.graphs):
                node_update = layer["Node"](x, edge_attr, plan, x_0, edge_ids=plan.perm)
                if i < last:
                    edge_attr = edge_attr + layer["Edge"](x, batch.edge_index, edge_attr, edge_attr_0)
                x = x + node_update

        comp = self.roost(weights, r_fea, self_idx, nbr_idx, r_cry, n_crystals=n_cry)   # :587
        pooled = self.cry_pool(x, comp, batch.batch, plan=cry_plan)                      # :588
        if self.mean_pooling:
            pooled = pooled.view(n_cry, self.msg_heads, self.elem_fea_len).mean(dim=1)   # :590-592
        if return_graph_embedding:
            return pooled
        return self.output_nn(pooled, last_layer=last_layer)

    def __repr__(self):
        return type(self).__name__

    def get_output_parameters(self):
        return self.output_nn.parameters()

    def get_hidden_parameters(self):
        return itertools.chain(self.embedding.parameters(), self.nbr_embedding.parameters(),
                               self.graphs.parameters(), self.roost.parameters(), self.cry_pool.parameters())
