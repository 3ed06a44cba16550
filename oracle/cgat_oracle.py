"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference CGAtNet forward (hyllios/CGAT).

This file is the parity oracle.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import it; the product (cgat_b200/) never does.

It restates, scatter-free and dependency-free (plain torch on CPU; works in fp32 or fp64), the
arithmetic of the reference's hot path, function by function, citing reference file:line
(paths relative to the reference root).  It consumes a reference-compatible ``state_dict``
(same names/shapes as the reference's CGAtNet) so reference checkpoints and the
seeded weights of oracle/weights.py drive it directly.  Gradients come from torch autograd over
this restatement.

PINNING: the reference ships no tests or golden vectors.  The restatement is pinned
against outputs of the UNMODIFIED reference modules run in the build container
(oracle/make_golden.py -> tests/golden/*.npz, checked by tests/test_oracle_golden.py).
Residual risk, stated in DESIGN.md: the third-party torch_scatter / torch_geometric semantics are
reproduced by stand-ins (oracle/standins.py) from their documented behaviour.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------- segment helpers
def seg_sum(src, index, size):
    """torch_scatter.scatter_add(src, index, dim=0, dim_size=size) restated with index_add."""
    out = torch.zeros((size,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    return out.index_add(0, index, src)


def seg_max(src, index, size):
    """torch_scatter.scatter_max(...)[0]: per-segment max, empty segments = 0."""
    idx = index.view((-1,) + (1,) * (src.dim() - 1)).expand_as(src)
    out = torch.zeros((size,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    return out.scatter_reduce(0, idx, src, reduce="amax", include_self=False)


def pyg_softmax(src, index, size):
    """torch_geometric.utils.softmax as called at CGAT/CGAT.py:59 and :323:
    exp(src - segmax[index]) / (segsum(exp)[index] + 1e-16), along dim 0."""
    mx = seg_max(src.detach(), index, size)
    e = (src - mx[index]).exp()
    return e / (seg_sum(e, index, size)[index] + 1e-16)


# ----------------------------------------------------------------------------- dense blocks
def multi_head_network(sd, pre, fea, heads):
    """MultiHeadNetwork.forward, CGAT/CGAT.py:103-109 (ctor :70-101).
    Grouped Conv1d(k=1) over the input repeated `heads` times == `heads` independent 2-layer MLPs;
    LeakyReLU() uses the default slope 0.01 (CGAT/CGAT.py:95)."""
    w1, b1 = sd[pre + "fc_in.weight"], sd[pre + "fc_in.bias"]      # (H*Hd, In, 1), (H*Hd,)
    w2, b2 = sd[pre + "fc_out.weight"], sd[pre + "fc_out.bias"]    # (H*Out, Hd, 1), (H*Out,)
    n, din = fea.shape
    hd = w1.shape[0] // heads
    dout = w2.shape[0] // heads
    w1 = w1.view(heads, hd, din)
    w2 = w2.view(heads, dout, hd)
    hid = torch.einsum("ni,hki->nhk", fea, w1) + b1.view(1, heads, hd)
    hid = F.leaky_relu(hid, 0.01)
    out = torch.einsum("nhk,hok->nho", hid, w2) + b2.view(1, heads, dout)
    return out                                                     # (n, H, Out)


def simple_network(sd, pre, fea):
    """SimpleNetwork.forward, CGAT/message_changed.py:58-63 (and roost_message.py:351-355)."""
    i = 0
    while pre + f"fcs.{i}.weight" in sd:
        fea = F.leaky_relu(F.linear(fea, sd[pre + f"fcs.{i}.weight"], sd[pre + f"fcs.{i}.bias"]), 0.01)
        i += 1
    return F.linear(fea, sd[pre + "fc_out.weight"], sd[pre + "fc_out.bias"])


def residual_network(sd, pre, fea, rezero, last_layer=True):
    """ResidualNetwork.forward + Rezero, CGAT/message_changed.py:120-135, :74-75."""
    i = 0
    while pre + f"fcs.{i}.weight" in sd:
        h = F.relu(F.linear(fea, sd[pre + f"fcs.{i}.weight"], sd[pre + f"fcs.{i}.bias"]))
        if rezero:
            h = sd[pre + f"rezeros.{i}.alpha"] * h
        rk = pre + f"res_fcs.{i}.weight"
        fea = h + (F.linear(fea, sd[rk]) if rk in sd else fea)
        i += 1
    if last_layer:
        return F.linear(fea, sd[pre + "fc_out.weight"], sd[pre + "fc_out.bias"])
    return fea


# ----------------------------------------------------------------------------- hypernetwork
def fc_block(sd, pre, h):
    """FCBlock.forward, CGAT/Hypernetworksmp.py:82-83 (ctor :36-69): 4x(Linear+Tanh) then Linear."""
    i = 0
    while pre + f"net.{i}.net.0.weight" in sd:
        h = torch.tanh(F.linear(h, sd[pre + f"net.{i}.net.0.weight"], sd[pre + f"net.{i}.net.0.bias"]))
        i += 1
    return F.linear(h, sd[pre + f"net.{i}.weight"], sd[pre + f"net.{i}.bias"])


def hyper_fc_apply(sd, pre, hyper_in, y):
    """HyperFC.forward + the predicted nn.Sequential applied to y
    (CGAT/Hypernetworksmp.py:176-185, 109-114, 243-254, 205-209, 103-107).
    Layers 0..2 are HyperLayer (hyper_linear + LayerNorm(no affine, eps 1e-5) + Tanh), layer 3 is a
    bare HyperLinear (outermost_linear=True, :277-284)."""
    f_in = y.shape[1]
    j = 0
    while True:
        if pre + f"layers.{j}.hyper_linear.hypo_params.net.0.net.0.weight" in sd:
            p = fc_block(sd, pre + f"layers.{j}.hyper_linear.hypo_params.", hyper_in)
            last = False
        elif pre + f"layers.{j}.hypo_params.net.0.net.0.weight" in sd:
            p = fc_block(sd, pre + f"layers.{j}.hypo_params.", hyper_in)
            last = True
        else:
            break
        out_ch = p.shape[1] // (f_in + 1)
        w = p[:, : f_in * out_ch].reshape(-1, out_ch, f_in)      # (N, out, in)   :247,252
        b = p[:, f_in * out_ch: f_in * out_ch + out_ch]            # (N, out)       :248-251
        y = torch.einsum("noi,ni->no", w, y) + b                   # BatchLinear    :205-209
        if not last:
            y = torch.tanh(F.layer_norm(y, (out_ch,), eps=1e-5))   # norm_nl        :103-107
        f_in = out_ch
        j += 1
    return y


# ----------------------------------------------------------------------------- message passing
def gat_conv_nodes(sd, pre, x, edge_index, edge_attr, x_0, heads, first, clamp_damping=True):
    """GATConvNodes.forward/message/aggregate/update, CGAT/CGAT.py:307-335 with PyG's default flow
    source_to_target: x_j = x[edge_index[0]], x_i = x[edge_index[1]], softmax/aggregation index =
    edge_index[1]."""
    src, dst = edge_index[0], edge_index[1]
    n = x.shape[0]
    m = torch.cat([x[dst], edge_attr, x[src]], dim=-1)            # :320
    alpha = multi_head_network(sd, pre + "MH_A.", m, heads)        # :321
    msg = multi_head_network(sd, pre + "MH_M.", m, heads)          # :322
    alpha = pyg_softmax(alpha, dst, n)                             # :323 (scalar attention broadcasts)
    aggr = seg_sum(msg * alpha, dst, n)                            # :326 + aggr='add' (:275)
    aggr = aggr.mean(dim=1)                                        # :329
    if first:                                                      # :330-331  H_Net_0(h_0=x, x=aggr)
        return hyper_fc_apply(sd, pre + "Pooling_NN.Hyper.", x, aggr)
    d = sd[pre + "Pooling_NN.damping"]                             # :332-333  H_Net(h_0=x_0, h_t=x, x=aggr)
    if clamp_damping:                                              # Hypernetworksmp.py:310-311
        with torch.no_grad():
            d.data = d.data.clamp(0.0, 1.0)
    hyper_in = d * x_0 + (1 - d) * aggr                            # :312 (h_t unused)
    return hyper_fc_apply(sd, pre + "Pooling_NN.Hyper.", hyper_in, aggr)


def gat_conv_edges(sd, pre, x, edge_index, edge_attr, heads, as_written=False, no_hyper=True, first=False,
                   edge_attr_0=None, clamp_damping=True):
    """GATConvEdges.forward, CGAT/CGAT.py:208-230.  no_hyper=True: lines :209-223 compute an attention whose
    result is overwritten at :224-225; `as_written=True` executes that dead work too (used only to time the
    reference 'as written').  no_hyper=False (:226-229): the head-softmaxed message drives a per-EDGE hypernetwork
    update, H_Net_0(h_0=edge_attr, x=aggr) in the first layer, H_Net(h_0=edge_attr_0, h_t=edge_attr, x=aggr) after."""
    if not no_hyper:
        m = torch.cat([x[edge_index[0]], edge_attr, x[edge_index[1]]], dim=-1)     # :209-211 (x_i = x[source] here)
        a = multi_head_network(sd, pre + "MH_A.", m, heads).exp()                  # :212, :214
        mm = multi_head_network(sd, pre + "MH_M.", m, heads)                       # :213
        a = a / a.sum(dim=1, keepdim=True)                                         # :216-219 (softmax over HEADS)
        aggr = (mm * a).mean(dim=1)                                                # :222-223
        if first:                                                                  # :226-227
            return hyper_fc_apply(sd, pre + "Pooling_NN.Hyper.", edge_attr, aggr)
        d = sd[pre + "Pooling_NN.damping"]                                         # :228-229, Hypernetworksmp.py:309-313
        if clamp_damping:
            with torch.no_grad():
                d.data = d.data.clamp(0.0, 1.0)
        return hyper_fc_apply(sd, pre + "Pooling_NN.Hyper.", d * edge_attr_0 + (1 - d) * aggr, aggr)
    if as_written:
        m = torch.cat([x[edge_index[0]], edge_attr, x[edge_index[1]]], dim=-1)
        a = multi_head_network(sd, pre + "MH_A.", m, heads).exp()
        mm = multi_head_network(sd, pre + "MH_M.", m, heads)
        a = a / a.sum(dim=1, keepdim=True)
        _ = (mm * a).mean(dim=1)
    return simple_network(sd, pre + "Pooling_NN.", edge_attr)      # :224-225


def weighted_attention(sd, pre, fea, index, weights, size, has_message_nn):
    """WeightedAttention.forward, CGAT/roost_message.py:302-317."""
    gate = simple_network(sd, pre + "gate_nn.", fea)               # :305
    gate = gate - seg_max(gate.detach(), index, size)[index]       # :307
    gate = (weights ** sd[pre + "pow"]) * gate.exp()               # :308
    gate = gate / (seg_sum(gate, index, size)[index] + 1e-13)      # :311
    if has_message_nn:
        fea = simple_network(sd, pre + "message_nn.", fea)         # :313
    return seg_sum(gate * fea, index, size)                        # :315


def roost(sd, pre, weights, fea, self_idx, nbr_idx, crystal_idx, n_crystals):
    """Roost.forward + MessageLayer.forward, CGAT/roost_message.py:212-264, 112-153."""
    x = F.linear(fea, sd[pre + "embedding.weight"], sd[pre + "embedding.bias"])   # :240
    x = torch.cat([x, weights], dim=1)                                            # :245
    l = 0
    while pre + f"graphs.{l}.pooling.0.pow" in sd:
        nbr_w = weights[nbr_idx, :]                                               # :139
        cat = torch.cat([x[self_idx, :], x[nbr_idx, :]], dim=1)                   # :140-142
        head = weighted_attention(sd, pre + f"graphs.{l}.pooling.0.", cat, self_idx, nbr_w,
                                  x.shape[0], True)                               # :146-149
        x = head + x                                                              # :153-154 (mean over 1 head)
        l += 1
    return weighted_attention(sd, pre + "cry_pool.0.", x, crystal_idx, weights, n_crystals, False)  # :253-260


def mh_attention(sd, pre, fea, cry_fea, index, size, heads):
    """MHAttention.forward, CGAT/CGAT.py:50-62."""
    m = multi_head_network(sd, pre + "MH_M.", fea, heads)                          # :53
    a_in = torch.cat([fea, cry_fea[index]], dim=1)                                 # :55-57 (stack+transpose+reshape)
    alpha = multi_head_network(sd, pre + "MH_A.", a_in, heads)                     # :58
    alpha = pyg_softmax(alpha, index, size)                                        # :59
    return seg_sum((alpha * m).reshape(fea.shape[0], -1), index, size)             # :60-61


# ----------------------------------------------------------------------------- the model
def cgat_forward(sd, cfg, graph, roost_in, last_layer=True, return_graph_embedding=False,
                 as_written=False, return_intermediates=False):
    """CGAtNet.forward, CGAT/CGAT.py:540-600, update_edges=True / no_hyper=True branch.

    cfg: dict(n_graph, msg_heads, mean_pooling, rezero).  graph: x, edge_index, edge_attr, batch.
    roost_in: (weights, fea, self_idx, nbr_idx, crystal_idx)."""
    heads, n_graph = cfg["msg_heads"], cfg["n_graph"]
    edge_index, batch = graph.edge_index, graph.batch
    n_cry = int(batch[-1]) + 1
    inter = {}
    edge_attr = sd["nbr_embedding.weight"][graph.edge_attr]                        # :569
    x = F.linear(graph.x, sd["embedding.weight"])                                 # :570
    x_0 = x                                                                        # :571
    edge_attr_0 = edge_attr                                                        # :574
    no_hyper = cfg.get("no_hyper", True)
    for l in range(n_graph):                                                       # :580-585
        node_upd = gat_conv_nodes(sd, f"graphs.{l}.Node.", x, edge_index, edge_attr, x_0, heads, l == 0)
        edge_attr = edge_attr + gat_conv_edges(sd, f"graphs.{l}.Edge.", x, edge_index, edge_attr, heads,
                                               as_written, no_hyper, l == 0, edge_attr_0)
        x = x + node_upd
        inter[f"x{l + 1}"] = x
    weights, fea, self_idx, nbr_idx, crystal_idx = roost_in
    cry = roost(sd, "roost.", weights, fea, self_idx, nbr_idx, crystal_idx, n_cry)  # :587
    inter["roost"] = cry
    cry = mh_attention(sd, "cry_pool.", x, cry, batch, n_cry, heads)                # :588
    if cfg["mean_pooling"]:                                                         # :590-592
        cry = cry.view(n_cry, heads, -1).mean(dim=1)
    inter["pooled"] = cry
    if return_graph_embedding:
        return cry
    out = residual_network(sd, "output_nn.", cry, cfg["rezero"], last_layer)       # :595/599
    if return_intermediates:
        return out, inter
    return out


# ----------------------------------------------------------------------------- integer structure
def csr_by_destination(edge_index, n_nodes):
    """The integer structures the CUDA path builds (no reference counterpart):
    perm = stable argsort of destinations, rowptr = exclusive cumsum of in-degrees."""
    dst = edge_index[1]
    perm = torch.sort(dst, stable=True)[1]
    deg = torch.bincount(dst, minlength=n_nodes)
    rowptr = torch.zeros(n_nodes + 1, dtype=torch.long)
    rowptr[1:] = torch.cumsum(deg, 0)
    return perm, rowptr


def segment_ptr(index, size):
    """ptr (size+1) of a sorted segment index vector (crystal_ptr, Roost rowptrs)."""
    cnt = torch.bincount(index, minlength=size)
    ptr = torch.zeros(size + 1, dtype=torch.long)
    ptr[1:] = torch.cumsum(cnt, 0)
    return ptr
