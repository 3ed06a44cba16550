"""-m gpu: the CUDA kernels, through the C ABI, against the CPU oracle on the same seeded inputs."""
import numpy as np
import pytest
import torch

from cgat_b200 import _lib, graph, ops, synthetic
from oracle import cgat_oracle as O
from tests._cases import assert_close

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _rand_edges(n, e, seed):
    g = torch.Generator().manual_seed(seed)
    src = torch.randint(0, n, (e,), generator=g)
    dst = torch.randint(0, n, (e,), generator=g)
    attr = torch.randint(1, 25, (e,), generator=g)
    return torch.stack([src, dst]), attr


@pytest.mark.parametrize("n,e,seed", [(1, 1, 0), (7, 0, 1), (50, 600, 2), (5000, 60000, 3), (3, 5000, 4),
                                      (100000, 1200000, 5)])
def test_csr_build_bit_exact(n, e, seed):
    ei, attr = _rand_edges(n, e, seed)
    plan = graph.build_edge_plan(ei.to(DEV), attr.to(DEV), n)
    perm, rowptr = O.csr_by_destination(ei, n)
    assert torch.equal(plan.perm.cpu().long(), perm)
    assert torch.equal(plan.rowptr.cpu().long(), rowptr)
    assert torch.equal(plan.src.cpu().long(), ei[0][perm])
    assert torch.equal(plan.dst.cpu().long(), ei[1][perm])
    assert torch.equal(plan.rank.cpu().long(), attr[perm])


def test_csr_build_on_synthetic_batch():
    sb = synthetic.make_batch(500, 12, seed=1)
    g = sb.graph
    plan = graph.build_edge_plan(g.edge_index.to(DEV), g.edge_attr.to(DEV), g.num_nodes)
    perm, rowptr = O.csr_by_destination(g.edge_index, g.num_nodes)
    assert torch.equal(plan.perm.cpu().long(), perm) and torch.equal(plan.rowptr.cpu().long(), rowptr)


@pytest.mark.parametrize("sizes", [[3, 0, 0, 5, 1], [0, 0, 4], [1], [2, 2, 2, 0]])
def test_segment_ptr(sizes):
    idx = torch.repeat_interleave(torch.arange(len(sizes)), torch.tensor(sizes))
    plan = graph.build_segment_plan(idx.to(DEV), len(sizes))
    assert torch.equal(plan.ptr.cpu().long(), O.segment_ptr(idx, len(sizes)))
    assert torch.equal(plan.index.cpu().long(), idx)


def _softmax_case(n_seg, heads, f, fa, with_u, seed, max_len=40):
    g = torch.Generator().manual_seed(seed)
    lens = torch.randint(0, max_len, (n_seg,), generator=g)
    lens[0] = 0  # an empty segment
    idx = torch.repeat_interleave(torch.arange(n_seg), lens)
    n = idx.numel()
    gate = 3 * torch.randn(n, heads, fa, generator=g)
    val = torch.randn(n, heads, f, generator=g)
    u = (torch.rand(n, generator=g) + 0.05) if with_u else None
    return idx, gate, val, u


@pytest.mark.parametrize("heads,f,fa,with_u,eps", [(5, 128, 128, False, 1e-16), (5, 128, 1, False, 1e-16),
                                                   (1, 128, 1, True, 1e-13), (3, 20, 20, True, 1e-16),
                                                   (8, 256, 256, False, 1e-16)])
def test_seg_softmax_fwd_bwd(heads, f, fa, with_u, eps):
    n_seg = 97
    idx, gate, val, u = _softmax_case(n_seg, heads, f, fa, with_u, seed=heads * 7 + f)
    # oracle (fp64 on CPU)
    gd, vd = gate.double().requires_grad_(True), val.double().requires_grad_(True)
    ud = u.double().requires_grad_(True) if with_u else None
    e = (gd - O.seg_max(gd.detach(), idx, n_seg)[idx]).exp()
    if with_u:
        e = e * ud.view(-1, 1, 1)
    alpha = e / (O.seg_sum(e, idx, n_seg)[idx] + eps)
    ref = O.seg_sum(alpha * vd, idx, n_seg)
    w = torch.randn(ref.shape, generator=torch.Generator().manual_seed(1)).double()
    (ref * w).sum().backward()
    # kernel
    plan = graph.build_segment_plan(idx.to(DEV), n_seg)
    gc, vc = gate.to(DEV).requires_grad_(True), val.to(DEV).requires_grad_(True)
    uc = u.to(DEV).requires_grad_(True) if with_u else None
    out = ops.seg_softmax(gc, vc, plan, u=uc, eps=eps)
    (out * w.float().to(DEV)).sum().backward()
    assert_close(out.detach(), ref.detach(), "out", atol=1e-5, rtol=1e-5)
    assert_close(gc.grad, gd.grad, "d_gate", atol=1e-5, rtol=1e-4)
    assert_close(vc.grad, vd.grad, "d_value", atol=1e-5, rtol=1e-4)
    if with_u:
        assert_close(uc.grad, ud.grad, "d_u", atol=1e-4, rtol=1e-3)


def test_seg_softmax_deterministic_and_long_segments():
    idx, gate, val, _ = _softmax_case(11, 5, 128, 128, False, seed=3, max_len=3000)
    plan = graph.build_segment_plan(idx.to(DEV), 11)
    a = ops.seg_softmax(gate.to(DEV), val.to(DEV), plan)
    b = ops.seg_softmax(gate.to(DEV), val.to(DEV), plan)
    assert torch.equal(a, b)
    e = (gate.double() - O.seg_max(gate.double(), idx, 11)[idx]).exp()
    ref = O.seg_sum(e / (O.seg_sum(e, idx, 11)[idx] + 1e-16) * val.double(), idx, 11)
    assert_close(a, ref, "long segments", atol=1e-5, rtol=1e-5)


def test_library_loaded_and_counting():
    before = _lib.launch_count()
    graph.build_segment_plan(torch.zeros(4, dtype=torch.int64, device=DEV), 1)
    assert _lib.launch_count() > before


@pytest.mark.parametrize("n", [1, 130, 700, 3000])
def test_hyper_linear_fused_f256(n):
    """The F = 256 instantiation of the fused hyper-linear kernels (BASELINE.json configs[3]: hidden 256): forward,
    both activation gradients (two 128-column halves per predicted-weight row) and the weight gradient (one N = 256
    accumulator per (output channel, row half)) against fp64."""
    f = 256
    g = torch.Generator().manual_seed(n)
    z = torch.tanh(torch.randn(n, f, generator=g))
    y = torch.randn(n, f, generator=g)
    w = torch.randn(f * f + f, f, generator=g) * 0.0125 * 0.7
    b = torch.randn(f * f + f, generator=g) * 0.09
    zd, yd, wd, bd = (t.double().requires_grad_(True) for t in (z, y, w, b))
    p = zd @ wd.t() + bd
    ref = torch.einsum("noi,ni->no", p[:, : f * f].view(n, f, f), yd) + p[:, f * f:]
    gw = torch.randn(n, f, generator=g).double() * 1e-3          # a realistically small upstream gradient
    (ref * gw).sum().backward()
    zc, yc, wc, bc = (t.to(DEV).requires_grad_(True) for t in (z, y, w, b))
    before = _lib.launch_count()
    out = ops.hyper_linear(zc, wc, bc, yc, f)
    (out * gw.float().to(DEV)).sum().backward()
    assert _lib.launch_count() - before >= 6, "the fused F = 256 kernels did not run"
    assert_close(out.detach(), ref.detach(), "hyper fwd", atol=4e-5, rtol=1e-4)
    sc = gw.abs().max().item()
    assert_close(zc.grad, zd.grad, "g_z", atol=1e-4 * sc, rtol=1e-3)
    assert_close(yc.grad, yd.grad, "g_y", atol=1e-4 * sc, rtol=1e-3)
    assert_close(wc.grad, wd.grad, "g_w", atol=1e-5 * wd.grad.abs().max().item() + 1e-4 * sc, rtol=1e-3)
    assert_close(bc.grad, bd.grad, "g_b", atol=1e-5 * bd.grad.abs().max().item() + 1e-4 * sc, rtol=1e-3)


@pytest.mark.parametrize("f16", [True, False], ids=["f16x3", "tf32x3"])
@pytest.mark.parametrize("n", [1, 119, 128, 700, 5559])
def test_hyper_linear_fused_fwd_bwd(n, f16, monkeypatch):
    """cgat_hyper_rowdot_fwd[_f16] / cgat_hyper_rowscale[_f16] (+ e-term GEMM) against the reference arithmetic in
    fp64 (HyperLinear.forward + BatchLinear.forward, reference CGAT/Hypernetworksmp.py:243-254, 205-209), with the
    MMA operands as scaled fp16 hi/lo pairs and as tf32 hi/lo pairs: same tolerance for both."""
    monkeypatch.setattr(ops, "_F16X3", f16)
    f = 128
    g = torch.Generator().manual_seed(n)
    z = torch.tanh(torch.randn(n, f, generator=g))
    y = torch.randn(n, f, generator=g)
    w = torch.randn(f * f + f, f, generator=g) * 0.0125
    b = torch.randn(f * f + f, generator=g) * 0.09
    zd, yd, wd, bd = (t.double().requires_grad_(True) for t in (z, y, w, b))
    p = zd @ wd.t() + bd
    ref = torch.einsum("noi,ni->no", p[:, : f * f].view(n, f, f), yd) + p[:, f * f:]
    gw = torch.randn(n, f, generator=g).double()
    (ref * gw).sum().backward()
    zc, yc, wc, bc = (t.to(DEV).requires_grad_(True) for t in (z, y, w, b))
    out = ops.hyper_linear(zc, wc, bc, yc, f)
    (out * gw.float().to(DEV)).sum().backward()
    assert_close(out.detach(), ref.detach(), "hyper fwd", atol=2e-5, rtol=1e-4)
    assert_close(zc.grad, zd.grad, "g_z", atol=1e-4, rtol=1e-3)
    assert_close(yc.grad, yd.grad, "g_y", atol=1e-4, rtol=1e-3)
    # sums over all n atoms of O(1) terms: the noise floor scales with the magnitude of the sums
    assert_close(wc.grad, wd.grad, "g_w", atol=1e-5 * wd.grad.abs().max().item() + 1e-4, rtol=1e-3)
    assert_close(bc.grad, bd.grad, "g_b", atol=1e-5 * bd.grad.abs().max().item() + 1e-4, rtol=1e-3)


def test_hyper_linear_work_splits_agree():
    """The f16 hyper kernels choose between two work decompositions (hyper_f16.cu): CTAs tied to one atom tile with a
    slice of the output channels when there are enough SMs per tile, contiguous (tile, chunk) item ranges otherwise
    (more than ~72 tile pairs: the 5 000-crystal inference batches).  The op is independent per atom row, so 20 000 atoms
    in one call (item ranges) must reproduce the two 10 000-atom halves (tile-aligned): forward bit for bit — every row
    sees the same MMA sequence either way — and the gradients to rounding (partial sums meet in a different order)."""
    f, n = 128, 20000
    g = torch.Generator().manual_seed(7)
    z = torch.tanh(torch.randn(n, f, generator=g)).to(DEV)
    y = torch.randn(n, f, generator=g).to(DEV)
    w = (torch.randn(f * f + f, f, generator=g) * 0.0125).to(DEV)
    b = (torch.randn(f * f + f, generator=g) * 0.09).to(DEV)
    gw = torch.randn(n, f, generator=g).to(DEV)

    def run(lo, hi):
        zc, yc = z[lo:hi].clone().requires_grad_(True), y[lo:hi].clone().requires_grad_(True)
        wc, bc = w.clone().requires_grad_(True), b.clone().requires_grad_(True)
        out = ops.hyper_linear(zc, wc, bc, yc, f)
        (out * gw[lo:hi]).sum().backward()
        return out.detach(), zc.grad, yc.grad, wc.grad, bc.grad

    whole = run(0, n)
    a, c = run(0, n // 2), run(n // 2, n)
    assert torch.equal(whole[0], torch.cat([a[0], c[0]])), "forward differs between the two work splits"
    for i, name in ((1, "g_z"), (2, "g_y")):
        assert_close(whole[i], torch.cat([a[i], c[i]]).double(), name, atol=2e-5, rtol=1e-5)
    for i, name in ((3, "g_w"), (4, "g_b")):
        ref = (a[i] + c[i]).double()
        assert_close(whole[i], ref, name, atol=1e-5 * ref.abs().max().item(), rtol=1e-4)


def test_pack_repack_on_weight_update():
    f = 128
    w = torch.randn(f * f + f, f, device=DEV) * 0.01
    p1 = ops.packed_kmajor(w, f * f).clone()
    assert ops.packed_kmajor(w, f * f).data_ptr() == ops.packed_kmajor(w, f * f).data_ptr()
    w.add_(1.0)
    p2 = ops.packed_kmajor(w, f * f)
    assert not torch.equal(p1, p2)
    # a different tensor that lands on the same address must not see the old packing
    ptr = w.data_ptr()
    del w, p2
    w2 = torch.randn(f * f + f, f, device=DEV) * 0.01
    p3 = ops.packed_kmajor(w2, f * f)
    assert w2.data_ptr() != ptr or not torch.equal(p1, p3)


@pytest.mark.parametrize("n_cry,k,heads,lo,hi", [(3, 12, 5, 2, 20), (40, 12, 5, 2, 20), (2, 24, 2, 200, 256),
                                                 (700, 12, 5, 2, 20), (1, 4, 1, 2, 3)])
@pytest.mark.parametrize("f16", [True, False], ids=["f16x3", "tf32x3"])
def test_edge_attention_fused_matches_oracle(n_cry, k, heads, lo, hi, f16, monkeypatch):
    """cgat_edge_attn_fwd[_f16] (+ projection GEMMs) against the reference arithmetic of
    GATConvNodes.message/aggregate (reference CGAT/CGAT.py:319-329) restated in fp64; both operand formats."""
    from cgat_b200.CGAT import MultiHeadNetwork
    monkeypatch.setattr(ops, "_F16X3_EDGE", f16)
    f, fe = 128, 128
    sb = synthetic.make_batch(n_cry, k, seed=n_cry, atoms_lo=lo, atoms_hi=hi)
    gidx = sb.graph
    n = gidx.num_nodes
    torch.manual_seed(n_cry)
    width = 2 * f + fe
    mh_a = MultiHeadNetwork(width, f, int(width / 1.5), heads)
    mh_m = MultiHeadNetwork(width, f, int(width / 1.5), heads)
    x = torch.randn(n, f) * 0.5
    tab = torch.randn(k + 1, fe)
    # fp64 oracle
    sd = {}
    for pre, mod in (("A.", mh_a), ("M.", mh_m)):
        for kk, v in mod.state_dict().items():
            sd[pre + kk] = v.double()
    src, dst = gidx.edge_index
    m = torch.cat([x.double()[dst], tab.double()[gidx.edge_attr], x.double()[src]], dim=1)
    alpha = O.pyg_softmax(O.multi_head_network(sd, "A.", m, heads), dst, n)
    ref = O.seg_sum(O.multi_head_network(sd, "M.", m, heads) * alpha, dst, n).mean(dim=1)
    # fused path
    plan = graph.build_edge_plan(gidx.edge_index.to(DEV), gidx.edge_attr.to(DEV), n)
    mh_a, mh_m = mh_a.to(DEV), mh_m.to(DEV)
    xc = x.to(DEV).requires_grad_(True)
    out = ops.edge_attention(xc, tab.to(DEV), plan, mh_a, mh_m, heads)
    assert_close(out.detach(), ref, "edge attention fwd", atol=2e-5, rtol=1e-4)
    out2 = ops.edge_attention(xc, tab.to(DEV), plan, mh_a, mh_m, heads)
    assert torch.equal(out, out2), "fused edge attention is not deterministic"
    # backward runs (recompute path) and matches the oracle's gradient w.r.t. x
    xd = x.double().requires_grad_(True)
    m = torch.cat([xd[dst], tab.double()[gidx.edge_attr], xd[src]], dim=1)
    alpha = O.pyg_softmax(O.multi_head_network(sd, "A.", m, heads), dst, n)
    refg = O.seg_sum(O.multi_head_network(sd, "M.", m, heads) * alpha, dst, n).mean(dim=1)
    w = torch.randn(n, f, generator=torch.Generator().manual_seed(5)).double()
    (refg * w).sum().backward()
    (out * w.float().to(DEV)).sum().backward()
    from tests._cases import assert_grad_close
    assert_grad_close(xc.grad, xd.grad, "edge attention d_x")


@pytest.mark.parametrize("n_cry,k,heads,lo,hi", [(3, 12, 5, 2, 20), (60, 12, 5, 2, 20), (2, 24, 2, 200, 256),
                                                 (500, 12, 5, 2, 20)])
@pytest.mark.parametrize("f16", [True, False], ids=["f16x3", "tf32x3"])
def test_edge_attention_fused_backward_all_grads(n_cry, k, heads, lo, hi, f16, monkeypatch):
    """Fused edge-attention backward (bwd_prep + dgrad x2 + wgrad + first-layer GEMMs): every gradient
    against autograd through the fp64 restatement of GATConvNodes.message/aggregate."""
    monkeypatch.setattr(ops, "_F16X3_EDGE", f16)
    _edge_backward_case(n_cry, k, heads, lo, hi, 128)


@pytest.mark.parametrize("n_cry,k,heads,lo,hi,f", [(40, 12, 8, 2, 20, 256), (3, 24, 8, 100, 120, 256),
                                                   (30, 12, 3, 2, 20, 256), (30, 12, 5, 2, 20, 128)])
def test_edge_attention_fused_wide(n_cry, k, heads, lo, hi, f):
    """The generalised f16 path: F = 256 as two virtual heads per head and a hidden width that is not a multiple of 64
    (F = 256 -> Hd = int(640 / 1.5) = 426, zero-padded to 448) — BASELINE.json configs[3] (hidden 256, 8 heads) —
    forward and every gradient against the fp64 restatement; the launch counter proves the fused kernels ran."""
    before = _lib.launch_count()
    _edge_backward_case(n_cry, k, heads, lo, hi, f, kink_tolerant_table=True)
    assert _lib.launch_count() - before >= 10


@pytest.mark.parametrize("n_cry,k,heads,f", [(40, 12, 5, 128), (25, 12, 4, 256)])
def test_edge_attention_fused_scalar_gate(n_cry, k, heads, f):
    """vector_attention=False (reference CGAT/CGAT.py:282-287: one gate logit per head) on the fused kernels: the gate
    row is repeated over the channels through an autograd `expand`, so the kernels run unchanged and the per-channel
    gate gradients are summed back onto the single row.  Forward and every gradient against the fp64 restatement."""
    before = _lib.launch_count()
    _edge_backward_case(n_cry, k, heads, 2, 20, f, kink_tolerant_table=True, scalar_gate=True)
    assert _lib.launch_count() - before >= 10


def _edge_backward_case(n_cry, k, heads, lo, hi, f, kink_tolerant_table=False, scalar_gate=False):
    from cgat_b200.CGAT import MultiHeadNetwork
    from tests._cases import assert_grad_close
    fe = 128
    sb = synthetic.make_batch(n_cry, k, seed=100 + n_cry, atoms_lo=lo, atoms_hi=hi)
    gidx = sb.graph
    n = gidx.num_nodes
    torch.manual_seed(n_cry)
    width = 2 * f + fe
    mh_a = MultiHeadNetwork(width, 1 if scalar_gate else f, int(width / 1.5), heads)
    mh_m = MultiHeadNetwork(width, f, int(width / 1.5), heads)
    x = torch.randn(n, f) * 0.5
    tab = torch.randn(k + 1, fe)
    w = torch.randn(n, f, generator=torch.Generator().manual_seed(5))
    # fp64 oracle with autograd
    sd = {}
    for pre, mod in (("A.", mh_a), ("M.", mh_m)):
        for kk, v in mod.state_dict().items():
            sd[pre + kk] = v.double().requires_grad_(True)
    xd, tabd = x.double().requires_grad_(True), tab.double().requires_grad_(True)
    src, dst = gidx.edge_index
    m = torch.cat([xd[dst], tabd[gidx.edge_attr], xd[src]], dim=1)
    alpha = O.pyg_softmax(O.multi_head_network(sd, "A.", m, heads), dst, n)
    ref = O.seg_sum(O.multi_head_network(sd, "M.", m, heads) * alpha, dst, n).mean(dim=1)
    (ref * w.double()).sum().backward()
    # fused
    plan = graph.build_edge_plan(gidx.edge_index.to(DEV), gidx.edge_attr.to(DEV), n)
    mh_a, mh_m = mh_a.to(DEV), mh_m.to(DEV)
    xc, tabc = x.to(DEV).requires_grad_(True), tab.to(DEV).requires_grad_(True)
    out = ops.edge_attention(xc, tabc, plan, mh_a, mh_m, heads)
    (out * w.to(DEV)).sum().backward()
    assert_close(out.detach(), ref.detach(), "fwd", atol=2e-5, rtol=1e-4)
    assert_grad_close(xc.grad, xd.grad, "d_x")
    # d_table sums E * 2*H*Hd signed terms per entry (heavy cancellation): calibrate the fp32 noise floor with the
    # library-GEMM formulation of the same op and require the fused path to sit at that floor
    xu, tabu = x.to(DEV).requires_grad_(True), tab.to(DEV).requires_grad_(True)
    outu = ops.edge_attention_unfused(xu, tabu, plan, mh_a.w_in(), mh_a.fc_in.bias, mh_a.w_out(), mh_a.fc_out.bias,
                                      mh_m.w_in(), mh_m.fc_in.bias, mh_m.w_out(), mh_m.fc_out.bias, heads)
    (gu_x, gu_tab) = torch.autograd.grad((outu * w.to(DEV)).sum(), [xu, tabu])
    floor = (gu_tab.cpu().double() - tabd.grad).abs().max().item()
    err = (tabc.grad.cpu().double() - tabd.grad).abs().max().item()
    print(f"d_table: fused err {err:.3e}, library-fp32 err {floor:.3e}, ref max {tabd.grad.abs().max().item():.3e}")
    if kink_tolerant_table:
        # at 8 heads x 426 hidden units a LeakyReLU pre-activation within fp32 rounding of 0 is likely in ONE of the two
        # evaluation orders (P[dst] + P[src] + T[rank] here, one GEMM over the concatenated input in the library):
        # scripts/edge_wide_diag.py shows the fused and library errors identical to three digits when neither flips,
        # and a single hidden unit's row off when one does — so the table gradient is held to the kink-aware criterion
        assert_grad_close(tabc.grad, tabd.grad, "d_table")
    else:
        assert err <= max(4 * floor, 1e-4 + 1e-3 * tabd.grad.abs().max().item() * 0.1), (err, floor)
    for pre, mod in (("A.", mh_a), ("M.", mh_m)):
        for kk, p in mod.named_parameters():
            ref_g = sd[pre + kk].grad
            scale = max(ref_g.abs().max().item(), 1e-3)
            # a scalar gate's hidden units carry the gradient of all 128 channels at once: when one pre-activation sits
            # on the LeakyReLU kink, that unit's bias / weight-row gradient moves by a larger share of the maximum
            assert_grad_close(p.grad, ref_g, f"d_{pre}{kk}", atol=1e-4 + 1e-5 * scale,
                              outlier_cap=5e-2 if scalar_gate else 2e-2)
    # deterministic
    xc2 = x.to(DEV).requires_grad_(True)
    for p in list(mh_a.parameters()) + list(mh_m.parameters()):
        p.grad = None
    out2 = ops.edge_attention(xc2, tabc, plan, mh_a, mh_m, heads)
    (out2 * w.to(DEV)).sum().backward()
    assert torch.equal(xc2.grad, xc.grad), "fused backward is not deterministic"


def _trunk_case(n, n_j, seed):
    f = 128
    g = torch.Generator().manual_seed(seed)
    h = torch.randn(n, f, generator=g) * 0.5
    layers, tails = [], []
    for _ in range(n_j):
        layers.append([(torch.randn(f, f, generator=g) * (2.0 / f) ** 0.5, torch.randn(f, generator=g) * 0.1)
                       for _ in range(4)])
        tails.append((torch.randn(f * f + f, f, generator=g) * 0.0125, torch.randn(f * f + f, generator=g) * 0.09,
                      f * f))
    return h, layers, tails


@pytest.mark.parametrize("n,n_j", [(1, 1), (119, 4), (128, 3), (700, 4), (5559, 4)])
def test_hyper_trunks_fused_fwd_bwd(n, n_j):
    """cgat_hyper_trunk_fwd / _bwd + cgat_gemm3x_tn_batched against the reference arithmetic in fp64
    (FCBlock = [Linear+Tanh]x4, reference CGAT/Hypernetworksmp.py:36-83; bias tail of HyperLinear :243-254)."""
    f = 128
    h, layers, tails = _trunk_case(n, n_j, 100 + n)
    g = torch.Generator().manual_seed(n)
    gz = [torch.randn(n, f, generator=g) for _ in range(n_j)]
    ge = [torch.randn(n, f, generator=g) for _ in range(n_j)]

    def run(dev, dt):
        hh = h.to(dev, dt).requires_grad_(True)
        ll = [[(w.to(dev, dt).requires_grad_(True), b.to(dev, dt).requires_grad_(True)) for w, b in l] for l in layers]
        tt = [(w.to(dev, dt), b.to(dev, dt), r) for w, b, r in tails]
        zs, es = ops.hyper_trunks(hh, ll, tt)
        if es[0] is None:  # unfused formulation: the tail is applied by the caller
            es = [z @ w[r:].t() + b[r:] for z, (w, b, r) in zip(zs, tt)]
        loss = sum((z * a.to(dev, dt)).sum() + (e * c.to(dev, dt)).sum() for z, e, a, c in zip(zs, es, gz, ge))
        loss.backward()
        return zs, es, hh.grad, [[(w.grad, b.grad) for w, b in l] for l in ll]

    zs, es, gh, gl = run(DEV, torch.float32)
    zr, er, ghr, glr = run("cpu", torch.float64)
    for j in range(n_j):
        assert_close(zs[j].detach(), zr[j].detach(), f"z[{j}]", atol=1e-5, rtol=1e-5)
        assert_close(es[j].detach(), er[j].detach(), f"e[{j}]", atol=1e-5, rtol=1e-5)
        for s in range(4):
            scale = glr[j][s][0].abs().max().item()
            assert_close(gl[j][s][0], glr[j][s][0], f"g_w[{j}][{s}]", atol=1e-5 * scale + 1e-5, rtol=1e-4)
            assert_close(gl[j][s][1], glr[j][s][1], f"g_b[{j}][{s}]", atol=1e-5 * scale + 1e-5, rtol=1e-4)
    assert_close(gh, ghr, "g_h", atol=1e-5 * ghr.abs().max().item() + 1e-5, rtol=1e-4)


def test_hyper_trunks_repack_on_weight_update():
    h, layers, tails = _trunk_case(200, 2, 7)
    hh = h.to(DEV)
    ll = [[(w.to(DEV), b.to(DEV)) for w, b in l] for l in layers]
    tt = [(w.to(DEV), b.to(DEV), r) for w, b, r in tails]
    with torch.no_grad():
        z1, _ = ops.hyper_trunks(hh, ll, tt)
        z1 = [z.clone() for z in z1]
        ll[1][2][0].mul_(0.5)
        z2, _ = ops.hyper_trunks(hh, ll, tt)
    assert torch.equal(z1[0], z2[0]) and not torch.equal(z1[1], z2[1])


@pytest.mark.parametrize("n_parts,n,acc", [(1, 1000, False), (5, 4096, False), (18, 128 * 130 + 3, False), (8, 5632 * 128, True),
                                           (32, 64, True)])
def test_sum_parts_matches_fp64(n_parts, n, acc):
    """cgat_sum_parts: fixed-order sum of split-K / split-atom partial results (vector and scalar paths)."""
    g = torch.Generator().manual_seed(n_parts * 7 + n)
    parts = torch.randn(n_parts, n, generator=g).to(DEV)
    base = torch.randn(n, generator=g).to(DEV)
    out = base.clone() if acc else None
    res = ops.sum_parts(parts, out=out, accumulate=acc)
    ref = parts.double().sum(0) + (base.double() if acc else 0)
    assert (res.double() - ref).abs().max().item() <= 1e-6 * (1 + n_parts ** 0.5) * 4
    res2 = ops.sum_parts(parts, out=base.clone() if acc else None, accumulate=acc)
    assert torch.equal(res, res2), "sum_parts is not deterministic"


def test_adamw_flat_matches_torch_adamw():
    """cgat_adamw_flat against torch.optim.AdamW (the reference's optimizer, CGAT/lightning_module.py:328-344) over
    four steps with changing gradients and a learning-rate change, incl. the folded 1/world gradient scale."""
    from cgat_b200 import _lib
    n = 4 * 70001
    g = torch.Generator().manual_seed(3)
    p0 = torch.randn(n, generator=g)
    grads = [torch.randn(n, generator=g) * (10.0 ** (i - 2)) for i in range(4)]
    ref = torch.nn.Parameter(p0.clone().double())
    opt = torch.optim.AdamW([ref], lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2)
    p = p0.clone().to(DEV)
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    step, lr = torch.zeros(1, device=DEV), torch.full((1,), 1e-3, device=DEV)
    for i, gr in enumerate(grads):
        if i == 2:
            lr.fill_(5e-4)
            opt.param_groups[0]["lr"] = 5e-4
        ref.grad = gr.double() / 2
        opt.step()
        gd = gr.to(DEV)
        _lib.call("cgat_adamw_flat", _lib.ptr(p), _lib.ptr(gd), _lib.ptr(m), _lib.ptr(v), n, _lib.ptr(lr),
                  _lib.ptr(step), 0.9, 0.999, 1e-8, 1e-2, 0.5, _lib.stream())
        err = (p.double().cpu() - ref.detach()).abs().max().item()
        assert err < 2e-6, f"step {i}: {err}"
    assert float(step) == 4.0


@pytest.mark.parametrize("n,rows", [(1, 1), (500, 501), (5000, 5001), (37, 37)])
def test_l1_loss_fwd_bwd(n, rows):
    from cgat_b200 import optim
    g = torch.Generator().manual_seed(n)
    out = torch.randn(rows, 2, generator=g).to(DEV).requires_grad_(True)
    tgt = torch.randn(rows, 1, generator=g).to(DEV)
    loss = optim.l1_loss(out, tgt, n)
    (3.0 * loss).backward()
    od = out.detach().double().cpu().requires_grad_(True)
    ref = (od[:n, :1] - tgt.double().cpu()[:n]).abs().mean()
    (3.0 * ref).backward()
    assert abs(float(loss) - float(ref)) < 1e-6
    assert (out.grad.double().cpu() - od.grad).abs().max().item() < 1e-7
    # the slice form graphed.GraphedTrainStep uses: pred = out[:n, :1] (row stride 2)
    out2 = out.detach().clone().requires_grad_(True)
    optim.l1_loss(out2[:n, :1], tgt[:n]).backward()
    assert (out2.grad.double().cpu() * 3 - od.grad).abs().max().item() < 1e-6


@pytest.mark.parametrize("n_store,n_sel,k,seed", [(60, 25, 12, 0), (300, 300, 12, 1), (8, 40, 24, 2), (40, 1, 6, 3)])
def test_device_collation_bit_exact(n_store, n_sel, k, seed):
    """store.CrystalStore.collate (cgat_collate_plan / cgat_collate_fill) against the host path the reference uses per
    batch — per-crystal samples -> Batch.from_data_list / collate_batch semantics (batching.collate) -> pad_batch —
    for a random selection (repeats and arbitrary order allowed): every tensor identical, integers and floats."""
    import numpy as np
    from cgat_b200 import batching, store, synthetic
    sb = synthetic.make_batch(n_store, k, seed=seed)
    st = store.CrystalStore.from_batch(sb).to(DEV)
    rng = np.random.default_rng(seed)
    sel = rng.integers(0, n_store, size=n_sel)
    dev = st.collate(sel)
    samples = []
    for c in sel:
        part = synthetic.split_batch(sb, int(c), int(c) + 1)
        samples.append((part.graph, part.roost[:4]))
    host = batching.pad_batch(batching.collate(samples))
    assert dev.graph.num_graphs == host.graph.num_graphs == n_sel + 1
    for name, a, b in zip(("x", "edge_index", "edge_attr", "batch", "y"), dev.graph.tensors(), host.graph.tensors()):
        assert a.dtype == b.dtype and a.shape == b.shape, (name, a.dtype, b.dtype, a.shape, b.shape)
        assert torch.equal(a.cpu(), b), name
    for name, a, b in zip(("weights", "fea", "self_idx", "nbr_idx", "crystal_idx"), dev.roost, host.roost):
        assert a.dtype == b.dtype and a.shape == b.shape, (name, a.dtype, b.dtype, a.shape, b.shape)
        assert torch.equal(a.cpu(), b), name
    assert list(dev.n_atoms) == list(host.n_atoms)
    # and the model sees the same thing
    assert batching.signature(dev) == batching.signature(host)


def test_status_flags_for_out_of_range_input():
    """What the reference's nn.Embedding / index_select refuse (a shell rank beyond the table, a node id beyond the
    batch) raises a sticky device flag instead of gathering out of bounds; clean input leaves the flags clear."""
    _lib.status_flags(reset=True)
    sb = synthetic.make_batch(5, 12, seed=0)
    g = sb.graph
    n = g.num_nodes
    plan = graph.build_edge_plan(g.edge_index.to(DEV), g.edge_attr.to(DEV), n, 13)
    assert _lib.status_flags(reset=True) == 0
    assert int(plan.rank.max()) <= 12
    bad_attr = g.edge_attr.clone()
    bad_attr[3] = 13                                        # one past the (K+1)-row table
    plan = graph.build_edge_plan(g.edge_index.to(DEV), bad_attr.to(DEV), n, 13)
    assert _lib.status_flags(reset=False) & 4
    assert int(plan.rank.max()) <= 12                       # clamped: nothing downstream reads out of bounds
    with pytest.raises(_lib.CgatLibraryError, match="shell rank"):
        _lib.check_status()
    assert _lib.status_flags() == 0                         # check_status cleared them
    bad_ei = g.edge_index.clone()
    bad_ei[0, 7] = n + 5
    bad_ei[1, 9] = -1
    plan = graph.build_edge_plan(bad_ei.to(DEV), g.edge_attr.to(DEV), n, 13)
    flags = _lib.status_flags(reset=True)
    assert flags & 1 and flags & 2 and not flags & 4
    assert int(plan.src.max()) < n and int(plan.rowptr[-1]) == g.edge_index.shape[1] - 1   # the bad edge was dropped
