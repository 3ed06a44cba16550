"""-m gpu: CGAtNet on the B200 through the C ABI vs (a) the reference goldens and (b) the CPU oracle,
forward and every gradient, within the north-star tolerance (1e-4 abs / 1e-3 rel), plus
size-independent properties at the bench workload's size."""
import numpy as np
import pytest
import torch

import cgat_b200
from cgat_b200 import _lib, synthetic, weights
from oracle import cgat_oracle as O
from tests._cases import (ATOL, CASES, RTOL, assert_close, assert_grad_close, golden_shapes, grad_digest, grad_stats,
                          load_golden, oracle_cfg, record_parity, training_scalar)

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _gpu_model(name):
    mkw, bkw, wseed = CASES[name]
    model = weights.load_seeded(cgat_b200.CGAtNet(200, **mkw), wseed).to(DEV)
    return model, synthetic.make_batch(**bkw)


@pytest.mark.parametrize("name", list(CASES))
def test_forward_matches_reference_golden(name, golden_dir):
    gold = load_golden(golden_dir, name)
    model, sb = _gpu_model(name)
    d = sb.to(DEV)
    before = _lib.launch_count()
    with torch.no_grad():
        out = model(d.graph, (t for t in d.roost))
        emb = model(d.graph, d.roost, return_graph_embedding=True)
        pen = model(d.graph, d.roost, last_layer=False)
    assert _lib.launch_count() > before, "no cgat_b200 kernel was launched"
    assert_close(out, gold["out"], f"{name}: out")
    assert_close(emb, gold["embedding"], f"{name}: embedding")
    assert_close(pen, gold["penultimate"], f"{name}: penultimate")



def _library_fp32_grads(mkw, wseed, bkw):
    """The same model on the same device with every fused tensor-core kernel switched off (ops._FUSED = False:
    PyTorch's own fp32 GEMMs + this library's segment kernels) — the measured fp32 noise floor of this network on
    this GPU, LeakyReLU kink flips included.  Returns {name: grad}."""
    from cgat_b200 import ops
    was = ops._FUSED
    ops._FUSED = False
    try:
        model = weights.load_seeded(cgat_b200.CGAtNet(200, **mkw), wseed).to(DEV)
        d = synthetic.make_batch(**bkw).to(DEV)
        training_scalar(model(d.graph, d.roost), d.graph.y).backward()
        return {k: p.grad for k, p in model.named_parameters()}
    finally:
        ops._FUSED = was


def _check_strict_parity(name, grads, ref_grads, lib_grads):
    """North-star criterion (1e-4 abs + 1e-3 rel per element) on EVERY gradient element, recorded to
    gpurun_out/grad_parity_<name>.json.  The only allowance is measured, not assumed: PyTorch's own fp32 path on the
    same device is held to the same strict count, and the fused path may not exceed it by more than a handful —
    a LeakyReLU pre-activation within fp32 rounding of 0 flips side under ANY fp32 evaluation order and moves one
    hidden unit's weight-gradient row (default_k12: the same 142 elements of graphs.4.Node.MH_M.fc_in.weight are out
    for the fused kernels, the tf32 kernels and the library path alike; profiles/r02a_graddiag_*)."""
    stats, lib_stats = {}, {}
    for k, g in grads.items():
        if ref_grads[k] is None:
            assert g is None or float(g.abs().max()) == 0.0, f"{k}: dead parameter got a gradient"
            continue
        assert g is not None, f"{k}: missing gradient"
        stats[k] = grad_stats(g, ref_grads[k])
        lib_stats[k] = grad_stats(lib_grads[k], ref_grads[k])
    lib_tot = dict(bad=sum(v["bad"] for v in lib_stats.values()), far=sum(v["far"] for v in lib_stats.values()),
                   max_ratio=max(v["max_ratio"] for v in lib_stats.values()),
                   tensors_with_violations={k: v for k, v in lib_stats.items() if v["bad"]})
    tot = record_parity(name, stats, extra=dict(library_fp32_same_device=lib_tot))
    print(f"{name}: strict violations fused {tot}; library fp32 on the same device {lib_tot['bad']} / far "
          f"{lib_tot['far']} / max ratio {lib_tot['max_ratio']:.2f}")
    assert tot["bad"] <= 8 + 1.5 * lib_tot["bad"], (name, tot, lib_tot["bad"])
    assert tot["far"] <= 2 + lib_tot["far"], (name, tot, lib_tot["far"])
    assert tot["max_ratio"] <= max(3.0, 1.25 * lib_tot["max_ratio"]), (name, tot, lib_tot["max_ratio"])
    # scalar-free guard against a wrong (not merely noisy) tensor: every tensor's relative L2 error is small
    for k, st in stats.items():
        assert st["bad"] == 0 or st["rel_l2"] <= max(2e-3, 4 * lib_stats[k]["rel_l2"]), (name, k, st)


@pytest.mark.parametrize("name", list(CASES))
def test_all_gradients_match_oracle(name, golden_dir):
    """Every parameter gradient against the CPU oracle run here on the same seeded inputs, and the
    committed reference digests / full tensors."""
    mkw, bkw, wseed = CASES[name]
    gold = load_golden(golden_dir, name)
    model, sb = _gpu_model(name)
    d = sb.to(DEV)
    out = model(d.graph, d.roost)
    training_scalar(out, d.graph.y).backward()
    # the oracle in fp64 is the noise-free statement of the reference's arithmetic
    sd = weights.seeded_state_dict(golden_shapes(gold), wseed, torch.float64)
    for v in sd.values():
        v.requires_grad_(True)
    sb64 = synthetic.make_batch(dtype=torch.float64, **bkw)
    ref_out = O.cgat_forward(sd, oracle_cfg(mkw), sb64.graph, sb64.roost)
    training_scalar(ref_out, sb64.graph.y).backward()
    assert_close(out.detach(), ref_out.detach(), f"{name}: out vs oracle")
    none_ref = set(map(str, gold["none_grads"]))
    for k, p in model.named_parameters():
        if k in none_ref:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, f"{k}: dead parameter got a gradient"
    _check_strict_parity(name, {k: p.grad for k, p in model.named_parameters()}, {k: v.grad for k, v in sd.items()},
                         _library_fp32_grads(mkw, wseed, bkw))
    # the committed full gradient tensors of the UNMODIFIED reference (fp32 CPU run): strict per element
    for key in gold.files:
        if key.startswith("grad::"):
            st = grad_stats(dict(model.named_parameters())[key[6:]].grad, gold[key])
            assert st["bad"] <= max(1, 1e-3 * st["n"]) and st["far"] == 0, (name, key, st)


def _chunked_oracle_grads(mkw, wseed, shapes, bkw, chunk):
    """fp64 oracle gradients of training_scalar over a LARGE batch, accumulated over chunks of crystals: crystals are
    independent and the scalar is a mean over crystals with one global constant (max|y|), so the sum of the chunk
    gradients is the gradient of the whole batch while the CPU memory stays that of one chunk."""
    sd = weights.seeded_state_dict(shapes, wseed, torch.float64)
    for v in sd.values():
        v.requires_grad_(True)
    sb64 = synthetic.make_batch(dtype=torch.float64, **bkw)
    C = sb64.num_crystals
    ymax = sb64.graph.y.abs().max()
    outs = []
    for lo in range(0, C, chunk):
        part = synthetic.split_batch(sb64, lo, min(C, lo + chunk))
        out = O.cgat_forward(sd, oracle_cfg(mkw), part.graph, part.roost)
        y = part.graph.y.view(-1, 1) / ymax
        (((out[:, :1] - y).abs().sum() + 0.1 * out[:, 1].sum()) / C).backward()
        outs.append(out.detach())
    return sd, torch.cat(outs)


BIG_CASES = {
    # BASELINE.json configs[1] at its full size: 500 crystals, K = 12, default net (the bench workload)
    "cfg2_500_crystals": (CASES["default_k12"][0], dict(n_crystals=500, max_nbr=12, seed=1), 0, 50),
    # configs[4] on the FUSED kernels (F = 128; the golden large_cell_k24 case is F = 32 = library fallback):
    # 200-256-atom cells, 24 neighbours, long softmax segments
    "large_cell_k24_f128": (dict(CASES["default_k12"][0], neighbor_number=24),
                            dict(n_crystals=3, max_nbr=24, seed=4, atoms_lo=200, atoms_hi=256), 4, 1),
}


@pytest.mark.parametrize("name", list(BIG_CASES))
def test_full_size_gradients_match_oracle(name):
    """Predictions and every parameter gradient at the BASELINE sizes against the fp64 CPU oracle (chunked over
    crystals), under the same criterion as the golden cases; strict violation counts are recorded."""
    mkw, bkw, wseed, chunk = BIG_CASES[name]
    model = weights.load_seeded(cgat_b200.CGAtNet(200, **mkw), wseed).to(DEV)
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    d = synthetic.make_batch(**bkw).to(DEV)
    out = model(d.graph, d.roost)
    training_scalar(out, d.graph.y).backward()
    sd, ref_out = _chunked_oracle_grads(mkw, wseed, shapes, bkw, chunk)
    assert_close(out.detach(), ref_out, f"{name}: out vs oracle")
    _check_strict_parity(name, {k: p.grad for k, p in model.named_parameters()}, {k: v.grad for k, v in sd.items()},
                         _library_fp32_grads(mkw, wseed, bkw))


def test_bench_size_properties():
    """cfg2 size (500 crystals, K=12, default net): determinism, batch-split equivalence
    f(A ∪ B) = cat(f(A), f(B)), and crystal-order equivariance."""
    mkw = CASES["default_k12"][0]
    model = weights.load_seeded(cgat_b200.CGAtNet(200, **mkw), 0).to(DEV)
    sb = synthetic.make_batch(500, 12, seed=1)
    d = sb.to(DEV)
    with torch.no_grad():
        full = model(d.graph, d.roost)
        again = model(d.graph, d.roost)
        assert torch.equal(full, again), "forward is not deterministic"
        cut = 211
        a, b = synthetic.split_batch(sb, 0, cut).to(DEV), synthetic.split_batch(sb, cut, 500).to(DEV)
        parts = torch.cat([model(a.graph, a.roost), model(b.graph, b.roost)])
        assert_close(parts, full, "batch split", atol=2e-5, rtol=1e-4)
        swapped = torch.cat([model(b.graph, b.roost), model(a.graph, a.roost)])
        assert_close(swapped, torch.cat([full[cut:], full[:cut]]), "crystal order", atol=2e-5, rtol=1e-4)
    assert torch.isfinite(full).all()


def test_edge_order_invariance():
    """Permuting the neighbour slots of each atom (edge order within a source) changes nothing but
    fp32 summation order: the softmax segments are sets."""
    mkw, bkw, wseed = CASES["mixed_flags"]
    model, sb = _gpu_model("mixed_flags")
    g = sb.graph
    K = bkw["max_nbr"]
    n = g.num_nodes
    perm = torch.stack([torch.randperm(K, generator=torch.Generator().manual_seed(i)) for i in range(n)])
    flat = (torch.arange(n).view(-1, 1) * K + perm).reshape(-1)
    g2 = synthetic.GraphBatch(g.x, g.edge_index[:, flat], g.edge_attr[flat], g.batch, g.y)
    with torch.no_grad():
        a = model(g.to(DEV), tuple(t.to(DEV) for t in sb.roost))
        b = model(g2.to(DEV), tuple(t.to(DEV) for t in sb.roost))
    assert_close(a, b, "edge order", atol=2e-5, rtol=1e-4)


def test_padded_batch_matches_unpadded():
    """batching.pad_batch on the device path: predictions of the real crystals are unchanged by the dummy crystal."""
    from cgat_b200 import batching
    model, sb = _gpu_model("default_k12")
    pb = batching.pad_batch(sb)
    C = sb.num_crystals
    with torch.no_grad():
        ref = model(sb.graph.to(DEV), tuple(t.to(DEV) for t in sb.roost))
        pad = model(pb.graph.to(DEV), tuple(t.to(DEV) for t in pb.roost))
    assert pad.shape[0] == C + 1 and torch.isfinite(pad).all()
    assert_close(pad[:C], ref, "padded forward", atol=2e-5, rtol=1e-4)


def _train_setup(seed=0):
    mkw = CASES["default_k12"][0]
    model = weights.load_seeded(cgat_b200.CGAtNet(200, **mkw), seed).to(DEV)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3, weight_decay=1e-6, fused=True, capturable=True)
    return model, opt


def test_fused_optimizer_updates_reach_the_packed_operands():
    """Three eager train steps with AdamW(fused=True) — which does not bump parameter version counters — follow the
    same trajectory as AdamW(foreach=True), i.e. the pre-packed tensor-core operands are rebuilt after every step."""
    from cgat_b200 import distributed as cdist
    crit = torch.nn.L1Loss()
    batches = [synthetic.make_batch(40, 12, seed=s).to(DEV) for s in (1, 2, 3)]
    losses = {}
    for kind in ("fused", "foreach"):
        mkw = CASES["default_k12"][0]
        model = weights.load_seeded(cgat_b200.CGAtNet(200, **mkw), 0).to(DEV)
        opt = torch.optim.AdamW(model.parameters(), lr=1e-3, weight_decay=1e-6, **{kind: True})
        sync = cdist.GradSync(model, 1)
        losses[kind] = []
        for d in batches:
            y = d.graph.y
            loss = crit(model(d.graph, d.roost)[:, :1], (y / y.abs().max()).view(-1, 1))
            loss.backward()
            opt.step()
            sync.zero_grad()
            losses[kind].append(float(loss))
    for i, (a, b) in enumerate(zip(losses["fused"], losses["foreach"])):
        assert abs(a - b) <= 1e-5 + 1e-4 * abs(b), f"step {i}: fused {a} vs foreach {b}"
    assert abs(losses["fused"][1] - losses["fused"][0]) > 0


def test_flat_adamw_follows_torch_adamw():
    """optim.FlatAdamW (one cgat_adamw_flat launch over the flat parameter buffer, gradients packed per bucket from
    autograd hooks) follows torch.optim.AdamW step for step on the real model; state_dict names/shapes unchanged."""
    from cgat_b200 import optim
    batches = [synthetic.make_batch(40, 12, seed=s).to(DEV) for s in (1, 2, 3, 1)]
    losses = {}
    keys = None
    for kind in ("flat", "torch"):
        mkw = CASES["default_k12"][0]
        model = weights.load_seeded(cgat_b200.CGAtNet(200, **mkw), 0).to(DEV)
        keys = keys or {k: tuple(v.shape) for k, v in model.state_dict().items()}
        if kind == "flat":
            opt = optim.FlatAdamW(model, lr=1e-3, weight_decay=1e-6)
            assert {k: tuple(v.shape) for k, v in model.state_dict().items()} == keys
        else:
            opt = torch.optim.AdamW(model.parameters(), lr=1e-3, weight_decay=1e-6, foreach=True)
        losses[kind] = []
        for d in batches:
            y = d.graph.y
            loss = optim.l1_loss(model(d.graph, d.roost)[:, :1], (y / y.abs().max()).view(-1, 1))
            loss.backward()
            if kind == "flat":
                opt.sync.finish()
            opt.step()
            opt.zero_grad(set_to_none=True)
            losses[kind].append(float(loss.detach()))
    for i, (a, b) in enumerate(zip(losses["flat"], losses["torch"])):
        assert abs(a - b) <= 1e-5 + 1e-4 * abs(b), f"step {i}: flat {a} vs torch {b}"
    assert abs(losses["flat"][1] - losses["flat"][0]) > 0


def test_graphed_flat_adamw_matches_eager():
    """GraphedTrainStep with FlatAdamW + the own L1 kernel (what bench.py runs): graph replays across two buckets
    follow the eager loop of the same optimizer."""
    from cgat_b200 import batching, graphed, optim
    batches = [batching.pad_batch(synthetic.make_batch(40, 12, seed=s)) for s in (1, 2, 1, 2)]
    batches += [batching.pad_batch(synthetic.make_batch(90, 12, seed=3)), batching.pad_batch(synthetic.make_batch(40, 12, seed=2))]
    tg = [(b.graph.y / b.graph.y.abs().max()).view(-1, 1).to(DEV) for b in batches]
    mkw = CASES["default_k12"][0]
    model_e = weights.load_seeded(cgat_b200.CGAtNet(200, **mkw), 0).to(DEV)
    opt_e = optim.FlatAdamW(model_e, lr=1e-3, weight_decay=1e-6)
    eager = []
    for b, t in zip(batches, tg):
        d = b.to(DEV)
        n_real = d.graph.num_graphs - 1
        loss = optim.l1_loss(model_e(d.graph, d.roost)[:n_real, :1], t[:n_real])
        loss.backward()
        opt_e.sync.finish()
        opt_e.step()
        opt_e.zero_grad()
        eager.append(float(loss.detach()))
    model_g = weights.load_seeded(cgat_b200.CGAtNet(200, **mkw), 0).to(DEV)
    opt_g = optim.FlatAdamW(model_g, lr=1e-3, weight_decay=1e-6)
    runner = graphed.GraphedTrainStep(model_g, opt_g, optim.l1_loss)
    graph = [float(runner.step(b.pin_memory(), t)) for b, t in zip(batches, tg)]
    assert runner.captures >= 2 and runner.replayed_launches > 0
    for i, (a, b) in enumerate(zip(graph, eager)):
        assert abs(a - b) <= 1e-6 + 1e-5 * abs(b), f"step {i}: graphed loss {a} vs eager {b}"
    assert_close(opt_g.flat_p, opt_e.flat_p, "flat parameters after 6 steps", atol=1e-6, rtol=1e-5)


def test_graphed_step_matches_eager():
    """graphed.GraphedTrainStep (one CUDA graph per shape bucket: forward, loss, backward, AdamW) follows the eager
    loop step for step: same losses and same weights after steps that cross two buckets and revisit the first."""
    from cgat_b200 import batching, graphed
    from cgat_b200 import distributed as cdist
    batches = [batching.pad_batch(synthetic.make_batch(40, 12, seed=s)) for s in (1, 2, 1, 2, 1)]
    batches += [batching.pad_batch(synthetic.make_batch(90, 12, seed=3)),
                batching.pad_batch(synthetic.make_batch(40, 12, seed=2))]      # back to the first bucket's graph
    assert len({batching.signature(b) for b in batches}) >= 2
    tg = []
    for b in batches:
        y = b.graph.y
        tg.append((y / y.abs().max()).view(-1, 1).to(DEV))
    crit = torch.nn.L1Loss()

    model_e, opt_e = _train_setup()
    sync_e = cdist.GradSync(model_e, 1)
    eager_losses = []
    for b, t in zip(batches, tg):
        d = b.to(DEV)
        n_real = d.graph.num_graphs - 1
        loss = crit(model_e(d.graph, d.roost)[:n_real, :1], t[:n_real])
        loss.backward()
        opt_e.step()
        sync_e.zero_grad()
        eager_losses.append(float(loss))

    model_g, opt_g = _train_setup()
    runner = graphed.GraphedTrainStep(model_g, opt_g, crit, cdist.GradSync(model_g, 1))
    before = _lib.launch_count()
    graph_losses = [float(runner.step(b.pin_memory(), t)) for b, t in zip(batches, tg)]
    assert runner.captures >= 2 and runner.replayed_launches > 0 and _lib.launch_count() > before
    for i, (a, b) in enumerate(zip(graph_losses, eager_losses)):
        assert abs(a - b) <= 1e-6 + 1e-5 * abs(b), f"step {i}: graphed loss {a} vs eager {b}"
    pe, pg = dict(model_e.named_parameters()), dict(model_g.named_parameters())
    for k in ("embedding.weight", "graphs.0.Node.MH_M.fc_in.weight", "graphs.4.Node.MH_A.fc_out.weight",
              "graphs.2.Node.Pooling_NN.Hyper.layers.1.hyper_linear.hypo_params.net.4.weight", "output_nn.fc_out.weight"):
        assert_close(pg[k].detach(), pe[k].detach(), f"weights after 7 steps: {k}", atol=1e-6, rtol=1e-5)


def test_graphed_forward_matches_eager():
    from cgat_b200 import batching, graphed
    model, _ = _gpu_model("default_k12")
    runner = graphed.GraphedForward(model)
    for s in (5, 6, 5, 7):
        sb = synthetic.make_batch(64, 12, seed=s)
        pb = batching.pad_batch(sb)
        out = runner(pb.pin_memory()).clone()
        with torch.no_grad():
            ref = model(sb.graph.to(DEV), tuple(t.to(DEV) for t in sb.roost))
        assert_close(out, ref, f"graphed forward seed {s}", atol=2e-5, rtol=1e-4)
    assert runner.captures >= 1 and runner.replayed_launches > 0
