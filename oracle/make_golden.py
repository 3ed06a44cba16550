"""TEST INFRASTRUCTURE ONLY — generates tests/golden/*.npz by running the UNMODIFIED reference.

Runs only in the build container (needs /root/reference).  It imports the reference's
CGAT/{CGAT,roost_message,message_changed,Hypernetworksmp}.py with the stand-ins of
oracle/standins.py, builds the reference CGAtNet, loads the seeded weights of
cgat_b200/weights.py, runs forward + backward on the seeded synthetic batches of
cgat_b200/synthetic.py and stores the outputs and a digest of the gradients.  Inputs and weights
are NOT stored: they are regenerated from their seeds on the other side.

    python oracle/make_golden.py            # rewrites every fixture
    python oracle/make_golden.py edge_hypernet      # only the named ones
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import standins  # noqa: E402
from cgat_b200 import synthetic, weights  # noqa: E402

# name -> (model kwargs, batch kwargs, weight seed)
CASES = {
    "default_k12": (dict(elem_fea_len=128, n_graph=5, msg_heads=5, neighbor_number=12, mean_pooling=False,
                         rezero=True, update_edges=True, vector_attention=True, global_vector_attention=True,
                         n_graph_roost=3),
                    dict(n_crystals=12, max_nbr=12, seed=0), 0),
    "default_k24": (dict(elem_fea_len=128, n_graph=5, msg_heads=5, neighbor_number=24, mean_pooling=False,
                         rezero=True, update_edges=True, vector_attention=True, global_vector_attention=True,
                         n_graph_roost=3),
                    dict(n_crystals=6, max_nbr=24, seed=1), 1),
    "scalar_attn_meanpool": (dict(elem_fea_len=64, n_graph=3, msg_heads=3, neighbor_number=12, mean_pooling=True,
                                  rezero=False, update_edges=True, vector_attention=False,
                                  global_vector_attention=False, n_graph_roost=2),
                             dict(n_crystals=16, max_nbr=12, seed=2), 2),
    "mixed_flags": (dict(elem_fea_len=32, n_graph=2, msg_heads=4, neighbor_number=8, mean_pooling=False,
                         rezero=True, update_edges=True, vector_attention=True, global_vector_attention=False,
                         n_graph_roost=1),
                    dict(n_crystals=20, max_nbr=8, seed=3), 3),
    "large_cell_k24": (dict(elem_fea_len=32, n_graph=2, msg_heads=2, neighbor_number=24, mean_pooling=False,
                            rezero=True, update_edges=True, vector_attention=True, global_vector_attention=True,
                            n_graph_roost=3),
                       dict(n_crystals=2, max_nbr=24, seed=4, atoms_lo=200, atoms_hi=256), 4),
    "edge_hypernet": (dict(elem_fea_len=64, n_graph=3, msg_heads=2, neighbor_number=8, mean_pooling=False,
                           rezero=True, update_edges=True, vector_attention=True, global_vector_attention=True,
                           n_graph_roost=1, no_hyper=False),
                      dict(n_crystals=8, max_nbr=8, seed=5), 5),
    "wide_f256_h8": (dict(elem_fea_len=256, n_graph=2, msg_heads=8, neighbor_number=12, mean_pooling=False,
                          rezero=True, update_edges=True, vector_attention=True, global_vector_attention=True,
                          n_graph_roost=3),
                     dict(n_crystals=10, max_nbr=12, seed=6), 6),
}

# gradient tensors stored in full (small); every other gradient is stored as (sum, abs-sum, l2)
FULL_GRADS = ("nbr_embedding.weight", "graphs.1.Node.Pooling_NN.damping", "graphs.0.Node.MH_A.fc_out.bias",
              "graphs.0.Node.MH_M.fc_in.bias", "graphs.0.Edge.Pooling_NN.fc_out.bias", "cry_pool.MH_A.fc_out.bias",
              "roost.graphs.0.pooling.0.pow", "roost.cry_pool.0.pow", "output_nn.fc_out.weight",
              "output_nn.rezeros.0.alpha", "roost.embedding.bias",
              "graphs.0.Node.Pooling_NN.Hyper.layers.3.hypo_params.net.4.bias")


def run_case(name, ref):
    mkw, bkw, wseed = CASES[name]
    torch.manual_seed(0)
    model = ref["CGAT"].CGAtNet(200, **mkw)
    weights.load_seeded(model, wseed)
    sb = synthetic.make_batch(**bkw)
    batch = standins.Batch(x=sb.graph.x, edge_index=sb.graph.edge_index, edge_attr=sb.graph.edge_attr, y=sb.graph.y)
    batch.batch = sb.graph.batch
    out = model(batch, (t for t in sb.roost))
    emb = model(batch, (t for t in sb.roost), return_graph_embedding=True)
    pen = model(batch, (t for t in sb.roost), last_layer=False)
    # training-shaped scalar: L1 on the first output column against y (reference lightning_module.py:237-240)
    target = sb.graph.y.view(-1, 1) / sb.graph.y.abs().max()
    loss = (out[:, :1] - target).abs().mean() + 0.1 * out[:, 1].mean()
    model.zero_grad()
    loss.backward()
    rec = {"out": out.detach().numpy(), "embedding": emb.detach().numpy(), "penultimate": pen.detach().numpy(),
           "loss": np.array(loss.item(), dtype=np.float64)}
    names, digests, none_grads = [], [], []
    for k, p in model.named_parameters():
        if p.grad is None:
            none_grads.append(k)
            continue
        g = p.grad.double()
        names.append(k)
        digests.append([g.sum().item(), g.abs().sum().item(), g.pow(2).sum().sqrt().item()])
        if k in FULL_GRADS:
            rec["grad::" + k] = p.grad.numpy().copy()
    rec["grad_names"] = np.array(names)
    rec["grad_digest"] = np.array(digests, dtype=np.float64)
    rec["none_grads"] = np.array(none_grads)
    sd = model.state_dict()
    rec["state_names"] = np.array(list(sd.keys()))
    rec["state_shapes"] = np.array([",".join(map(str, v.shape)) for v in sd.values()])
    print(f"{name}: C={sb.num_crystals} N={sb.graph.x.shape[0]} E={sb.graph.edge_index.shape[1]} "
          f"out[0]={out[0].tolist()} loss={loss.item():.6f} none_grads={len(none_grads)}")
    return rec


def main():
    ref = standins.import_reference()
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    only = set(sys.argv[1:])          # python oracle/make_golden.py [case ...]: default every fixture
    for name in CASES:
        if only and name not in only:
            continue
        rec = run_case(name, ref)
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", f"ref_{name}.npz"), **rec)


if __name__ == "__main__":
    main()
