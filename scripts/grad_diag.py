"""Gradient-noise diagnostic: every parameter gradient of a golden case on the GPU against the fp64 oracle.
Run once per numerics switch (read at import time):
    python scripts/grad_diag.py default_k12                      # fused kernels
    CGAT_B200_TRUNK=0 python scripts/grad_diag.py default_k12    # library GEMMs for the hypernetwork trunks
    CGAT_B200_FUSED=0 python scripts/grad_diag.py default_k12    # library GEMMs everywhere (PyTorch's own fp32)
Prints, per tensor with any element outside (1e-4 abs, 1e-3 rel): count, count beyond 3x, max error; and the
median relative L2 error over all tensors."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import cgat_b200
from cgat_b200 import synthetic, weights
from oracle import cgat_oracle as O
from tests._cases import CASES, golden_shapes, load_golden, oracle_cfg, training_scalar

name = sys.argv[1] if len(sys.argv) > 1 else "default_k12"
mkw, bkw, wseed = CASES[name]
gold = load_golden(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden"), name)
dev = "cuda:0"
model = weights.load_seeded(cgat_b200.CGAtNet(200, **mkw), wseed).to(dev)
sb = synthetic.make_batch(**bkw)
d = sb.to(dev)
out = model(d.graph, d.roost)
training_scalar(out, d.graph.y).backward()
sd = weights.seeded_state_dict(golden_shapes(gold), wseed, torch.float64)
for v in sd.values():
    v.requires_grad_(True)
sb64 = synthetic.make_batch(dtype=torch.float64, **bkw)
ref = O.cgat_forward(sd, oracle_cfg(mkw), sb64.graph, sb64.roost)
training_scalar(ref, sb64.graph.y).backward()
print(f"switches: FUSED={os.environ.get('CGAT_B200_FUSED', '1')} TRUNK={os.environ.get('CGAT_B200_TRUNK', '1')}  "
      f"out err {(out.detach().cpu().double() - ref.detach()).abs().max().item():.2e}")
rels, tot_bad, tot_far = [], 0, 0
for k, p in model.named_parameters():
    if p.grad is None or sd[k].grad is None:
        continue
    a, b = p.grad.detach().cpu().double(), sd[k].grad
    err = (a - b).abs()
    tol = 1e-4 + 1e-3 * b.abs()
    bad, far = int((err > tol).sum()), int((err > 3 * tol).sum())
    rel = (err.norm() / (b.norm() + 1e-30)).item()
    rels.append(rel)
    tot_bad += bad
    tot_far += far
    if bad:
        print(f"  {k:90s} bad {bad:6d}/{b.numel():8d} far {far:5d} max {err.max().item():.2e} ref {b.abs().max().item():.2e} relL2 {rel:.2e}")
rels.sort()
print(f"median relL2 {rels[len(rels) // 2]:.2e}  max relL2 {rels[-1]:.2e}  total bad {tot_bad} far {tot_far}")
