"""Scratch: per-tensor gradient error of the CUDA path vs the fp64 oracle (default_k12 case)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cgat_b200
from cgat_b200 import synthetic, weights
from oracle import cgat_oracle as O
from tests._cases import CASES, oracle_cfg, training_scalar
name = sys.argv[1] if len(sys.argv) > 1 else "default_k12"
mkw, bkw, wseed = CASES[name]
model = weights.load_seeded(cgat_b200.CGAtNet(200, **mkw), wseed).cuda()
sb = synthetic.make_batch(**bkw); d = sb.to("cuda")
out = model(d.graph, d.roost); training_scalar(out, d.graph.y).backward()
shapes = {k: v.shape for k, v in model.state_dict().items()}
sd = weights.seeded_state_dict(shapes, wseed, torch.float64)
for v in sd.values(): v.requires_grad_(True)
sb64 = synthetic.make_batch(dtype=torch.float64, **bkw)
o = O.cgat_forward(sd, oracle_cfg(mkw), sb64.graph, sb64.roost); training_scalar(o, sb64.graph.y).backward()
print("out err", (out.detach().cpu().double() - o.detach()).abs().max().item())
rows = []
for k, p in model.named_parameters():
    if p.grad is None: continue
    g = p.grad.detach().cpu().double(); r = sd[k].grad
    e = (g - r).abs(); v = e / (1e-4 + 1e-3 * r.abs())
    rows.append((v.max().item(), int((v > 1).sum()), v.numel(), e.max().item(), r.abs().max().item(), (e.pow(2).sum().sqrt() / r.pow(2).sum().sqrt()).item(), k))
rows.sort(reverse=True)
for r in rows[:14]: print("viol %.2f bad %d/%d maxerr %.2e refmax %.2e relL2 %.2e %s" % r)
print("total bad", sum(r[1] for r in rows), "of", sum(r[2] for r in rows), "median relL2", sorted(r[5] for r in rows)[len(rows)//2])
