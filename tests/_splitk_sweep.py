"""Run by tests/test_gpu_gemm.py::test_split_k_sweep_does_not_hang in a time-bounded subprocess (a hang must not
take the whole GPU suite with it).  Exercises (a) the raw C ABI with split counts that leave EMPTY parts — the
round-1 deadlock — and (b) hyper_trunks forward + backward at the atom counts that hung."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from cgat_b200 import _lib, ops

DEV = "cuda:0"


def raw_tn(k, m, n, n_split):
    g = torch.Generator().manual_seed(k + n_split)
    a = torch.randn(k, m, generator=g).to(DEV)
    b = torch.randn(k, n, generator=g).to(DEV)
    part = torch.full((n_split, m, n), float("nan"), device=DEV)
    _lib.call("cgat_gemm3x_tn", a.data_ptr(), a.stride(0), b.data_ptr(), b.stride(0), part.data_ptr(), n, m * n,
              m, n, k, n_split, _lib.stream())
    ref = a.double().t() @ b.double()
    err = (part.sum(0).double() - ref).abs().max().item() / (a.double().abs().t() @ b.double().abs()).max().item()
    assert err < 1e-5, ("tn", k, n_split, err)


def raw_nt_splitk(m, n, k, n_split):
    g = torch.Generator().manual_seed(k + n_split)
    a = torch.randn(m, k, generator=g).to(DEV)
    w = torch.randn(n, k, generator=g).to(DEV)
    part = torch.full((n_split, m, n), float("nan"), device=DEV)
    _lib.call("cgat_gemm3x_nt_splitk", a.data_ptr(), a.stride(0), w.data_ptr(), w.stride(0), part.data_ptr(), n, m * n,
              m, n, k, n_split, _lib.stream())
    ref = a.double() @ w.double().t()
    err = (part.sum(0).double() - ref).abs().max().item() / (a.double().abs() @ w.double().abs().t()).max().item()
    assert err < 1e-5, ("nt_splitk", k, n_split, err)


def trunks(n):
    f, n_j = 128, 4
    g = torch.Generator().manual_seed(n)
    h = (torch.randn(n, f, generator=g) * 0.5).to(DEV).requires_grad_(True)
    layers = [[((torch.randn(f, f, generator=g) * (2.0 / f) ** 0.5).to(DEV).requires_grad_(True),
                (torch.randn(f, generator=g) * 0.1).to(DEV).requires_grad_(True)) for _ in range(4)]
              for _ in range(n_j)]
    tails = [((torch.randn(f * f + f, f, generator=g) * 0.0125).to(DEV), (torch.randn(f * f + f, generator=g) * 0.09).to(DEV),
              f * f) for _ in range(n_j)]
    zs, es = ops.hyper_trunks(h, layers, tails)
    (sum(z.sum() for z in zs) + sum(e.sum() for e in es)).backward()
    # weight gradient of the first tanh layer of trunk 0 against fp64: dW = D^T h with D = dL/d(pre-activation)
    hd = h.detach().double()
    w0, b0 = layers[0][0][0].detach().double(), layers[0][0][1].detach().double()
    assert torch.isfinite(layers[0][0][0].grad).all() and torch.isfinite(h.grad).all()
    return float(layers[0][0][0].grad.abs().sum())


def main():
    # (a) empty parts on purpose: 5376 rows in 18 parts of 320 -> the 18th starts at 5440 (nk was -1 in round 1)
    for k, ns in ((5376, 18), (5888, 18), (6400, 18), (4609, 18), (64, 7), (33, 40)):
        raw_tn(k, 128, 128, ns)
    for k, ns in ((5120, 10), (640, 9), (96, 8), (5376, 18)):
        raw_nt_splitk(200, 128, k, ns)
    # (b) the model-level sizes (raw atom counts of eager training and the padded buckets)
    for n in (4609, 5185, 5376, 5377, 5888, 6400):
        trunks(n)
    torch.cuda.synchronize()
    print("SPLITK_SWEEP_OK")


if __name__ == "__main__":
    main()
    sys.exit(0)
