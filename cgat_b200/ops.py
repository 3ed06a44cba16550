"""autograd.Functions over the C ABI (include/cgat_b200.h).  Host glue only: every Function
allocates outputs, calls one or more kernels on the current stream, and wires the hand-written
backward kernels into autograd.  No CPU path."""
from __future__ import annotations

import os

import torch

from . import _lib
from .graph import SegmentPlan


def _f32c(t):
    if t.dtype != torch.float32:
        raise TypeError("cgat_b200 kernels are fp32")
    return t.contiguous()


class _SegSoftmax(torch.autograd.Function):
    """out[s,h,:] = sum_{t in seg s} alpha[t,h,:] * value[t,h,:],
    alpha = u * exp(gate - segmax) / (segsum(u * exp(gate - segmax)) + eps).

    Replaces torch_geometric.utils.softmax + scatter_add (reference CGAT/CGAT.py:59-61, :323-326)
    and scatter_max / scatter_add in WeightedAttention (reference CGAT/roost_message.py:307-315)."""

    @staticmethod
    def forward(ctx, gate, value, u, ptr, seg_of_row, n_seg, eps):
        gate, value = _f32c(gate), _f32c(value)
        u = None if u is None else _f32c(u)
        n_rows, heads, f = value.shape
        fa = gate.shape[2]
        out = torch.empty((n_seg, heads, f), dtype=torch.float32, device=value.device)
        smax = torch.empty((n_seg, heads, fa), dtype=torch.float32, device=value.device)
        sden = torch.empty_like(smax)
        _lib.call("cgat_seg_softmax_fwd", _lib.ptr(gate), _lib.ptr(value), _lib.ptr(u), _lib.ptr(ptr), n_seg,
                  heads, f, fa, eps, _lib.ptr(out), _lib.ptr(smax), _lib.ptr(sden), _lib.stream(),
                  work=dict(key="seg_softmax_fwd", bound="hbm",
                            bytes=4 * (n_rows * heads * (f + fa) + n_seg * heads * (f + 2 * fa)),
                            note="reads gate+value once, writes out+max+den"))
        ctx.save_for_backward(gate, value, u, seg_of_row, out, smax, sden)
        ctx.eps = eps
        return out

    @staticmethod
    def backward(ctx, d_out):
        gate, value, u, seg_of_row, out, smax, sden = ctx.saved_tensors
        d_out = _f32c(d_out)
        n_rows, heads, f = value.shape
        fa = gate.shape[2]
        d_gate = torch.empty_like(gate)
        d_value = torch.empty_like(value)
        _lib.call("cgat_seg_softmax_bwd", _lib.ptr(gate), _lib.ptr(value), _lib.ptr(u), _lib.ptr(seg_of_row),
                  _lib.ptr(out), _lib.ptr(smax), _lib.ptr(sden), _lib.ptr(d_out), n_rows, heads, f, fa,
                  ctx.eps, _lib.ptr(d_gate), _lib.ptr(d_value), _lib.stream(),
                  work=dict(key="seg_softmax_bwd", bound="hbm",
                            bytes=4 * (2 * n_rows * heads * (f + fa) + out.shape[0] * heads * (2 * f + 2 * fa)),
                            note="reads gate+value, writes d_gate+d_value, per-segment stats once"))
        d_u = None
        if u is not None and ctx.needs_input_grad[2]:
            # alpha ∝ exp(gate + log u)  =>  dL/du = (dL/dgate summed over heads/channels) / u
            d_u = d_gate.sum(dim=(1, 2)) / u
        return d_gate, d_value, d_u, None, None, None, None


def seg_softmax(gate, value, plan: SegmentPlan | None = None, *, ptr=None, seg_of_row=None, n_seg=None,
                u=None, eps=1e-16):
    """gate (n,H,Fa), value (n,H,F) -> (n_seg,H,F).  Either a SegmentPlan or (ptr, seg_of_row, n_seg)."""
    if plan is not None:
        ptr, seg_of_row, n_seg = plan.ptr, plan.index, plan.n_seg
    return _SegSoftmax.apply(gate, value, u, ptr, seg_of_row, n_seg, float(eps))


# ----------------------------------------------------------------------------------------------
# Dense pieces.  Stage A: library GEMMs (torch / cuBLAS fp32, TF32 disabled by torch default) with
# the hand-written segment kernels in between.  The fused sm_100a kernels replace these one by one.
# ----------------------------------------------------------------------------------------------
LEAKY_SLOPE = 0.01
# numerics A/B switch for development (1 = fused tensor-core kernels; 0 = library GEMMs + segment kernels)
_FUSED = os.environ.get("CGAT_B200_FUSED", "1") != "0"
_TN_WGRAD = os.environ.get("CGAT_B200_TN", "1") != "0"   # weight gradients on cgat_gemm3x_tn vs library mm
_TRUNK = os.environ.get("CGAT_B200_TRUNK", "1") != "0"   # hypernetwork trunks: fused chain kernel vs library GEMMs
# activation x weight tensor-core kernels on kind::f16 with scaled fp16 hi/lo operands (1) or kind::tf32 hi/lo (0)
_F16X3 = os.environ.get("CGAT_B200_F16X3", "1") != "0"
_F16X3_EDGE = os.environ.get("CGAT_B200_F16X3_EDGE", "1") != "0"   # the same for the edge-attention forward / bwd_prep
# gradient-operand kernels (hyper weight gradient) on kind::f16 with a per-launch power-of-two scale (1) or kind::tf32 (0)
_F16X3_GRAD = os.environ.get("CGAT_B200_F16X3_GRAD", "1") != "0"
_EDGE_W2_PRESCALE = 64.0   # cgat_edge_attn_*_f16 expect W2 packed as f16(w * 2^6) hi/lo with an unscaled lo
# EXPERIMENTAL, off by default (not yet measured on the GPU): the small MLPs around the fused kernels (Roost, crystal
# pool, output network, edge table) on the 3-pass tensor-core GEMMs instead of the library's SIMT sgemm kernels
_LINEAR3X = os.environ.get("CGAT_B200_LINEAR3X", "0") == "1"


def multi_head_mlp(fea, w_in, b_in, w_out, b_out, heads):
    """fea (n,In) -> (n,H,Out): per head h  W2_h leaky_relu(W1_h fea + b1_h) + b2_h
    (reference MultiHeadNetwork.forward, CGAT/CGAT.py:103-109)."""
    n = fea.shape[0]
    out_dim, hd = w_out.shape[1], w_out.shape[2]
    if _LINEAR3X and fea.is_cuda and n > 0 and fea.shape[1] % 4 == 0 and hd % 4 == 0 and out_dim % 4 == 0:
        # own 3-pass tensor-core GEMMs: one wide first layer (bias + LeakyReLU in its epilogue), then one GEMM per
        # head on that head's column block of the hidden activations (row-strided operand, no copy)
        hid = linear_act(fea, w_in, b_in, 1)                                                  # (n, H*Hd)
        outs = [_Linear3x.apply(hid[:, h * hd:(h + 1) * hd], w_out[h], b_out[h * out_dim:(h + 1) * out_dim], 0)
                for h in range(heads)]
        return torch.stack(outs, dim=1)                                                       # (n, H, Out)
    hid = torch.nn.functional.leaky_relu(torch.addmm(b_in, fea, w_in.t()), LEAKY_SLOPE)      # (n, H*Hd)
    hid = hid.view(n, heads, hd).transpose(0, 1)                                              # (H, n, Hd)
    out = torch.baddbmm(b_out.view(heads, 1, out_dim), hid, w_out.transpose(1, 2))            # (H, n, Out)
    return out.transpose(0, 1)                                                                # (n, H, Out)


def edge_attention_unfused(x, edge_table, plan, w1a, b1a, w2a, b2a, w1m, b1m, w2m, b2m, heads, edge_ids=None):
    """Node-attention aggregation, reference CGAT/CGAT.py:319-329 (Appendix A of SURVEY.md) — library GEMMs +
    the segmented-softmax kernel.  Used for the shapes the fused kernel is not instantiated for and as
    the recompute path of its backward.

    m_t = [x[dst]; e(rank_t); x[src]]; the first MLP layer is linear in m_t, so it is evaluated per
    atom and per rank and only summed per edge:  W1 m_t = (x W1_i^T)[dst] + (e W1_e^T + b1)[rank] + (x W1_j^T)[src].
    Edges are visited in destination-sorted order (plan)."""
    n, f = x.shape
    fe = edge_table.shape[1]
    hhd = w1a.shape[0]
    hd = hhd // heads
    w1 = torch.cat([w1a, w1m], dim=0)                              # (2*H*Hd, 2F+Fe)
    p_dst = x @ w1[:, :f].t()                                      # (N, 2*H*Hd)
    p_src = x @ w1[:, f + fe:].t()
    t_rank = torch.addmm(torch.cat([b1a, b1m]), edge_table, w1[:, f:f + fe].t())   # (K+1, 2*H*Hd) / (E, 2*H*Hd)
    pre = (p_dst.index_select(0, plan.dst) + p_src.index_select(0, plan.src)
           + t_rank.index_select(0, plan.rank if edge_ids is None else edge_ids))
    hid = torch.nn.functional.leaky_relu(pre, LEAKY_SLOPE)
    e = hid.shape[0]
    hid_a = hid[:, :hhd].view(e, heads, hd).transpose(0, 1)        # (H, E, Hd)
    hid_m = hid[:, hhd:].view(e, heads, hd).transpose(0, 1)
    gate = torch.baddbmm(b2a.view(heads, 1, -1), hid_a, w2a.transpose(1, 2)).transpose(0, 1).contiguous()
    msg = torch.baddbmm(b2m.view(heads, 1, -1), hid_m, w2m.transpose(1, 2)).transpose(0, 1).contiguous()
    agg = seg_softmax(gate, msg, ptr=plan.rowptr, seg_of_row=plan.dst, n_seg=n, eps=1e-16)   # (N,H,F)
    return agg.mean(dim=1)


# ----------------------------------------------------------------------------------------------
# Tensor-core pieces (tcgen05, three error-compensated passes per product — kind::tf32 or kind::f16; fp32 in / out)
# ----------------------------------------------------------------------------------------------
def gemm3x(a, w, bias=None, act=0, out=None):
    """act(a @ w.T + bias) on the tensor cores (cgat_gemm3x_nt). a (M,K), w (N,K): rows contiguous."""
    M, K = a.shape
    N = w.shape[0]
    if a.stride(1) != 1 or w.stride(1) != 1:
        raise ValueError("gemm3x needs row-contiguous operands")
    if out is None:
        out = torch.empty((M, N), dtype=torch.float32, device=a.device)
    if not a.is_cuda:
        raise _lib.CgatLibraryError("cgat_b200 kernels need CUDA tensors (there is no CPU path)")
    tiles = ((M + 127) // 128) * ((N + (63 if N <= 64 else 127)) // (64 if N <= 64 else 128))
    if tiles <= 74 and K >= 256 and N % 4 == 0 and out.is_contiguous():
        # few output tiles and a long contraction (the Roost / crystal-pool / output MLPs: M = crystals or elements):
        # one 128 x 128 tile per CTA would leave most SMs idle and walk K serially, so split K over CTAs and apply
        # bias + activation while the parts are summed
        n_split = plan_split(K, max(2, min(K // 128, 148 // tiles)))
        if n_split > 1:
            part = torch.empty((n_split, M, N), dtype=torch.float32, device=a.device)
            _lib.call("cgat_gemm3x_nt_splitk", a.data_ptr(), a.stride(0), w.data_ptr(), w.stride(0), part.data_ptr(), N,
                      M * N, M, N, K, n_split, _lib.stream(),
                      work=dict(key="gemm3x_nt", bound="tensor", flops=2.0 * M * N * K))
            _lib.call("cgat_sum_parts_bias_act", _lib.ptr(part), n_split, M * N, _lib.ptr(bias), _lib.ptr(out), M, N, act,
                      _lib.stream(), work=dict(key="sum_parts", bound="hbm", bytes=4.0 * M * N * (n_split + 1)))
            return out
    _lib.call("cgat_gemm3x_nt", a.data_ptr(), a.stride(0), w.data_ptr(), w.stride(0), _lib.ptr(bias),
              out.data_ptr(), out.stride(0), M, N, K, act, _lib.stream(),
              work=dict(key="gemm3x_nt", bound="tensor", flops=2.0 * M * N * K))
    return out


class _Linear3x(torch.autograd.Function):
    """act(x @ w.T + b) with forward on cgat_gemm3x_nt (bias + activation in its epilogue) and backward on
    cgat_gemm3x_nt (dL/dx) / cgat_gemm3x_tn (dL/dw).  act: 0 none, 1 LeakyReLU(0.01), 3 ReLU — both have
    sign(out) == sign(pre-activation), so only the output is saved."""

    @staticmethod
    def forward(ctx, x, weight, bias, act):
        if x.dtype != torch.float32:
            raise TypeError("cgat_b200 kernels are fp32")
        if x.stride(1) != 1 or x.stride(0) % 4 or x.data_ptr() % 16:
            x = x.contiguous()             # row-strided column blocks of a wider matrix are taken as they are
        w = weight.contiguous()
        out = gemm3x(x, w, None if bias is None else bias.contiguous(), act)
        ctx.act, ctx.has_bias = act, bias is not None
        ctx.save_for_backward(x, w, out if act else None)
        return out

    @staticmethod
    def backward(ctx, g):
        x, w, out = ctx.saved_tensors
        g = _f32c(g)
        if ctx.act == 1:
            g = torch.where(out > 0, g, g * LEAKY_SLOPE)
        elif ctx.act == 3:
            g = g * (out > 0)
        g_x = gemm3x(g, w.t().contiguous()) if ctx.needs_input_grad[0] else None     # (M,N) @ (N,K)
        g_w = gemm3x_tn(g, x) if ctx.needs_input_grad[1] else None                    # g^T x: (N,K)
        g_b = g.sum(dim=0) if ctx.has_bias and ctx.needs_input_grad[2] else None
        return g_x, g_w, g_b, None


def linear_act(x, weight, bias=None, act=0):
    """act(x @ weight.T + bias) for a 2-D x; act: 0 none, 1 LeakyReLU(0.01), 3 ReLU.  The library path (cuBLAS fp32
    via torch) unless CGAT_B200_LINEAR3X=1 selects the 3-pass tensor-core GEMMs."""
    # both contractions of the layer (over K forward, over N_out in dL/dx) need a multiple of 4: the 127-wide Roost
    # embedding and the 1- / 2-wide output heads stay on the library
    if (_LINEAR3X and x.is_cuda and x.dim() == 2 and x.shape[1] % 4 == 0 and weight.shape[0] % 4 == 0
            and x.shape[0] > 0):
        return _Linear3x.apply(x, weight, bias, act)
    y = torch.nn.functional.linear(x, weight, bias)
    if act == 1:
        return torch.nn.functional.leaky_relu(y, LEAKY_SLOPE)
    if act == 3:
        return torch.relu(y)
    return y


def gemm3x_res(a, w_packed, n_out, bias=None, act=0):
    """act(a @ w.T + bias) for K <= 128 against a wide pre-packed weight (cgat_gemm3x_nt_res):
    a (M,K) row-contiguous, w_packed = packed_kmajor of w (n_out, K)."""
    M, K = a.shape
    if a.stride(1) != 1 or not a.is_cuda:
        raise ValueError("gemm3x_res needs a row-contiguous CUDA operand")
    out = torch.empty((M, n_out), dtype=torch.float32, device=a.device)
    _lib.call("cgat_gemm3x_nt_res", a.data_ptr(), a.stride(0), _lib.ptr(w_packed), _lib.ptr(bias), out.data_ptr(),
              n_out, M, n_out, K, act, _lib.stream(),
              work=dict(key="gemm3x_nt_res", bound="tensor", flops=2.0 * M * n_out * K,
                        bytes=4.0 * (M * K + M * n_out)))
    return out


def sum_parts(parts, out=None, accumulate=False):
    """out = [out +] parts.sum(dim=0) with this library's fixed-order reduction kernel (cgat_sum_parts): parts is a
    contiguous (n_parts, ...) stack of split-K / split-atom partial results.  Replaces torch.sum at every such
    place (176 library `reduce_kernel` launches per cfg2 step in round 1, profiles/r01n_*_launches_summary.txt)."""
    n_parts = parts.shape[0]
    if out is None:
        if n_parts == 1:
            return parts[0]
        out = torch.empty(parts.shape[1:], dtype=torch.float32, device=parts.device)
    n = out.numel()
    if n == 0 or n_parts == 0:
        return out if accumulate else out.zero_()
    if not parts.is_contiguous() or not out.is_contiguous():
        raise ValueError("sum_parts needs contiguous buffers")
    _lib.call("cgat_sum_parts", _lib.ptr(parts), n_parts, n, _lib.ptr(out), n, int(accumulate), _lib.stream(),
              work=dict(key="sum_parts", bound="hbm", bytes=4.0 * n * (n_parts + 1 + int(accumulate))))
    return out


def plan_split(k, want, granule=32):
    """Number of split-K parts actually used for a contraction of length k when `want` parts are asked for.
    The kernels round the per-part length up to `granule` (one 32-float K chunk):
        k_per = ceil(ceil(k / n) / granule) * granule,
    so for some k the last parts would start at or beyond k and be EMPTY (round 1: such a part hung
    cgat_gemm3x_tn — VERDICT r01 weak #1; the kernels now write zeros for it).  This returns
    n = ceil(k / k_per(want)) <= want, for which no part is empty: (n - 1) * k_per(n) < k."""
    want = max(1, int(want))
    if k <= 0:
        return 1
    k_per = -(-(-(-k // want)) // granule) * granule
    return max(1, -(-k // k_per))


N_SMS = 148


def best_split(tiles, k, max_split, granule=32, fixed=4.0, per_part=0.25):
    """Split-K count for `tiles` output tiles and a contraction of length k on one-CTA-per-SM kernels: the CTAs run in
    ceil(tiles * n / 148) waves of ceil(k / n / granule) chunks each, so the count that asks for the most CTAs is often
    NOT the fastest (15 splits of 20 tiles = 300 CTAs = a third wave for 4 CTAs; 5 splits of 44 tiles = 1.5 waves).
    Cost in chunk times: waves * (chunks + fixed per-CTA overhead) + per_part * n for the partial sum; ties -> fewer parts."""
    best, best_cost = 1, None
    for n in range(1, max(1, int(max_split)) + 1):
        n_eff = plan_split(k, n, granule)
        if n_eff != n:
            continue
        k_per = -(-(-(-k // n)) // granule) * granule
        waves = -(-tiles * n // N_SMS)
        cost = waves * (k_per // granule + fixed) + per_part * n
        if best_cost is None or cost < best_cost - 1e-9:
            best, best_cost = n, cost
    return best


def gemm3x_splitk(a, w, n_split=None):
    """a @ w.T for a long contraction with few output tiles (cgat_gemm3x_nt_splitk): a (M,K), w (N,K)."""
    M, K = a.shape
    N = w.shape[0]
    if a.stride(1) != 1 or w.stride(1) != 1 or not a.is_cuda:
        raise ValueError("gemm3x_splitk needs row-contiguous CUDA operands")
    tiles = ((M + 127) // 128) * ((N + 127) // 128)
    if n_split is None:
        n_split = best_split(tiles, K, min((K + 511) // 512, (2 * 148 + tiles - 1) // tiles))
    n_split = plan_split(K, n_split)
    part = torch.empty((n_split, M, N), dtype=torch.float32, device=a.device)
    _lib.call("cgat_gemm3x_nt_splitk", a.data_ptr(), a.stride(0), w.data_ptr(), w.stride(0), part.data_ptr(), N, M * N,
              M, N, K, n_split, _lib.stream(), work=dict(key="gemm3x_nt", bound="tensor", flops=2.0 * M * N * K))
    return sum_parts(part)


def gemm3x_tn(a, b, n_split=None):
    """a.T @ b on the tensor cores (cgat_gemm3x_tn): a (K,M), b (K,N) row-contiguous -> (M,N).  The
    weight-gradient shape: contraction over the rows (atoms / edges) of both operands."""
    K, M = a.shape
    N = b.shape[1]
    if a.stride(1) != 1 or b.stride(1) != 1 or not a.is_cuda:
        raise ValueError("gemm3x_tn needs row-contiguous CUDA operands")
    tiles = ((M + 127) // 128) * ((N + 127) // 128)
    if n_split is None:
        n_split = best_split(tiles, K, min((K + 255) // 256, (2 * 148 + tiles - 1) // tiles))
    n_split = plan_split(K, n_split)
    part = torch.empty((n_split, M, N), dtype=torch.float32, device=a.device)
    _lib.call("cgat_gemm3x_tn", a.data_ptr(), a.stride(0), b.data_ptr(), b.stride(0), part.data_ptr(), N, M * N,
              M, N, K, n_split, _lib.stream(), work=dict(key="gemm3x_tn", bound="tensor", flops=2.0 * M * N * K))
    return sum_parts(part)


# Packed operand caches are keyed on (the weight's autograd version, this epoch).  The version counter alone is
# not enough: fused optimizers (torch.optim.AdamW(fused=True), i.e. torch._fused_adamw_) update parameters WITHOUT
# bumping it, and a CUDA-graph replay updates them without Python running at all.  The epoch is bumped
#   * after every torch.optim.Optimizer.step() in the process (global post-step hook, registered below),
#   * by cgat_b200/graphed.py before every capture of a training graph (the capture then contains its own pack
#     launches) and after every replay (so that eager code repacks too).
_pack_epoch = 0


def invalidate_packed():
    global _pack_epoch
    _pack_epoch += 1


def _after_optimizer_step(optimizer, args, kwargs):
    invalidate_packed()


from torch.optim.optimizer import register_optimizer_step_post_hook as _register_post_step  # noqa: E402

_register_post_step(_after_optimizer_step)


def packed_kmajor(w, rows=None, transpose=False, f16=False, pre_scale=None):
    """Pre-split / pre-swizzled copy of w[:rows] (2-D view of a contiguous weight) for the fused
    kernels: tf32 hi/lo images; with f16=True fp16 hi/lo images, either (hi, lo * 2^11) for the kernels with a
    separate correction accumulator or, with pre_scale, (f16(w * pre_scale), unscaled lo).  Cached ON the tensor and
    keyed by its autograd version counter, so an in-place optimizer update (which bumps `_version`) triggers a
    repack and a freed tensor cannot leave a stale entry behind."""
    w2 = w.detach().view(w.shape[0], -1)
    rows = w2.shape[0] if rows is None else rows
    cache = w.__dict__.setdefault("_cgat_packed", {})
    key = (rows, int(transpose), bool(f16), pre_scale)
    hit = cache.get(key)
    if hit is not None and hit[0] == (w._version, _pack_epoch) and hit[2] == w.data_ptr():
        return hit[1]
    lib = _lib.load()
    n_rows, k = (w2.shape[1], rows) if int(transpose) == 1 else (rows, w2.shape[1])
    size = lib.cgat_packed_floats_f16(n_rows, k) if f16 else lib.cgat_packed_floats(n_rows, k)
    # always a FRESH buffer: the previous image may be held by a pending backward (ctx.save_for_backward) whose
    # forward ran before the weights changed; rewriting it in place through a raw pointer would hand that backward
    # the new weights without autograd's version check noticing.  The old image dies with its graph.
    buf = torch.empty(int(size), dtype=torch.float32, device=w.device)
    work = dict(key="pack_kmajor", bound="hbm", bytes=(8.0 if f16 else 12.0) * n_rows * k)
    if f16 and pre_scale is not None:
        _lib.call("cgat_pack_kmajor_f16s", _lib.ptr(w2), w2.stride(0), n_rows, k, int(transpose), float(pre_scale), 1.0,
                  _lib.ptr(buf), _lib.stream(), work=work)
    else:
        _lib.call("cgat_pack_kmajor_f16" if f16 else "cgat_pack_kmajor", _lib.ptr(w2), w2.stride(0), n_rows, k,
                  int(transpose), _lib.ptr(buf), _lib.stream(), work=work)
    cache[key] = ((w._version, _pack_epoch), buf, w.data_ptr())
    return buf


class _HyperLinear(torch.autograd.Function):
    """y_out[n] = reshape(W z[n] + b)[:F*F] y[n] + (W z[n] + b)[F*F:]   (reference HyperLinear.forward +
    BatchLinear.forward, CGAT/Hypernetworksmp.py:243-254, 205-209) with the (N, F*F+F) predicted-weight
    tensor never materialised in forward (cgat_hyper_rowdot_fwd).  `e` (optional) is the z-part of the
    bias tail, W[F*F:] z + b[F*F:], when the fused trunk kernel has already produced it."""

    @staticmethod
    def forward(ctx, z, weight, bias, y, e, w_packed, w_packed_bt, f16):
        z, y = _f32c(z), _f32c(y)
        ctx.f16 = f16
        n, f = y.shape
        ff = f * f
        bias = bias.contiguous()
        ctx.has_e = e is not None
        if e is None:
            # bias tail of the last Linear: e = z W[F*F:]^T + b[F*F:] (the fused trunk kernel supplies it otherwise)
            e = gemm3x(z, weight[ff:], bias[ff:])
        out = torch.empty_like(y)
        # the bias of the predicted weights, b[:F*F], is added inside the kernel's epilogue
        _lib.call("cgat_hyper_rowdot_fwd_f16" if f16 else "cgat_hyper_rowdot_fwd", _lib.ptr(z), _lib.ptr(y),
                  _lib.ptr(_f32c(e)), None, _lib.ptr(bias),
                  _lib.ptr(w_packed), _lib.ptr(out), n, f, _lib.stream(),
                  work=dict(key="hyper_rowdot_fwd", bound="tensor", flops=2.0 * n * f * ff,
                            note="f16x3: 3 kind::f16 passes per algorithmic flop" if f16 else
                            "3xTF32: 3 tensor passes per algorithmic flop"))
        ctx.save_for_backward(z, weight, bias, y, w_packed, w_packed_bt)
        return out

    @staticmethod
    def backward(ctx, g):
        z, weight, bias, y, w_packed, w_packed_bt = ctx.saved_tensors
        n, f = y.shape
        ff = f * f
        g = _f32c(g)
        parts = int(_lib.load().cgat_hyper_rowscale_parts_f16(n, f) if ctx.f16 else _lib.load().cgat_hyper_rowscale_parts(n, f))
        rowscale = "cgat_hyper_rowscale_f16" if ctx.f16 else "cgat_hyper_rowscale"
        work = dict(key="hyper_rowscale", bound="tensor", flops=2.0 * n * f * ff,
                    note="f16x3: 3 kind::f16 passes per algorithmic flop" if ctx.f16 else
                    "3xTF32: 3 tensor passes per algorithmic flop")
        # dL/dy[n,i] = sum_o g[n,o] (W z[n] + b)[o*F+i]: recomputed tile by tile on the tensor cores
        bias = bias.contiguous()
        buf = torch.empty((parts, n, f), dtype=torch.float32, device=y.device)
        f16_grad = ctx.f16 and _F16X3_GRAD
        if f16_grad:
            # the same launch also leaves max |g| in device memory: the range of cgat_hyper_wgrad_f16's gradient operand
            g_amax = torch.zeros(1, dtype=torch.float32, device=y.device)
            _lib.call("cgat_hyper_rowscale_f16_amax", _lib.ptr(z), _lib.ptr(g), _lib.ptr(bias), _lib.ptr(w_packed),
                      _lib.ptr(buf), _lib.ptr(g_amax), n, f, _lib.stream(), work=work)
        else:
            _lib.call(rowscale, _lib.ptr(z), _lib.ptr(g), _lib.ptr(bias), _lib.ptr(w_packed), _lib.ptr(buf),
                      n, f, _lib.stream(), work=work)
        g_y = sum_parts(buf)
        # dL/dz[n,k] = sum_o g[n,o] sum_i y[n,i] W[o*F+i,k] (+ the bias-tail rows unless they went through `e`)
        buf2 = torch.empty_like(buf)
        _lib.call(rowscale, _lib.ptr(y), _lib.ptr(g), None, _lib.ptr(w_packed_bt), _lib.ptr(buf2), n, f,
                  _lib.stream(), work=work)
        g_z = sum_parts(buf2)
        if not ctx.has_e:
            g_z = g_z + gemm3x(g, weight[ff:].t().contiguous())
        # weight gradient: dW[o*F+i,k] = sum_n g[n,o] y[n,i] z[n,k] — contraction over atoms, outer product on the fly
        lib = _lib.load()
        splits = int(lib.cgat_hyper_wgrad_splits(n)) if f == 128 else 1     # F = 256: 512 CTAs without an atom split
        wpart = torch.empty((splits, ff, f), dtype=torch.float32, device=y.device)
        # F = 128: the same launch also sums the bias-shaped gradients g^T [y | z] from the rows it stages
        tail = torch.empty((splits, f, 2 * f), dtype=torch.float32, device=y.device) if f16_grad and f == 128 else None
        if f16_grad:
            _lib.call("cgat_hyper_wgrad_f16", _lib.ptr(g), _lib.ptr(y), _lib.ptr(z), _lib.ptr(g_amax), _lib.ptr(wpart),
                      _lib.ptr(tail), n, f, _lib.stream(),
                      work=dict(key="hyper_wgrad", bound="tensor", flops=2.0 * n * f * ff,
                                note="f16x3: 3 kind::f16 passes per algorithmic flop, gradient operand scaled by 2^k"))
        else:
            _lib.call("cgat_hyper_wgrad", _lib.ptr(g), _lib.ptr(y), _lib.ptr(z), _lib.ptr(wpart), n, f, _lib.stream(),
                      work=dict(key="hyper_wgrad", bound="tensor", flops=2.0 * n * f * ff,
                                note="3xTF32: 3 tensor passes per algorithmic flop"))
        g_w = torch.empty_like(weight)
        sum_parts(wpart, out=g_w[:ff])
        if tail is not None:
            gyz = sum_parts(tail)                          # (F, 2F) = g^T [y | z]
        else:
            yz = torch.cat([y, z], dim=1)                  # bias-shaped rows: g^T [y | z]
            gyz = gemm3x_tn(g, yz)                         # (F, 2F), split over atoms to fill the SMs
        g_w[ff:] = gyz[:, f:]
        g_b = torch.cat([gyz[:, :f].reshape(ff), g.sum(dim=0)])
        return g_z, g_w, g_b, g_y, (g if ctx.has_e else None), None, None, None


def hyper_linear(z, weight, bias, y, out_ch, e=None):
    """y_out[n] = reshape(weight z[n] + bias)[:out*in] y[n] + (...)[out*in:]
    (reference HyperLinear.forward + BatchLinear.forward, CGAT/Hypernetworksmp.py:243-254, 205-209).
    `e`: weight[out*in:] z + bias[out*in:] if hyper_trunks has already computed it (fused path only)."""
    in_ch = y.shape[1]
    # fused kernels: F = 128 (tf32 or f16 operands) and F = 256 (BASELINE.json configs[3]; f16 operand kernels only)
    fused = (_FUSED and z.is_cuda and in_ch == out_ch == z.shape[1]
             and (in_ch == 128 or (in_ch == 256 and _F16X3 and _F16X3_GRAD)))
    if fused:
        return _HyperLinear.apply(z, weight, bias, y, e, packed_kmajor(weight, in_ch * out_ch, f16=_F16X3),
                                  packed_kmajor(weight, in_ch * out_ch, 2, f16=_F16X3) if torch.is_grad_enabled()
                                  else None, _F16X3)
    p = torch.addmm(bias, z, weight.t())                            # (N, out*in + out)
    w = p[:, : in_ch * out_ch].view(-1, out_ch, in_ch)
    b = p[:, in_ch * out_ch:]
    return torch.baddbmm(b.unsqueeze(2), w, y.unsqueeze(2)).squeeze(2)


# ----------------------------------------------------------------------------------------------
# Hypernetwork trunks: the J FCBlocks of a node layer share their input (reference
# CGAT/Hypernetworksmp.py:36-83, 243-254); one chained tensor-core kernel runs all of them.
# ----------------------------------------------------------------------------------------------
_TRUNK_STEPS = 5


def gemm3x_tn_batched(a_list, b_list, colsum=True):
    """[a.T @ b for a, b in zip(a_list, b_list)] in one launch (cgat_gemm3x_tn_batched): all operands
    (K, 128*m) row-contiguous with identical shapes.  Returns (C (batch, M, N), column sums of a (batch, M))."""
    K, M = a_list[0].shape
    N = b_list[0].shape[1]
    batch = len(a_list)
    dev = a_list[0].device
    tiles = batch * ((M + 127) // 128) * ((N + 127) // 128)
    n_split = plan_split(K, best_split(tiles, K, min((K + 255) // 256, (2 * 148) // tiles)))
    part = torch.empty((n_split, batch, M, N), dtype=torch.float32, device=dev)
    csum = torch.empty((n_split, batch, M), dtype=torch.float32, device=dev) if colsum else None
    _lib.call("cgat_gemm3x_tn_batched", _lib.ptr_array(a_list), _lib.ptr_array(b_list), batch, a_list[0].stride(0),
              b_list[0].stride(0), _lib.ptr(part), _lib.ptr(csum), M, N, K, n_split, _lib.stream(),
              work=dict(key="gemm3x_tn_batched", bound="tensor", flops=2.0 * batch * M * N * K))
    return sum_parts(part), (sum_parts(csum) if colsum else None)


def _trunk_packed(weights, transpose):
    """Packed chain operands of a list of F x F weights (cgat_hyper_trunk_pack), cached on the first tensor and
    keyed by the autograd versions / addresses of all of them (an optimizer step triggers a repack)."""
    key = (_pack_epoch,) + tuple((w._version, w.data_ptr()) for w in weights)
    cache = weights[0].__dict__.setdefault("_cgat_trunk_packed", {})
    hit = cache.get(transpose)
    if hit is not None and hit[0] == key:
        return hit[1]
    lib = _lib.load()
    f = weights[0].shape[1]
    # fresh buffer on every repack (a pending backward may still hold the previous image; see packed_kmajor)
    buf = torch.empty(int(lib.cgat_hyper_trunk_packed_floats(len(weights), f)), dtype=torch.float32,
                      device=weights[0].device)
    if any(w.stride(1) != 1 for w in weights):
        raise ValueError("trunk weights must be row-contiguous")
    _lib.call("cgat_hyper_trunk_pack", _lib.ptr_array([w.detach() for w in weights]),
              _lib.i64_array([w.stride(0) for w in weights]), len(weights), f, transpose, _lib.ptr(buf), _lib.stream(),
              work=dict(key="hyper_trunk_pack", bound="hbm", bytes=12.0 * len(weights) * f * f))
    cache[transpose] = (key, buf)
    return buf


class _HyperTrunk(torch.autograd.Function):
    """(Z, E) = trunks(h): Z[j] = t4 of hyper-layer j, E[j] = We_j Z[j] + be_j; see cgat_hyper_trunk_fwd.
    flat = per j: W1, b1, W2, b2, W3, b3, W4, b4, We, be (We / be detached: their gradients are produced by
    _HyperLinear, which owns the whole last Linear)."""

    @staticmethod
    def forward(ctx, h, n_j, *flat):
        h = _f32c(h)
        n, f = h.shape
        per = 2 * _TRUNK_STEPS
        ws = [flat[j * per + 2 * s] for j in range(n_j) for s in range(_TRUNK_STEPS)]
        bs = [flat[j * per + 2 * s + 1].contiguous() for j in range(n_j) for s in range(_TRUNK_STEPS)]
        T = torch.empty((n_j, 4, n, f), dtype=torch.float32, device=h.device)
        E = torch.empty((n_j, n, f), dtype=torch.float32, device=h.device)
        _lib.call("cgat_hyper_trunk_fwd", _lib.ptr(h), _lib.ptr(_trunk_packed(ws, 0)), _lib.ptr_array(bs), _lib.ptr(T),
                  _lib.ptr(E), n, f, n_j, _lib.stream(),
                  work=dict(key="hyper_trunk_fwd", bound="tensor", flops=2.0 * n * f * f * _TRUNK_STEPS * n_j,
                            note="3xTF32 chain of 5 GEMMs per hyper-layer, activation tile resident on the SM"))
        ctx.save_for_backward(h, T, *ws)
        ctx.n_j = n_j
        return T[:, 3], E

    @staticmethod
    def backward(ctx, dZ, dE):
        h, T = ctx.saved_tensors[:2]
        ws = ctx.saved_tensors[2:]
        n_j = ctx.n_j
        n, f = h.shape
        dev = h.device
        dZ = torch.zeros((n_j, n, f), dtype=torch.float32, device=dev) if dZ is None else _f32c(dZ)
        dE = torch.zeros((n_j, n, f), dtype=torch.float32, device=dev) if dE is None else _f32c(dE)
        # backward chain order per j: We, W4, W3, W2, W1 (transposed operands)
        wt = [ws[j * _TRUNK_STEPS + s] for j in range(n_j) for s in (4, 3, 2, 1, 0)]
        D = torch.empty((n_j, 4, n, f), dtype=torch.float32, device=dev)
        dH = torch.empty((n_j, n, f), dtype=torch.float32, device=dev)
        _lib.call("cgat_hyper_trunk_bwd", _lib.ptr(dE), _lib.ptr(dZ), _lib.ptr(T), _lib.ptr(_trunk_packed(wt, 1)),
                  _lib.ptr(D), _lib.ptr(dH), n, f, n_j, _lib.stream(),
                  work=dict(key="hyper_trunk_bwd", bound="tensor", flops=2.0 * n * f * f * _TRUNK_STEPS * n_j,
                            note="3xTF32 chain of 5 GEMMs per hyper-layer, activation tile resident on the SM"))
        g_h = sum_parts(dH)
        # weight / bias gradients of the 4 tanh layers of every trunk in one launch: dW_s = D_s^T (input of layer s)
        a_list = [D[j, s] for j in range(n_j) for s in range(4)]
        b_list = [h if s == 0 else T[j, s - 1] for j in range(n_j) for s in range(4)]
        g_w, g_b = gemm3x_tn_batched(a_list, b_list)
        grads = []
        for j in range(n_j):
            for s in range(4):
                grads += [g_w[j * 4 + s], g_b[j * 4 + s]]
            grads += [None, None]
        return (g_h, None, *grads)


def hyper_trunks(h, layers, tails):
    """The trunks of the J hyper-layers of a node layer on their shared hyper-input h.
    layers[j] = [(W, b), ...] tanh layers (reference FCBlock, CGAT/Hypernetworksmp.py:36-83);
    tails[j] = (weight, bias, rows) of the last Linear, whose rows [rows:] are the bias tail.
    Returns (zs, es): z_j and, on the fused path, e_j = weight[rows:] z_j + bias[rows:] (else None)."""
    n_j = len(layers)
    f = h.shape[1]
    fused = (_FUSED and _TRUNK and h.is_cuda and f == 128 and 1 <= n_j <= 4
             and all(len(l) == 4 and all(w.shape == (f, f) for w, _ in l) for l in layers)
             and all(w.shape[0] - rows == f and w.shape[1] == f for w, _, rows in tails))
    if not fused:
        zs = []
        for l in layers:
            z = h
            for w, b in l:
                z = torch.tanh(torch.nn.functional.linear(z, w, b))
            zs.append(z)
        return zs, [None] * n_j
    flat = []
    for l, (w, b, rows) in zip(layers, tails):
        for wi, bi in l:
            flat += [wi, bi]
        flat += [w.detach()[rows:], b.detach()[rows:]]
    Z, E = _HyperTrunk.apply(h, n_j, *flat)
    return list(Z.unbind(0)), list(E.unbind(0))


def _w2_transposed_packed(w2, heads, f16=False):
    """Packed W2^T per head, (H*Hd128, F): row h*Hd128+k, col c = W2[h*F+c, k] — the A operand of the dgrad MMAs
    (tf32 hi/lo images, or fp16 hi / 2^11-scaled lo images for cgat_edge_attn_dgrad_f16).  Hd128 = the hidden width
    rounded up to whole 128-row tiles (zero rows), so that every head starts on a tile boundary."""
    cache = w2.__dict__.setdefault("_cgat_packed", {})
    tag = "w2t16" if f16 else "w2t"
    hit = cache.get(tag)
    if hit is not None and hit[0] == (w2._version, _pack_epoch) and hit[2] == w2.data_ptr():
        return hit[1]
    hf, hd = w2.shape[0], w2.shape[1]
    hd128 = -(-hd // 128) * 128
    wt = w2.detach().view(heads, hf // heads, hd).transpose(1, 2)                       # (H, Hd, F)
    if hd128 != hd:
        wt = torch.nn.functional.pad(wt, (0, 0, 0, hd128 - hd))
    wt = wt.reshape(heads * hd128, hf // heads).contiguous()
    buf = packed_kmajor(wt, f16=f16)   # wt is a temporary: its packed image is a fresh buffer owned by this cache entry
    cache[tag] = ((w2._version, _pack_epoch), buf, w2.data_ptr())
    return buf


def _pad_heads(t, heads, hd, hd_pad):
    """(H*hd, ...) -> (H*hd_pad, ...): zero rows appended to every head's block."""
    if hd_pad == hd:
        return t
    v = t.reshape(heads, hd, *t.shape[1:])
    pad = [0, 0] * (v.dim() - 2) + [0, hd_pad - hd]
    return torch.nn.functional.pad(v, pad).reshape(heads * hd_pad, *t.shape[1:])


def _unpad_heads(t, blocks, hd, hd_pad):
    """(blocks*hd_pad, ...) -> (blocks*hd, ...): the inverse of _pad_heads over `blocks` consecutive head blocks."""
    if hd_pad == hd:
        return t
    return t.reshape(blocks, hd_pad, *t.shape[1:])[:, :hd].reshape(blocks * hd, *t.shape[1:])


def _first_layer_operands(w1a, w1m, b1a, b1m, f, fe, heads, hd_pad):
    """Regrouped first-layer weights of the gate / message MLPs: per-atom block (4*HHd, F) =
    [W1A_i; W1M_i; W1A_j; W1M_j] (+ its packed image when F <= 128), per-rank block (2*HHd, Fe) = [W1A_e; W1M_e] and
    the concatenated bias, with every head's hidden units zero-padded to hd_pad (a multiple of 64: the f16 kernels
    stage 64 hidden units per pipeline step; padded units have zero weights and bias, hence zero activation).
    Cached on the weight and keyed by the autograd versions, so screening inference builds them once and training
    once per optimizer step."""
    key = (_pack_epoch, hd_pad) + tuple((t._version, t.data_ptr()) for t in (w1a, w1m, b1a, b1m))
    cache = w1a.__dict__.setdefault("_cgat_first_layer", {})
    if cache.get("key") == key:
        return cache["val"]
    hhd0 = w1a.shape[0]
    hd = hhd0 // heads
    w1a2 = _pad_heads(w1a.detach().view(hhd0, -1), heads, hd, hd_pad)
    w1m2 = _pad_heads(w1m.detach().view(hhd0, -1), heads, hd, hd_pad)
    hhd = heads * hd_pad
    w_atom = torch.cat([w1a2[:, :f], w1m2[:, :f], w1a2[:, f + fe:], w1m2[:, f + fe:]], dim=0).contiguous()   # (4*HHd, F)
    w_rank = torch.cat([w1a2[:, f:f + fe], w1m2[:, f:f + fe]], dim=0).contiguous()                         # (2*HHd, Fe)
    b1 = torch.cat([_pad_heads(b1a.detach(), heads, hd, hd_pad), _pad_heads(b1m.detach(), heads, hd, hd_pad)])
    packed = None
    if f <= 128:
        lib = _lib.load()
        packed = torch.empty(int(lib.cgat_packed_floats(4 * hhd, f)), dtype=torch.float32, device=w1a.device)
        _lib.call("cgat_pack_kmajor", _lib.ptr(w_atom), w_atom.stride(0), 4 * hhd, f, 0, _lib.ptr(packed), _lib.stream(),
                  work=dict(key="pack_kmajor", bound="hbm", bytes=12.0 * 4 * hhd * f))
    cache["key"], cache["val"] = key, (w_atom, packed, w_rank, b1, w_atom.t().contiguous())
    return cache["val"]


class _EdgeAttentionFused(torch.autograd.Function):
    """Forward: per-atom / per-rank first-layer projections (cgat_gemm3x_nt_res / cgat_gemm3x_nt) + the fused
    gather / second-layer MMA / segmented-softmax kernel (cgat_edge_attn_fwd[_f16]).  No per-edge tensor is written.
    Backward: cgat_edge_attn_bwd_prep[_f16] (per-edge dL/da, dL/dv, LeakyReLU sign masks and the second-layer bias
    gradients), cgat_edge_attn_dgrad[_f16] (per-edge dL/d pre-activation), cgat_edge_attn_reduce (its per-destination /
    per-source / per-rank sums), cgat_edge_attn_wgrad[_f16], and the first-layer products on cgat_gemm3x_nt_splitk /
    cgat_gemm3x_tn.  Deterministic, atomic-free.  `hd_pad`: the hidden width every kernel sees (hidden units of a
    head zero-padded to a multiple of 64 on the f16 path: 426 -> 448 for BASELINE.json configs[3])."""

    @staticmethod
    def forward(ctx, x, edge_table, w1a, b1a, w2a, b2a, w1m, b1m, w2m, b2m, w2a_packed, w2m_packed, plan, heads, f16,
                hd_pad):
        x, edge_table = _f32c(x), _f32c(edge_table)
        ctx.f16 = f16
        n, f = x.shape
        fe = edge_table.shape[1]
        hd = w1a.shape[0] // heads
        hhd = heads * hd_pad
        w_atom, w_atom_packed, w_rank, b1, w_atom_t = _first_layer_operands(w1a, w1m, b1a, b1m, f, fe, heads, hd_pad)
        # (N, 4*HHd): K <= 128 on the resident-tile kernel, wider inputs on the plain 3-pass GEMM
        P = gemm3x_res(x, w_atom_packed, 4 * hhd) if w_atom_packed is not None else gemm3x(x, w_atom)
        T = gemm3x(edge_table, w_rank, b1)                                                           # (K+1, 2*HHd)
        train = any(ctx.needs_input_grad)
        out = torch.empty((n, heads, f), dtype=torch.float32, device=x.device)
        smax = torch.empty_like(out) if train else None
        sden = torch.empty_like(out) if train else None
        e = plan.n_edges
        b2a, b2m = b2a.contiguous(), b2m.contiguous()
        _lib.call("cgat_edge_attn_fwd_f16" if f16 else "cgat_edge_attn_fwd", _lib.ptr(P), _lib.ptr(T),
                  _lib.ptr(plan.rowptr), _lib.ptr(plan.src),
                  _lib.ptr(plan.dst), _lib.ptr(plan.rank), _lib.ptr(w2a_packed), _lib.ptr(w2m_packed),
                  _lib.ptr(b2a), _lib.ptr(b2m), _lib.ptr(out), _lib.ptr(smax), _lib.ptr(sden),
                  n, e, heads, f, hd_pad, 1e-16, _lib.stream(),
                  work=dict(key="edge_attn_fwd", bound="tensor", flops=2.0 * e * heads * hd * 2 * f,
                            bytes=4.0 * (e * (3 + 2 * 2 * hhd) + n * heads * f),
                            note="second MLP layer (E x Hd x F per head, gate+message) on tcgen05, "
                                 + ("f16x3" if f16 else "3xTF32")))
        if train:
            ctx.save_for_backward(x, edge_table, w1a, w1m, w2a, w2m, b2a, b2m, w2a_packed, w2m_packed, P, T, out, smax,
                                  sden, w_atom_t, w_rank)
            ctx.plan, ctx.heads, ctx.hd_pad = plan, heads, hd_pad
        return out

    @staticmethod
    def backward(ctx, g):
        (x, tab, w1a, w1m, w2a, w2m, b2a, b2m, w2a_packed, w2m_packed, P, T, out, smax, sden, w_atom_t,
         w_rank) = ctx.saved_tensors
        plan, heads, hd_pad = ctx.plan, ctx.heads, ctx.hd_pad
        n, f = x.shape
        fe = tab.shape[1]
        hd = w1a.shape[0] // heads          # the real hidden width; hd_pad is what the kernels see
        hhd = heads * hd_pad
        e = plan.n_edges
        kcn = (hd_pad + 31) // 32
        dev = x.device
        lib = _lib.load()
        g = _f32c(g)
        st = _lib.stream()
        # 1. per-edge dL/da, dL/dv and LeakyReLU sign masks
        d_gate = torch.empty((e, heads, f), dtype=torch.float32, device=dev)
        d_msg = torch.empty_like(d_gate)
        signs = torch.empty(2 * heads * kcn * max(e, 1), dtype=torch.int32, device=dev)
        flops2 = 2.0 * e * heads * hd * 2 * f
        # per-CTA column sums of d_msg | d_gate (the second-layer bias gradients) come out of the same kernel
        bsum = (torch.empty if e > 0 else torch.zeros)((int(lib.cgat_edge_attn_grid(e)), 2, heads, f),
                                                      dtype=torch.float32, device=dev)   # e == 0: no launch, no writes
        # max |dL/da|, |dL/dv| (device float, zeroed by the kernel's wrapper): the range of the f16 gradient operands.
        # The tf32 gradient kernels only exist for F = 128 and whole 128-unit hidden tiles.
        f16_grad = ctx.f16 and _F16X3_GRAD and hd_pad <= 512
        if not f16_grad and (f != 128 or hd_pad % 128 or hd_pad > 256):
            raise _lib.CgatLibraryError("edge attention backward: this shape needs the f16 gradient kernels "
                                        "(CGAT_B200_F16X3_EDGE=1, CGAT_B200_F16X3_GRAD=1)")
        dz_amax = torch.empty(1, dtype=torch.float32, device=dev) if f16_grad else None
        _lib.call("cgat_edge_attn_bwd_prep_f16" if ctx.f16 else "cgat_edge_attn_bwd_prep", _lib.ptr(P), _lib.ptr(T),
                  _lib.ptr(plan.rowptr), _lib.ptr(plan.src),
                  _lib.ptr(plan.dst), _lib.ptr(plan.rank), _lib.ptr(w2a_packed), _lib.ptr(w2m_packed), _lib.ptr(b2a),
                  _lib.ptr(b2m), _lib.ptr(out), _lib.ptr(smax), _lib.ptr(sden), _lib.ptr(g), _lib.ptr(d_gate),
                  _lib.ptr(d_msg), _lib.ptr(signs), _lib.ptr(bsum), _lib.ptr(dz_amax), n, e, heads, f, hd_pad, 1e-16, st,
                  work=dict(key="edge_attn_bwd_prep", bound="tensor", flops=flops2))
        # 2. dgrad on the tensor cores -> per-edge d_pre, then its per-destination / per-source / per-rank sums
        #    (HBM-bound, cgat_edge_attn_reduce)
        wt_a, wt_m = _w2_transposed_packed(w2a, heads, f16_grad), _w2_transposed_packed(w2m, heads, f16_grad)
        n_ranks = tab.shape[0]
        d_pre = torch.empty((e, 2 * hhd), dtype=torch.float32, device=dev)
        if f16_grad:
            _lib.call("cgat_edge_attn_dgrad_f16", _lib.ptr(d_gate), _lib.ptr(d_msg), _lib.ptr(signs),
                      _lib.ptr(plan.rowptr), _lib.ptr(plan.dst), _lib.ptr(wt_a), _lib.ptr(wt_m), _lib.ptr(dz_amax),
                      _lib.ptr(d_pre), n, e, heads, f, hd_pad, st,
                      work=dict(key="edge_attn_dgrad", bound="tensor", flops=flops2,
                                note="f16x3, gradient operand scaled by 2^k"))
        else:
            _lib.call("cgat_edge_attn_dgrad", _lib.ptr(d_gate), _lib.ptr(d_msg), _lib.ptr(signs), _lib.ptr(plan.rowptr),
                      _lib.ptr(plan.dst), None, None, _lib.ptr(wt_a), _lib.ptr(wt_m), None, 4 * hhd, 0, None,
                      n_ranks, _lib.ptr(d_pre), n, e, heads, f, hd_pad, st,
                      work=dict(key="edge_attn_dgrad", bound="tensor", flops=flops2))
        so = plan.by_source()
        chunks, groups = int(lib.cgat_edge_attn_reduce_chunks(n)), int(lib.cgat_edge_attn_reduce_groups(n))
        d_p = torch.empty((n, 4 * hhd), dtype=torch.float32, device=dev)
        d_rank = torch.empty((groups, n_ranks, 2 * hhd), dtype=torch.float32, device=dev)
        rank_scratch = torch.empty((chunks, n_ranks, 2 * hhd), dtype=torch.float32, device=dev)
        counters = torch.empty(int(lib.cgat_edge_attn_reduce_counters(n, 2 * hhd)), dtype=torch.int32, device=dev)
        _lib.call("cgat_edge_attn_reduce", _lib.ptr(d_pre), 2 * hhd, _lib.ptr(plan.rowptr), _lib.ptr(so.rowptr),
                  _lib.ptr(so.row), _lib.ptr(so.rank), _lib.ptr(d_p), 4 * hhd, 0, 2 * hhd, _lib.ptr(d_rank),
                  _lib.ptr(rank_scratch), _lib.ptr(counters), n_ranks, n, 2 * hhd, st,
                  work=dict(key="edge_attn_reduce", bound="hbm",
                            bytes=4.0 * (2 * e * 2 * hhd + n * 4 * hhd + groups * n_ranks * 2 * hhd),
                            note="reads d_pre twice (by destination, by source), writes dL/dP + per-rank partials"))
        del d_pre
        d_t = sum_parts(d_rank)                                                     # (K+1, 2*HHd)
        # 3. second-layer weight / bias gradients
        if f16_grad:
            splits = int(lib.cgat_edge_attn_wgrad_f16_splits(heads, f, hd_pad))
            part = torch.empty((splits, 2, heads, f, hd_pad), dtype=torch.float32, device=dev)
            _lib.call("cgat_edge_attn_wgrad_f16", _lib.ptr(P), _lib.ptr(T), _lib.ptr(plan.src), _lib.ptr(plan.dst),
                      _lib.ptr(plan.rank), _lib.ptr(d_gate), _lib.ptr(d_msg), _lib.ptr(dz_amax), _lib.ptr(part), e, heads,
                      f, hd_pad, st, work=dict(key="edge_attn_wgrad", bound="tensor", flops=flops2,
                                               note="f16x3, gradient operand scaled by 2^k"))
        else:
            splits = int(lib.cgat_edge_attn_wgrad_splits(heads))
            part = torch.empty((splits, 2, heads, f, hd_pad), dtype=torch.float32, device=dev)
            _lib.call("cgat_edge_attn_wgrad", _lib.ptr(P), _lib.ptr(T), _lib.ptr(plan.src), _lib.ptr(plan.dst),
                      _lib.ptr(plan.rank), _lib.ptr(d_gate), _lib.ptr(d_msg), _lib.ptr(part), e, heads, f, hd_pad, st,
                      work=dict(key="edge_attn_wgrad", bound="tensor", flops=flops2))
        d_w2 = sum_parts(part)                                                      # (2, H, F, hd_pad)
        if hd_pad != hd:
            d_w2 = d_w2[..., :hd]
        g_w2a, g_w2m = d_w2[0].reshape(w2a.shape), d_w2[1].reshape(w2m.shape)
        g_b2 = sum_parts(bsum)
        g_b2m, g_b2a = g_b2[0].reshape(-1), g_b2[1].reshape(-1)
        # 4. first layer: P = x w_atom^T, T = tab w_rank^T + b1 (gradients of the zero-padded hidden units are dropped)
        g_x = gemm3x_splitk(d_p, w_atom_t)
        g_watom = _unpad_heads(gemm3x_tn(d_p, x), 4 * heads, hd, hd_pad)            # (4*H*Hd, F)
        g_tab = d_t @ w_rank
        g_wrank = _unpad_heads(d_t.t() @ tab, 2 * heads, hd, hd_pad)                # (2*H*Hd, Fe)
        g_b1 = _unpad_heads(d_t.sum(dim=0), 2 * heads, hd, hd_pad)
        hh = heads * hd
        g_w1a = torch.cat([g_watom[:hh], g_wrank[:hh], g_watom[2 * hh:3 * hh]], dim=1).reshape(w1a.shape)
        g_w1m = torch.cat([g_watom[hh:2 * hh], g_wrank[hh:], g_watom[3 * hh:]], dim=1).reshape(w1m.shape)
        return (g_x, g_tab, g_w1a, g_b1[:hh], g_w2a, g_b2a, g_w1m, g_b1[hh:], g_w2m, g_b2m, None, None, None, None,
                None, None)


def edge_attention_heads_unfused(x, edge_table, plan, w1a, b1a, w2a, b2a, w1m, b1m, w2m, b2m, heads):
    """(N, H, F) per-head aggregation through the unfused formulation (see edge_attention_unfused)."""
    n, f = x.shape
    fe = edge_table.shape[1]
    hhd = w1a.shape[0]
    hd = hhd // heads
    w1 = torch.cat([w1a, w1m], dim=0)
    p_dst = x @ w1[:, :f].t()
    p_src = x @ w1[:, f + fe:].t()
    t_rank = torch.addmm(torch.cat([b1a, b1m]), edge_table, w1[:, f:f + fe].t())
    pre = p_dst.index_select(0, plan.dst) + p_src.index_select(0, plan.src) + t_rank.index_select(0, plan.rank)
    hid = torch.nn.functional.leaky_relu(pre, LEAKY_SLOPE)
    e = hid.shape[0]
    hid_a = hid[:, :hhd].view(e, heads, hd).transpose(0, 1)
    hid_m = hid[:, hhd:].view(e, heads, hd).transpose(0, 1)
    gate = torch.baddbmm(b2a.view(heads, 1, -1), hid_a, w2a.transpose(1, 2)).transpose(0, 1).contiguous()
    msg = torch.baddbmm(b2m.view(heads, 1, -1), hid_m, w2m.transpose(1, 2)).transpose(0, 1).contiguous()
    return seg_softmax(gate, msg, ptr=plan.rowptr, seg_of_row=plan.dst, n_seg=n, eps=1e-16)


def edge_attention(x, edge_table, plan, mh_a, mh_m, heads, edge_ids=None):
    """(N, F): mean over heads of the attention-weighted messages arriving at each atom (reference
    GATConvNodes.message + aggregate + the head-mean of update, CGAT/CGAT.py:319-329).  `mh_a`, `mh_m` are
    the gate / message MultiHeadNetwork modules (parameters fc_in/fc_out in the reference's Conv1d layout).
    edge_ids (no_hyper=False only): edge_table holds per-EDGE features and row edge_ids[t] belongs to the t-th edge of
    the destination-sorted order; that variant runs the unfused formulation (the fused kernel gathers a per-rank table)."""
    f = x.shape[1]
    hd = mh_a.hidden_dim
    # scalar attention (vector_attention=False, reference CGAT/CGAT.py:282-287: the gate net emits ONE logit per head
    # that weighs all channels): run as vector attention whose gate rows are that one row repeated — `expand` is an
    # autograd op, so the F per-channel gate gradients the kernels produce are summed back onto the single row
    scalar_gate = mh_a.output_dim == 1 and f > 1
    base_ok = (_FUSED and edge_ids is None and x.is_cuda and (mh_a.output_dim == f or scalar_gate)
               and mh_m.output_dim == f and heads <= 8 and edge_table.shape[1] % 4 == 0 and edge_table.shape[0] <= 32)
    # the original instantiation: F = 128, whole 128-unit hidden tiles (tf32 kernels, or f16 forward + either backward)
    small_ok = base_ok and f == 128 and hd % 128 == 0 and hd <= 256
    # generalised f16 path (BASELINE.json configs[3]: F = 256, 8 heads, Hd = 426): F in {128, 256} run as F / 128
    # virtual heads, hidden units zero-padded to a multiple of 64 (426 -> 448), gradient kernels on kind::f16 too
    hd_pad = -(-hd // 64) * 64
    wide_ok = (base_ok and _F16X3_EDGE and _F16X3_GRAD and f in (128, 256) and heads * (f // 128) <= 16
               and hd_pad <= 512)
    if small_ok or wide_ok:
        f16 = (_F16X3_EDGE and hd % 64 == 0) if small_ok else True
        pk = dict(f16=True, pre_scale=_EDGE_W2_PRESCALE) if f16 else {}
        w2a, b2a = mh_a.fc_out.weight, mh_a.fc_out.bias
        if scalar_gate:
            w2a = w2a.view(heads, 1, hd, 1).expand(heads, f, hd, 1).reshape(heads * f, hd, 1)
            b2a = b2a.view(heads, 1).expand(heads, f).reshape(heads * f)
        out = _EdgeAttentionFused.apply(x, edge_table, mh_a.fc_in.weight, mh_a.fc_in.bias, w2a, b2a,
                                        mh_m.fc_in.weight, mh_m.fc_in.bias, mh_m.fc_out.weight,
                                        mh_m.fc_out.bias, packed_kmajor(w2a, **pk),
                                        packed_kmajor(mh_m.fc_out.weight, **pk), plan, heads, f16,
                                        hd_pad if f16 else hd)
        return out.mean(dim=1)
    return edge_attention_unfused(x, edge_table, plan, mh_a.w_in(), mh_a.fc_in.bias, mh_a.w_out(), mh_a.fc_out.bias,
                                  mh_m.w_in(), mh_m.fc_in.bias, mh_m.w_out(), mh_m.fc_out.bias, heads, edge_ids)
