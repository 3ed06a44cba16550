"""-m "not gpu": the arithmetic claim behind the kind::f16 kernels (tc_common.cuh), emulated in
numpy against fp64: an fp16 hi/lo pair carries the same 22 significand bits as the tf32 hi/lo pair as long as the lo
part stays in fp16's normal range — (a) lo scaled by 2^11 with its own accumulator (hyper_f16.cu), (b) operands
pre-scaled by powers of two with one accumulator (edge_attn_fwd.cu, kF16)."""
import numpy as np
import pytest


def _tf32(x):
    u = np.asarray(x, np.float32).view(np.uint32).astype(np.uint64)
    return ((u + 0x1000) & 0xFFFFE000).astype(np.uint32).view(np.float32)   # the integer form used on the device


def _split_tf32(x):
    hi = _tf32(x)
    return hi.astype(np.float64), _tf32((x - hi).astype(np.float32)).astype(np.float64), 1.0


def _split_f16(x, pre=1.0, lo_scale=2048.0):
    xs = (x * np.float32(pre)).astype(np.float32)
    hi = xs.astype(np.float16).astype(np.float32)
    lo = ((xs - hi) * np.float32(lo_scale)).astype(np.float16)
    return hi.astype(np.float64) / pre, lo.astype(np.float64) / pre, 1.0 / lo_scale


def _three_pass(a, b, sa, sb):
    ah, al, ia = sa(a)
    bh, bl, ib = sb(b)
    return ah @ bh.T + ia * (al @ bh.T) + ib * (ah @ bl.T)


def _rel_l2(r, ref):
    return float(np.sqrt(((r - ref) ** 2).mean()) / np.sqrt((ref ** 2).mean()))


def test_integer_tf32_rounding_matches_round_to_nearest_away():
    rng = np.random.default_rng(0)
    x = (rng.standard_normal(100000) * 10.0 ** rng.integers(-20, 20, 100000)).astype(np.float32)
    hi = _tf32(x)
    ulp = np.abs(x) * 2.0 ** -10
    assert np.all(np.abs(hi.astype(np.float64) - x) <= 0.5 * ulp * (1 + 1e-6))            # nearest
    assert np.all((hi.view(np.uint32) & 0x1FFF) == 0)                                      # 10 explicit mantissa bits


@pytest.mark.parametrize("scale_a", [1.0, 1e-2, 1e-4, 1e2])
def test_scaled_lo_split_matches_tf32_split(scale_a):
    """(a): hyper-linear operands — tanh outputs times kaiming-scaled weights — and rescaled variants."""
    rng = np.random.default_rng(1)
    a = (scale_a * np.tanh(rng.standard_normal((256, 128)))).astype(np.float32)
    b = (0.0125 * rng.standard_normal((128, 128))).astype(np.float32)
    ref = a.astype(np.float64) @ b.astype(np.float64).T
    e_tf32 = _rel_l2(_three_pass(a, b, _split_tf32, _split_tf32), ref)
    e_f16 = _rel_l2(_three_pass(a, b, _split_f16, _split_f16), ref)
    e_fp32 = _rel_l2((a @ b.T).astype(np.float64), ref)
    assert e_tf32 < 1.2e-7
    if scale_a >= 1e-2:
        assert e_f16 < 1.1 * e_tf32 + 1e-9, (e_tf32, e_f16)       # indistinguishable in fp16's normal range
    # at 1e-4 the hi parts approach fp16's subnormals and the error doubles — still below a plain fp32 GEMM's rounding
    assert e_f16 < e_fp32, (e_f16, e_fp32)


def test_prescaled_single_accumulator_split():
    """(b): LeakyReLU hidden activations (x 2^4) times W2 (x 2^6), unscaled lo parts, one accumulator."""
    rng = np.random.default_rng(2)
    pre = rng.standard_normal((256, 256)) * 1.5
    hid = np.where(pre > 0, pre, 0.01 * pre).astype(np.float32)
    w2 = rng.uniform(-1 / 16, 1 / 16, (128, 256)).astype(np.float32)
    ref = hid.astype(np.float64) @ w2.astype(np.float64).T
    r = _three_pass(hid, w2, lambda x: _split_f16(x, 16.0, 1.0), lambda x: _split_f16(x, 64.0, 1.0))
    e = _rel_l2(r, ref)
    e_tf32 = _rel_l2(_three_pass(hid, w2, _split_tf32, _split_tf32), ref)
    assert e < 2e-7 and e < 3 * e_tf32, (e, e_tf32)
    assert np.abs(hid * 16).max() < 65504 and np.abs(w2 * 64).max() < 65504


def test_unscaled_lo_without_prescale_is_why_the_scales_exist():
    """Without either remedy the lo parts of typical weights are fp16 subnormals and the product loses ~3 bits."""
    rng = np.random.default_rng(3)
    a = np.tanh(rng.standard_normal((256, 128))).astype(np.float32)
    b = (0.0125 * rng.standard_normal((128, 128))).astype(np.float32)
    ref = a.astype(np.float64) @ b.astype(np.float64).T
    bad = _rel_l2(_three_pass(a, b, lambda x: _split_f16(x, 1.0, 1.0), lambda x: _split_f16(x, 1.0, 1.0)), ref)
    good = _rel_l2(_three_pass(a, b, _split_f16, _split_f16), ref)
    assert bad > 3 * good
