"""CPU model of the numerics the tensor-core engine relies on: fp16 operands with fp16 residual planes,
fp32 accumulation.  No GPU, no product code: these tests pin the arithmetic claims the kernels are built on, so that a change of
the split scheme has a reference to be checked against.

  x ~ hi + lo with hi = fp16(x), lo = fp16(x - hi);   x*w ~ x_hi*w_hi + x_lo*w_hi + x_hi*w_lo   (lo*lo ~ 2^-22 dropped)
"""
import numpy as np
import pytest


def split(x):
    hi = x.astype(np.float16)
    lo = (x - hi.astype(np.float32)).astype(np.float16)
    return hi, lo


def dot32(a16, b16):
    """products of fp16 values are exact in fp32; accumulate in float64 here to isolate the operand error from summation order"""
    return (a16.astype(np.float64) * b16.astype(np.float64)).sum(axis=-1)


@pytest.mark.parametrize("k", [144, 576, 4096])
def test_three_term_product_is_fp32_grade(k):
    g = np.random.default_rng(k)
    x = np.maximum(g.standard_normal((256, k)).astype(np.float32), 0) * 2.0          # post-ReLU activations
    w = (g.standard_normal((k,)) * 0.05).astype(np.float32)
    exact = (x.astype(np.float64) * w.astype(np.float64)).sum(axis=1)
    xh, xl = split(x)
    wh, wl = split(w * 8192.0)                                                       # weights stored times a power of two
    plain = dot32(xh, wh) / 8192.0
    sw = (dot32(xh, wh) + dot32(xh, wl)) / 8192.0                                    # weight residual only (AffNet layers 2-6)
    full = (dot32(xh, wh) + dot32(xl, wh) + dot32(xh, wl)) / 8192.0                  # both residuals (OriNet, layer 1, heads)
    scale = np.abs(x).astype(np.float64) @ np.abs(w).astype(np.float64)
    e_plain, e_sw, e_full = [np.abs(v - exact).max() / scale.max() for v in (plain, sw, full)]
    assert e_plain > 1e-5 and e_plain > 30 * e_full   # plain fp16 operands: ~2^-12 per term (averaging over K helps a little)
    assert e_sw < e_plain                     # the activation rounding remains
    assert e_full < 4e-7                      # ~2^-22: fp32-grade


def test_power_of_two_scaling_keeps_weight_residuals_normal():
    """A BatchNorm-folded weight of ~1e-2 has a residual of ~5e-6: below fp16's smallest normal 6.1e-5 it keeps only a few bits.
    Stored times 2^k with the largest weight near 2^13 every residual keeps its 11 bits."""
    g = np.random.default_rng(5)
    w = (g.standard_normal(4096) * 1e-2).astype(np.float32)
    wh, wl = split(w)
    err_plain = np.abs((wh.astype(np.float64) + wl.astype(np.float64)) - w).max() / np.abs(w).max()
    ex = 13 - int(np.frexp(np.abs(w).max())[1])
    s = np.float32(2.0 ** ex)
    sh, sl = split(w * s)
    err_scaled = np.abs((sh.astype(np.float64) + sl.astype(np.float64)) / s - w).max() / np.abs(w).max()
    assert np.abs(sh.astype(np.float32)).max() < 2 ** 14 < 65504
    assert err_scaled < 2.0 ** -21 and err_plain > 4 * err_scaled


def test_truncating_accumulation_grows_with_chain_length_and_groups_help():
    """Model of what was measured on the heads (K = 4096): an accumulator that rounds toward zero after every add loses
    accuracy with the length of the running sum; eight interleaved partial accumulators added at the end recover most of it."""
    g = np.random.default_rng(9)
    k = 4096
    x = np.maximum(g.standard_normal((64, k)), 0).astype(np.float32)
    w = (g.standard_normal(k) * 0.02).astype(np.float32)
    p = (x * w).astype(np.float32)                                # per-term products, exact enough for the model
    exact = p.astype(np.float64).sum(axis=1)

    def trunc_add(acc, v):
        s = acc.astype(np.float64) + v.astype(np.float64)
        f = s.astype(np.float32)
        over = np.abs(f.astype(np.float64)) > np.abs(s)           # round-to-nearest went away from zero: step back one ulp
        return np.where(over, np.nextafter(f, np.float32(0)), f).astype(np.float32)

    one = np.zeros(64, np.float32)
    for j in range(k):
        one = trunc_add(one, p[:, j])
    groups = np.zeros((8, 64), np.float32)
    for j in range(k):
        groups[(j // 4) % 8] = trunc_add(groups[(j // 4) % 8], p[:, j])
    eight = groups.astype(np.float64).sum(axis=0)
    e1, e8 = np.abs(one - exact).max(), np.abs(eight - exact).max()
    assert e8 < 0.6 * e1
