"""The arithmetic claims of the fp32 parity mode's tensor-core MLP (csrc/mlp_tc32_sm100.cuh), checked on the CPU:
(1) the two-way IEEE-half split x = hi + 2^-11 lo' (hi = fp16(x), lo' = fp16(2^11 (x - hi))) carries 22 bits, and three products
    (Wh Ah, Wh Al', Wl' Ah) reproduce a K = 256 fp32 layer product to the level of a plain fp32 GEMM;
(2) the kernel's sine (Cody-Waite reduction by pi with two constants + odd degree-9 polynomial, coefficients parsed from the
    source so that they cannot drift apart) is accurate to 1.5e-7 on the argument range of the network.
Test infrastructure only: nothing here is on the product path."""
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "cips-3dplusplus_b200", "csrc", "mlp_tc32_sm100.cuh")


def _split(x):
    hi = x.astype(np.float16)
    lo = ((x - hi.astype(np.float32)) * np.float32(2048.0)).astype(np.float16)
    return hi, lo


def test_two_way_half_split_carries_22_bits():
    rng = np.random.default_rng(0)
    # activations (sines), SIREN hidden weights (|w| <= sqrt(6/256)/25) and larger trained-size weights
    for x in (np.sin(rng.uniform(-60, 60, 1 << 16)), rng.uniform(-0.0061, 0.0061, 1 << 16), rng.standard_normal(1 << 16)):
        x = x.astype(np.float32)
        hi, lo = _split(x)
        assert np.isfinite(hi.astype(np.float32)).all() and np.isfinite(lo.astype(np.float32)).all()
        recon = hi.astype(np.float64) + lo.astype(np.float64) / 2048.0
        err = np.abs(recon - x.astype(np.float64))
        # 11 + 11 bits: the scaled low part is a NORMAL fp16 wherever the value itself is (|x| > 2^-14), so the bound is
        # relative; below that it is the fp16 subnormal spacing of the low part, 2^-24 / 2^11
        assert (err <= np.abs(x) * 2.0 ** -22 + 2.0 ** -35).all(), float((err / np.maximum(np.abs(x), 1e-30)).max())


def test_three_half_products_match_an_fp32_gemm():
    rng = np.random.default_rng(1)
    K = N = 256
    a = np.sin(rng.uniform(-30, 30, (128, K))).astype(np.float32)                 # activations
    w = rng.uniform(-0.0061, 0.0061, (N, K)).astype(np.float32)                   # hidden-layer weights (frequency_init(25))
    exact = a.astype(np.float64) @ w.astype(np.float64).T
    ah, al = _split(a)
    wh, wl = _split(w)
    f = lambda t: t.astype(np.float64)
    acc0 = f(ah) @ f(wh).T                                                        # main accumulator (fp32 in TMEM; exact here)
    acc1 = f(al) @ f(wh).T + f(ah) @ f(wl).T                                      # small accumulator, carries the 2^11 scale
    got = acc0 + acc1 / 2048.0
    scale = np.sqrt((exact ** 2).mean())
    err_split = np.sqrt(((got - exact) ** 2).mean()) / scale
    err_fp32 = np.sqrt((((a @ w.T).astype(np.float64) - exact) ** 2).mean()) / scale
    print(f"rms error / rms value: three half products {err_split:.2e}, numpy fp32 GEMM {err_fp32:.2e}")
    assert err_split < 1e-7                    # the dropped Wl Al term and the split rounding: 2^-22-level, averaged over K
    assert err_split < 2.0 * err_fp32 + 5e-8   # no worse than what an fp32 GEMM's own rounding leaves


def _kernel_sine_constants():
    src = open(SRC).read()
    body = src[src.index("float sin_poly(float x)"):]
    body = body[:body.index("}")]
    nums = [float(m) for m in re.findall(r"(-?\d+\.\d+(?:e[+-]?\d+)?)f", body)]
    # order of appearance: 1/pi, magic, magic, -pi_hi, -pi_lo(positive literal), c9, c7, c5, c3
    inv_pi, magic, _, neg_pi_hi, pi_lo_corr, c9, c7, c5, c3 = nums[:9]
    return np.float32(inv_pi), np.float32(magic), np.float32(neg_pi_hi), np.float32(pi_lo_corr), [np.float32(c) for c in (c9, c7, c5, c3)]


def test_kernel_sine_polynomial_accuracy():
    inv_pi, magic, neg_pi_hi, pi_lo_corr, (c9, c7, c5, c3) = _kernel_sine_constants()
    assert abs(float(inv_pi) - 1 / np.pi) < 1e-7 and float(magic) == 12582912.0
    assert abs(float(neg_pi_hi) + np.pi) < 1e-6                                   # -float(pi)
    assert abs(float(neg_pi_hi) + float(pi_lo_corr) + np.pi) < 1e-14              # + (float(pi) - pi): together -pi to 1e-15
    x = np.linspace(-100.0, 100.0, 2_000_001).astype(np.float32)
    f32 = np.float32
    fma = lambda a, b, c: (a.astype(np.float64) * np.float64(b) + np.asarray(c, np.float64)).astype(np.float32)   # one rounding
    t = fma(x, inv_pi, magic)
    n = (t - magic).astype(np.float32)
    r = fma(n, neg_pi_hi, x)
    r = fma(n, pi_lo_corr, r)
    odd = (t.view(np.uint32) & 1).astype(bool)
    r = np.where(odd, -r, r).astype(np.float32)
    s = (r * r).astype(np.float32)
    q = fma(s, c9, c7)
    fma3 = lambda a, b, c: (a.astype(np.float64) * b.astype(np.float64) + np.float64(c)).astype(np.float32)
    q = fma3(q, s, c5)
    q = fma3(q, s, c3)
    rs = (r * s).astype(np.float32)
    y = (rs.astype(np.float64) * q.astype(np.float64) + r.astype(np.float64)).astype(np.float32)
    err = np.abs(y.astype(np.float64) - np.sin(x.astype(np.float64)))
    print(f"kernel sine on [-100, 100]: max abs error {err.max():.3e}, rms {np.sqrt((err ** 2).mean()):.3e}")
    assert err.max() < 1.5e-7
