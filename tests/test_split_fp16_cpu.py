"""CPU check of the error model behind the single-pass tensor matcher (csrc/match_tc.cu): the split-FP16
representation hi = fp16(2^12 x), lo = fp16(2^12 x - hi) with the three products hi*hi + hi*lo + lo*hi
reproduces the reference's sequential FP32 FMA chain (matching.cu:338-351) to well within
eps = T3_C1 * |a| * |b|.  (The tensor cores' own accumulation error is part of T3_C1's budget and is
covered on the GPU by the bit-exact comparisons of tests/test_match_gpu.py.)"""
import numpy as np

from cudasift_b200.synth import synth_descriptors

T3_C1 = 4.0e-5          # csrc/match_tc.cu
SCALE = 4096.0


def chain_fp32(a, b):
    """sum_k fma(a[k], b[k], acc) in FP32, k = 0..127 (float64 holds every product exactly)."""
    acc = np.zeros(a.shape[0], np.float32)
    for k in range(128):
        acc = (a[:, k].astype(np.float64) * b[:, k].astype(np.float64) + acc.astype(np.float64)).astype(np.float32)
    return acc


def split(x):
    xs = x.astype(np.float32) * np.float32(SCALE)
    hi = xs.astype(np.float16)
    lo = (xs - hi.astype(np.float32)).astype(np.float16)
    return hi.astype(np.float64), lo.astype(np.float64)


def test_split_fp16_error_bound():
    s1, s2 = synth_descriptors(4096, 1), synth_descriptors(4096, 2)
    a, b = s1["data"], s2["data"]
    exact = chain_fp32(a, b).astype(np.float64)
    ah, al = split(a)
    bh, bl = split(b)
    approx = ((ah * bh).sum(1) + (ah * bl).sum(1) + (al * bh).sum(1)) / (SCALE * SCALE)
    na = np.sqrt((a.astype(np.float64) ** 2).sum(1))
    nb = np.sqrt((b.astype(np.float64) ** 2).sum(1))
    err = np.abs(approx - exact) / (na * nb)
    # inputs + dropped lo*lo + the chain's own rounding: at most 3*2^-22 + 2^-17 of the 4.0e-5 budget
    assert err.max() < 3 * 2.0 ** -22 + 2.0 ** -17, err.max()
    assert err.max() < 0.25 * T3_C1
    # plain FP16 (the first version of the matcher) is ~2^-11 per operand: three orders of magnitude worse
    plain = (a.astype(np.float16).astype(np.float64) * b.astype(np.float16).astype(np.float64)).sum(1)
    assert (np.abs(plain - exact) / (na * nb)).max() > 20 * err.max()


def test_split_is_exact_sum_for_small_values():
    """x = hi + lo up to 2^-22 relative, and the 2^12 scale keeps lo out of the FP16 subnormals for the
    value range of SIFT descriptors (>= 1e-4 where it matters)."""
    rng = np.random.default_rng(3)
    x = rng.uniform(1e-4, 0.5, 100000).astype(np.float32)
    hi, lo = split(x)
    rel = np.abs((hi + lo) / SCALE - x.astype(np.float64)) / x
    assert rel.max() < 2.0 ** -21
    assert (np.abs(lo[lo != 0]) >= 2.0 ** -14).mean() > 0.99      # FP16 normal range
