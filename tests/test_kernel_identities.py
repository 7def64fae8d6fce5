"""Integer / floating-point identities the CUDA texel pass relies on (dp_group.cuh,
dp_device.cuh), checked with numpy on the CPU: each one lets the kernel drop instructions
without changing a single bit of OpenCV's fixed-point bilinear arithmetic."""
from fractions import Fraction

import numpy as np


def _taps(n, rng):
    p = [rng.integers(0, 1 << 24, n, dtype=np.uint64).astype(np.uint32) for _ in range(2)]
    for a in p:                      # extremes: all channels 255 / 0 (the x byte is always 0)
        a[:64] = 0x00FFFFFF
        a[64:128] = 0
    w1 = rng.integers(0, 32, n).astype(np.uint32)
    w1[:32], w1[32:64] = 31, 0
    return p[0], p[1], np.uint32(32) - w1, w1


def test_green_from_word_blend_minus_blue_red():
    """G << 8 part of the horizontal blend = blend of the whole BGRx word minus its B | R part."""
    rng = np.random.default_rng(0)
    a, b, w0, w1 = _taps(1 << 20, rng)
    m_br, m_g = np.uint32(0x00FF00FF), np.uint32(0xFF00)
    with np.errstate(over="ignore"):
        br = (a & m_br) * w0 + (b & m_br) * w1
        g_masked = (a & m_g) * w0 + (b & m_g) * w1
        word = a * w0 + b * w1
        assert (word.astype(np.uint64) == a.astype(np.uint64) * w0 + b.astype(np.uint64) * w1).all(), \
            "the word blend must not wrap (x byte = 0, partial sums <= 255 * 32)"
        assert np.array_equal(word - br, g_masked)
        # a tap with weight 0 may be any word (the neighbour of an ROI edge pixel)
        junk = rng.integers(0, 1 << 32, a.size, dtype=np.uint64).astype(np.uint32)
        z, full = np.zeros_like(w0), np.full_like(w0, 32)
        assert np.array_equal((a * full + junk * z) - ((a & m_br) * full + (junk & m_br) * z),
                              (a & m_g) * full)


def test_separable_weights_equal_opencv_15_bit_weights():
    """(sum of tap * (32-ax)(32-ay)*32 ... + 2^14) >> 15 == separable form with + 2^9 >> 10."""
    rng = np.random.default_rng(1)
    n = 1 << 18
    t = rng.integers(0, 256, (4, n)).astype(np.int64)
    ax, ay = rng.integers(0, 32, n), rng.integers(0, 32, n)
    w = [(32 - ax) * (32 - ay) * 32, ax * (32 - ay) * 32, (32 - ax) * ay * 32, ax * ay * 32]
    ocv = (sum(t[k] * w[k] for k in range(4)) + (1 << 14)) >> 15
    top, bot = t[0] * (32 - ax) + t[1] * ax, t[2] * (32 - ax) + t[3] * ax
    sep = (top * (32 - ay) + bot * ay + 512) >> 10
    assert np.array_equal(ocv, sep)


def test_clamp_orders_agree_for_non_negative_bound():
    """min(max(x, 0), m) == max(min(x, m), 0) for m >= 0 (one VIMNMX.RELU in SASS)."""
    rng = np.random.default_rng(2)
    x = rng.integers(-(1 << 31), 1 << 31, 1 << 16)
    for m in (0, 31, 32 * 7, 1 << 20):
        assert np.array_equal(np.minimum(np.maximum(x, 0), m), np.maximum(np.minimum(x, m), 0))


def test_product_of_two_floats_is_exact_in_double():
    """num + da*db with the product exact => one DFMA rounds like DMUL followed by DADD."""
    rng = np.random.default_rng(3)
    a = (rng.integers(0, 256, 4096) - rng.uniform(0, 255, 4096)).astype(np.float32)
    b = (rng.integers(0, 256, 4096) - rng.uniform(0, 255, 4096)).astype(np.float32)
    prod = a.astype(np.float64) * b.astype(np.float64)
    for x, y, p in zip(a[:512], b[:512], prod[:512]):
        assert Fraction(float(x)) * Fraction(float(y)) == Fraction(float(p))


def test_last_pass_is_the_only_ragged_one():
    """GL * (NP - 1) < s*s <= GL * NP for every group geometry: only the last texel pass can
    hold lanes past the patch, so the mask is applied there alone."""
    for s in range(2, 17):
        gl = 4 if s <= 8 else (8 if s <= 12 else 16)
        np_ = (s * s + gl - 1) // gl
        assert gl * (np_ - 1) < s * s <= gl * np_
