"""CPU check of the numerical CLAIM the device's fast epilogues rest on (tengine_b200/csrc/common.cuh):

    the fast value t (one multiply by a precomputed constant instead of the reference's multiply-multiply-divide chain)
    rounds to the reference's integer whenever t is farther than 2^-13 from a half-integer; elements inside that guard band
    are the only ones that need the literal arithmetic.

The reference recipes are restated here in numpy float32 (every numpy float32 operation rounds like the C code's float
operation), the fast path likewise; FMA is emulated in float64 (the product of two float32 is exact in float64; the extra
rounding of the sum is far below the guard margin).  No GPU, no oracle library: this pins the mathematics, the GPU tests pin
the kernels."""
import numpy as np

F = np.float32
MAGIC = F(12582912.0)
EPS = F(2.0 ** -13)


def round_half_away(x):
    return np.where(x >= 0, np.floor(x + F(0.5)), np.ceil(x - F(0.5))).astype(np.float32)


def guard_and_round(t):
    r = (t + MAGIC).astype(np.float32)
    d = (t - (r - MAGIC).astype(np.float32)).astype(np.float32)
    q = (r.view(np.int32) - MAGIC.view(np.int32)).astype(np.int64)
    return q, np.abs(d) > (F(0.5) - EPS)


def test_int8_conv_fast_path_equals_reference_outside_the_guard_band():
    rng = np.random.default_rng(11)
    n = 2_000_000
    s_in = F(rng.uniform(0.005, 0.08))
    s_out = F(rng.uniform(0.01, 0.2))
    s_w = rng.uniform(1e-4, 0.02, n).astype(np.float32)
    M = (np.float64(s_in) * s_w.astype(np.float64) / np.float64(s_out)).astype(np.float32)
    # accumulators whose scaled value covers the clamp range several times over
    t_target = rng.uniform(-300, 300, n)
    acc = np.rint(t_target / M.astype(np.float64)).astype(np.int64)
    bias = np.rint(rng.uniform(-50, 50, n) / M.astype(np.float64)).astype(np.int64)
    x = (acc + bias).astype(np.float32)  # exact: |acc + bias| < 2^24 by construction below
    keep = np.abs(acc + bias) < 2 ** 24
    for recipe in ("hcl", "ref"):
        for act in (-1, 0, 6):
            if recipe == "hcl":
                f = ((x * s_in).astype(np.float32) * s_w).astype(np.float32)
            else:
                f = (x * (s_in * s_w).astype(np.float32)).astype(np.float32)
            if act == 0:
                f = np.maximum(f, F(0))
            if act == 6:
                f = np.minimum(np.maximum(f, F(0)), F(6))
            q_ref = np.clip(round_half_away((f / s_out).astype(np.float32)), -127, 127).astype(np.int64)
            # fast path, FUSE form: t = fma((float)acc, M, fl(bias*M)); integer clamp after rounding
            bm = (bias.astype(np.float64) * M.astype(np.float64)).astype(np.float32)
            t = (acc.astype(np.float32).astype(np.float64) * M.astype(np.float64) + bm.astype(np.float64)).astype(np.float32)
            q, guarded = guard_and_round(t)
            lo = -127 if act < 0 else 0
            hi = 127 if act != 6 else min(127, int(round_half_away(np.array([F(6) / s_out], dtype=np.float32))[0]))
            q = np.clip(q, lo, hi)
            ok = keep & ~guarded & (np.abs(bm) <= 100)
            assert ok.sum() > n // 2
            assert np.array_equal(q[ok], q_ref[ok]), (recipe, act)
            # and the guard band is as thin as the design says (its elements take the slow path)
            assert guarded[keep].mean() < 1e-3


def test_uint8_and_pointwise_reciprocal_instead_of_division():
    rng = np.random.default_rng(12)
    n = 2_000_000
    s_out = F(rng.uniform(0.01, 0.2))
    r_out = (F(1.0) / s_out).astype(np.float32)
    f = rng.uniform(-300, 300, n).astype(np.float32) * s_out  # any float the reference's chain may have produced
    q_ref = round_half_away((f / s_out).astype(np.float32)).astype(np.int64)
    t = (f * r_out).astype(np.float32)
    q, guarded = guard_and_round(t)
    assert np.array_equal(q[~guarded], q_ref[~guarded])
    assert guarded.mean() < 1e-3


def test_same_scale_relu_and_maxpool_are_byte_operations():
    """relu_same_scale_kernel / pool_max_same_scale_kernel: with equal input and output quantisation the reference's
    dequantise -> op -> requantise is max(byte, zero point) resp. the byte-wise max (relu_kernel_ref_uint8.c:85,
    pooling_kernel_ref_uint8.c:131-197, relu_kernel_ref_int8.c, pooling_kernel_ref_int8.c)."""
    rng = np.random.default_rng(13)
    for _ in range(200):
        s = F(rng.uniform(1e-3, 0.5))
        z = int(rng.integers(0, 256))
        b = np.arange(256, dtype=np.int64)
        # uint8 relu: round(f / s + z) with f = max((b - z) * s, 0)
        f = np.maximum(((b - z).astype(np.float32) * s).astype(np.float32), F(0))
        q = np.clip(round_half_away(((f / s).astype(np.float32) + F(z)).astype(np.float32)), 0, 255).astype(np.int64)
        assert np.array_equal(q, np.maximum(b, z))
        # uint8 max pool over a window: max of dequantised values, round(v / s) + z, upper clamp only
        win = rng.integers(0, 256, (1000, 4))
        v = ((win - z).astype(np.float32) * s).astype(np.float32).max(axis=1)
        q = np.minimum(round_half_away((v / s).astype(np.float32)).astype(np.int64) + z, 255)
        assert np.array_equal(q, win.max(axis=1))
        # int8: relu -> max(b, 0); max pool -> round(imax * (s / s)) clamped to +-127
        bi = np.arange(-128, 128, dtype=np.int64)
        f = np.maximum((bi.astype(np.float32) * s).astype(np.float32), F(0))
        q = np.clip(round_half_away((f / s).astype(np.float32)), -127, 127).astype(np.int64)
        assert np.array_equal(q, np.maximum(bi, 0))
