"""CPU check of the CLAIMS the uint8 tensor-core path rests on (tengine_b200/csrc/engine.cu packing,
gemm_tcgen05.cu epilogue_unit_u8), restated in numpy:

 (1) algebra: with the TMA unit zero-filling taps outside the image,
        sum_{taps inside} (x - zx)(w - zw)  =  sum x*(w - 128)  +  (128 - zw) * sum(x)  +  y_int  +  border[pattern]
     where sum x*(w-128) is what the UMMA accumulates (unsigned A, B = w - 128 as int8), sum(x) is what the row-sum warps add up,
     y_int = -zx*sum_all(w) + K*zx*zw is the per-channel constant of an interior pixel and border[pattern] gives back
     zx*(sum_c w[t] - C*zw) for every tap t the pixel's window loses to the padding, looked up by the pattern index
     ((a*(KH+1) + b)*(pw+1) + c)*(KW+1) + d  (a / b rows cut at the top / bottom, c / d columns cut left / right);
 (2) numerics: the int8-form fast epilogue t = fl((float)(acc + bias) * M), M = fl(s_in*s_w/s_out), rounds to the reference's
     q = round(fl(fl(fl(acc*S) + fl(bias*S)) [act] / s_out)) + z_out whenever t is farther than 2^-13 from a half-integer, provided
     |bias*M| <= 250 (the per-layer check `u8_tc_fast`).
No GPU, no oracle library: this pins the mathematics; the GPU parity tests pin the kernels."""
import numpy as np
import pytest

F = np.float32
MAGIC = F(12582912.0)
EPS = F(2.0 ** -13)


def reference_integer_conv(x, w, zx, zw, stride, pad):
    """sum over the taps INSIDE the image of (x - zx)(w - zw): what the reference's uint8 convolution accumulates (padding
    contributes 0.0f there, conv_kernel_x86.c:167-181), in exact integers.  x [C,H,W], w [OC,C,KH,KW] -> [OC,OH,OW]"""
    C, H, W = x.shape
    OC, _, KH, KW = w.shape
    OH, OW = (H + 2 * pad - KH) // stride + 1, (W + 2 * pad - KW) // stride + 1
    out = np.zeros((OC, OH, OW), np.int64)
    xs, ws = x.astype(np.int64) - zx, w.astype(np.int64) - zw
    for oh in range(OH):
        for ow in range(OW):
            for kh in range(KH):
                for kw in range(KW):
                    ih, iw = oh * stride - pad + kh, ow * stride - pad + kw
                    if 0 <= ih < H and 0 <= iw < W:
                        out[:, oh, ow] += ws[:, :, kh, kw] @ xs[:, ih, iw]
    return out


@pytest.mark.parametrize("k,stride,pad", [(3, 1, 1), (3, 2, 1), (5, 1, 2), (1, 1, 0), (7, 2, 3)])
def test_tensor_core_decomposition_with_border_pattern_table(k, stride, pad):
    rng = np.random.default_rng(100 + k * 10 + stride)
    C, H, W, OC = 5, 9, 8, 6
    zx, zw = int(rng.integers(1, 255)), int(rng.integers(1, 255))
    x = rng.integers(0, 256, (C, H, W)).astype(np.int64)
    w = rng.integers(0, 256, (OC, C, k, k)).astype(np.int64)
    want = reference_integer_conv(x, w, zx, zw, stride, pad)
    KH = KW = k
    OH, OW = want.shape[1:]
    # what the device forms -------------------------------------------------------------------------------------------------------
    xpad = np.zeros((C, H + 2 * pad, W + 2 * pad), np.int64)  # the TMA unit zero-fills outside the image
    xpad[:, pad:pad + H, pad:pad + W] = x
    b_signed = w - 128  # packed B operand (int8 range)
    assert b_signed.min() >= -128 and b_signed.max() <= 127
    y_int = -zx * w.sum(axis=(1, 2, 3)) + C * KH * KW * zx * zw  # interior constant (bias left out: it just adds)
    tapc = zx * (w.sum(axis=1) - C * zw)  # [OC, KH, KW]: what a tap that falls into the padding must give back
    table = {}
    for a in range(pad + 1):
        for b in range(KH + 1):
            for c in range(pad + 1):
                for d in range(KW + 1):
                    pid = ((a * (KH + 1) + b) * (pad + 1) + c) * (KW + 1) + d
                    miss = np.zeros((KH, KW), bool)
                    for kh in range(KH):
                        for kw in range(KW):
                            miss[kh, kw] = kh < a or kh >= KH - b or kw < c or kw >= KW - d
                    table[pid] = (tapc * miss).sum(axis=(1, 2))
    got = np.zeros_like(want)
    for oh in range(OH):
        for ow in range(OW):
            win = xpad[:, oh * stride:oh * stride + KH, ow * stride:ow * stride + KW]
            acc = np.einsum("ockl,ckl->o", b_signed, win)  # UMMA: unsigned x signed
            sx = win.sum()  # row-sum warps: all bytes of the pixel's A rows (zeros where the tap is outside)
            ih0, iw0 = oh * stride - pad, ow * stride - pad
            a = min(max(-ih0, 0), pad)
            b = min(max(ih0 + KH - H, 0), KH)
            c = min(max(-iw0, 0), pad)
            d = min(max(iw0 + KW - W, 0), KW)
            pid = ((a * (KH + 1) + b) * (pad + 1) + c) * (KW + 1) + d
            got[:, oh, ow] = acc + (128 - zw) * sx + y_int + table[pid]
    assert np.array_equal(got, want)


def round_half_away(x):
    return np.where(x >= 0, np.floor(x + F(0.5)), np.ceil(x - F(0.5))).astype(np.float32)


@pytest.mark.parametrize("act", [-1, 0, 6])
def test_int8_form_epilogue_of_the_uint8_path_is_exact_outside_the_guard_band(act):
    rng = np.random.default_rng(7 + act)
    n = 2_000_000
    s_in, s_w, s_out = F(rng.uniform(0.005, 0.08)), F(rng.uniform(1e-4, 0.02)), F(rng.uniform(0.01, 0.2))
    z_out = int(rng.integers(0, 200))
    S = (s_in * s_w).astype(np.float32)  # in_w_scale of the HCL recipe (conv_kernel_x86.c:1723)
    M = F(np.float64(s_in) * np.float64(s_w) / np.float64(s_out))
    t_target = rng.uniform(-40.0, 300.0, n)
    bias = np.rint(rng.uniform(-250, 250, n) / np.float64(M)).astype(np.int64)  # |bias*M| <= 250: the bound u8_tc_fast checks
    acc = np.rint(t_target / np.float64(M)).astype(np.int64) - bias
    keep = (np.abs(acc) < 2 ** 24) & (np.abs(acc + bias) < 2 ** 24) & (np.abs(bias) < 2 ** 24)
    # reference (uint8 HCL epilogue): f = fl(acc*S) + fl(bias*S) ; act ; q = round(f / s_out) + z_out ; clamp 0..255
    f = ((acc.astype(np.float32) * S).astype(np.float32) + (bias.astype(np.float32) * S).astype(np.float32)).astype(np.float32)
    if act == 0:
        f = np.maximum(f, F(0))
    if act == 6:
        f = np.minimum(np.maximum(f, F(0)), F(6))
    q_ref = np.clip(round_half_away((f / s_out).astype(np.float32)).astype(np.int64) + z_out, 0, 255)
    # device: t = fl((float)(acc + bias) * M), round to nearest with the magic constant, integer clamp afterwards
    t = ((acc + bias).astype(np.float32) * M).astype(np.float32)
    r = (t + MAGIC).astype(np.float32)
    d = (t - (r - MAGIC).astype(np.float32)).astype(np.float32)
    q = (r.view(np.int32).astype(np.int64) - int(MAGIC.view(np.int32)))
    lo = 0 if act >= 0 else -10 ** 9
    hi = 10 ** 9 if act != 6 else int(round_half_away(np.array([F(6) / s_out], np.float32))[0])
    q = np.clip(np.clip(q, lo, hi) + z_out, 0, 255)
    guarded = np.abs(d) > (F(0.5) - EPS)
    ok = keep & ~guarded
    assert ok.sum() > n // 2
    assert np.array_equal(q[ok], q_ref[ok])
    assert guarded[keep].mean() < 1e-3  # the band is as thin as the design says
