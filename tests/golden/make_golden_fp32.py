#!/usr/bin/env python3
"""Generate tests/golden/ref_fp32_conv.npz: fp32 convolutions run by the UNMODIFIED reference CPU device (oracle/_ref) --
Winograd F(4,3) where its winograd_support() selects it (conv_kernel_x86.c:1896-1915) and the fp32 depthwise kernels
(conv_dw_kernel_x86.c, batch 1: conv_dw_hcl_x86.c:536) -- for the device's fp32 members of the path (tb200k_conv_winograd43_f32,
tb200k_conv_dw3x3_f32).  Run where /root/reference exists:  python tests/golden/make_golden_fp32.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.pyoracle import Reference, conv_f32  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
# name, n, c, h, w, oc, stride, pad, group, activation
CASES = [
    ("wino_a", 2, 32, 20, 22, 48, 1, 1, 1, 0),     # ragged 4x4 tiling (20 x 22), ReLU
    ("wino_b", 1, 64, 56, 56, 64, 1, 1, 1, -1),    # ResNet-50 stage-2 3x3 (C3's fp32 twin)
    ("wino_c", 3, 16, 13, 17, 32, 1, 1, 1, 6),     # odd sizes, ReLU6
    ("wino_d", 1, 24, 14, 14, 16, 1, 0, 1, -1),    # no padding
    ("dw_s1", 1, 32, 20, 22, 32, 1, 1, 32, 0),
    ("dw_s2", 1, 48, 21, 19, 48, 2, 1, 48, 6),
    ("dw_s1_nob", 1, 16, 12, 12, 16, 1, 1, 16, -1),
]


def main():
    ref = Reference()
    rng = np.random.default_rng(2024)
    d = {}
    for name, n, c, h, w, oc, s, p, g, act in CASES:
        x = rng.standard_normal((n, c, h, w)).astype(np.float32)
        wt = (rng.standard_normal((oc, c // g, 3, 3)) * (0.3 if g > 1 else 0.08)).astype(np.float32)
        b = None if name.endswith("nob") else rng.standard_normal(oc).astype(np.float32)
        y = conv_f32(ref, x, wt, b, s, p, g, act)
        d[name + "_x"], d[name + "_w"], d[name + "_y"] = x, wt, y
        if b is not None:
            d[name + "_b"] = b
        d[name + "_p"] = np.array([s, p, g, act], np.int32)
        print(name, y.shape, float(np.abs(y).max()))
    np.savez_compressed(os.path.join(OUT, "ref_fp32_conv.npz"), **d)


if __name__ == "__main__":
    main()
