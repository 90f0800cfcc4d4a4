"""CPU check of the fp32 fixtures (tests/golden/ref_fp32_conv.npz, produced by the unmodified reference): each equals a torch fp64
convolution of the same operands within the north star's fp32 tolerance (1e-4 relative to the tensor's magnitude), so the GPU test
that compares the device with these bytes is a meaningful check of both."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_fp32_conv.npz")


def test_reference_fp32_fixtures_agree_with_fp64():
    import torch

    d = np.load(GOLD)
    names = sorted({k[:-2] for k in d.files if k.endswith("_x")})
    assert len(names) >= 6
    for name in names:
        stride, pad, group, act = (int(v) for v in d[name + "_p"])
        b = torch.from_numpy(d[name + "_b"]).double() if name + "_b" in d.files else None
        y = torch.nn.functional.conv2d(torch.from_numpy(d[name + "_x"]).double(), torch.from_numpy(d[name + "_w"]).double(), b, stride=stride, padding=pad,
                                       groups=group)
        if act == 0:
            y = y.relu()
        elif act > 0:
            y = y.clamp(0, act)
        ref = d[name + "_y"]
        assert np.abs(y.numpy() - ref).max() <= 1e-4 * np.abs(ref).max(), name


def test_reference_live_fp32_winograd_matches_fixture(reference):
    """Where oracle/_ref is built: regenerate one case live and compare bit for bit (the fixture really is the reference's output)."""
    from oracle.pyoracle import conv_f32

    d = np.load(GOLD)
    stride, pad, group, act = (int(v) for v in d["wino_a_p"])
    y = conv_f32(reference, d["wino_a_x"], d["wino_a_w"], d["wino_a_b"], stride, pad, group, act)
    assert np.abs(y - d["wino_a_y"]).max() <= 1e-6 * np.abs(y).max()  # thread-count dependent summation order only
