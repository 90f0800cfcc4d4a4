"""The fp32 members of the path (north star: "fp32 paths within 1e-4 rel"): Winograd F(4x4,3x3) and
depthwise 3x3 through the kernel-level ABI, against bytes produced by the UNMODIFIED reference (tests/golden/ref_fp32_conv.npz,
generator tests/golden/make_golden_fp32.py) and, for shapes the fixture does not hold, a torch fp64 reference of the same op.
Tolerance, as the north star states it: max |device - reference| <= 1e-4 * max |reference|."""
import ctypes as C
import os

import numpy as np
import pytest

from tengine_b200 import abi

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_fp32_conv.npz")
TOL = 1e-4


def _run(ctx, x, w, b, stride, pad, group, act):
    import torch
    from tengine_b200 import runtime as rt

    lib = rt.lib()
    n, c, h, wd = x.shape
    oc = w.shape[0]
    oh, ow = (h + 2 * pad - 3) // stride + 1, (wd + 2 * pad - 3) // stride + 1
    s = abi.KConvShape()
    s.n, s.h, s.w, s.c, s.oh, s.ow, s.oc = n, h, wd, c, oh, ow, oc
    s.kh = s.kw = 3
    s.sh = s.sw = stride
    s.ph0 = s.pw0 = pad
    s.dh = s.dw = 1
    s.group = group
    dx, dw = torch.from_numpy(x).cuda(), torch.from_numpy(w).cuda()
    db = torch.from_numpy(b).cuda() if b is not None else None
    out = torch.empty((n, oc, oh, ow), dtype=torch.float32, device="cuda")
    bp = C.c_void_p(db.data_ptr()) if db is not None else None
    torch.cuda.synchronize()
    if group == 1:
        lib.tb200k_conv_winograd43_f32_workspace.restype = C.c_size_t
        ws = torch.empty(lib.tb200k_conv_winograd43_f32_workspace(C.byref(s)), dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize()
        rc = lib.tb200k_conv_winograd43_f32(C.c_void_p(dx.data_ptr()), C.c_void_p(dw.data_ptr()), bp, C.c_void_p(out.data_ptr()), C.byref(s), int(act),
                                            C.c_void_p(ws.data_ptr()), C.c_void_p(ctx.stream))
    else:
        rc = lib.tb200k_conv_dw3x3_f32(C.c_void_p(dx.data_ptr()), C.c_void_p(dw.data_ptr()), bp, C.c_void_p(out.data_ptr()), C.byref(s), int(act),
                                       C.c_void_p(ctx.stream))
    assert rc == 0, lib.tb200_last_error().decode()
    torch.cuda.synchronize()
    return out.cpu().numpy()


def _cases():
    d = np.load(GOLD)
    return sorted({k[:-2] for k in d.files if k.endswith("_x")})


@pytest.mark.parametrize("name", _cases())
def test_fp32_conv_vs_reference_fixture(ctx, name):
    d = np.load(GOLD)
    stride, pad, group, act = (int(v) for v in d[name + "_p"])
    b = d[name + "_b"] if name + "_b" in d.files else None
    got = _run(ctx, d[name + "_x"], d[name + "_w"], b, stride, pad, group, act)
    ref = d[name + "_y"]
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() <= TOL * np.abs(ref).max(), float(np.abs(got - ref).max() / np.abs(ref).max())


@pytest.mark.parametrize("shape", [(2, 128, 28, 28, 128, 1), (1, 256, 14, 14, 256, 1), (4, 64, 56, 56, 64, 64), (2, 512, 14, 14, 512, 512)],
                         ids=["resnet_3x3_28", "resnet_3x3_14", "dw64_56", "dw512_14"])
def test_fp32_conv_vs_torch_fp64(ctx, shape):
    """ResNet-50's fp32 3x3 layers (the shapes winograd_support() admits) and MobileNet-size depthwise layers."""
    import torch

    n, c, h, w, oc, group = shape
    rng = np.random.default_rng(sum(shape))
    x = rng.standard_normal((n, c, h, w)).astype(np.float32)
    wt = (rng.standard_normal((oc, c // group, 3, 3)) * (0.3 if group > 1 else 0.05)).astype(np.float32)
    b = rng.standard_normal(oc).astype(np.float32)
    for stride in ((1, 2) if group > 1 else (1,)):
        got = _run(ctx, x, wt, b, stride, 1, group, 0)
        ref = torch.nn.functional.conv2d(torch.from_numpy(x).double(), torch.from_numpy(wt).double(), torch.from_numpy(b).double(), stride=stride,
                                         padding=1, groups=group).relu().numpy()
        assert np.abs(got - ref).max() <= TOL * np.abs(ref).max(), (stride, float(np.abs(got - ref).max() / np.abs(ref).max()))
