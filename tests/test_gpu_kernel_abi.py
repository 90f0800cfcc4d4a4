"""The per-kernel entry points of the C ABI (tb200k_*, include/tengine_b200.h:196-237) on DEVICE pointers, each against the
CPU oracle on the same seeded layer.  These launchers use the literal reference arithmetic (epi_from_abi: fast_ok = 0), so
they are also an independent second construction of the epilogue constants.  torch only provides the device memory."""
import ctypes as C

import numpy as np
import pytest

from tengine_b200 import abi
from tengine_b200.graphdef import GraphDef

pytestmark = pytest.mark.gpu


def _cpad(c):
    return (c + 15) // 16 * 16


def _dev(a):
    import torch

    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _nhwc(x, cp):
    n, c, h, w = x.shape
    out = np.zeros((n, h, w, cp), x.dtype)
    out[..., :c] = x.transpose(0, 2, 3, 1)
    return out


def _epilogue(g, L, keep):
    u8 = g.data_type == abi.DT_UINT8
    tin, tout = g.tensors[L["inputs"][0]], g.tensors[L["output"]]
    oc = tout["dims"][1]
    ocp = _cpad(oc)
    bias = np.zeros(ocp, np.int32)
    if L["bias"] is not None:
        bias[:oc] = L["bias"]
    ws = np.ones(ocp, np.float32)
    ws[:oc] = L["weight_scales"][0] if u8 else L["weight_scales"]
    if u8:
        ws[:] = L["weight_scales"][0]
    db, dw = _dev(bias), _dev(ws)
    keep += [db, dw]
    e = abi.KEpilogue()
    e.bias, e.w_scale = db.data_ptr(), dw.data_ptr()
    e.in_scale, e.out_scale = tin["scale"], tout["scale"]
    e.in_zero, e.w_zero, e.out_zero = tin["zero_point"], L["weight_zero"], tout["zero_point"]
    e.activation, e.recipe, e.is_uint8 = L["activation"], L["recipe"], int(u8)
    e.fc_rounding = int(L["op"] == abi.OP_FC)
    e.w_scale_tensor = float(L["weight_scales"][0]) if u8 else 0.0
    return e


def _shape(g, L):
    n, c, h, w = g.dims(L["inputs"][0])
    _, oc, oh, ow = g.dims(L["output"])
    s = abi.KConvShape()
    s.n, s.h, s.w, s.c, s.oh, s.ow, s.oc = n, h, w, c, oh, ow, oc
    s.kh, s.kw, s.sh, s.sw = L["kernel_h"], L["kernel_w"], L["stride_h"], L["stride_w"]
    s.ph0, s.pw0, s.dh, s.dw, s.group = L["pad_h0"], L["pad_w0"], L["dilation_h"], L["dilation_w"], L["group"]
    return s


def _conv_graph(rng, dtype, n, c, h, w, oc, k, s, p, group, act, recipe, dilation=1):
    g = GraphDef(dtype)
    u8 = dtype == abi.DT_UINT8
    x = g.input(n, c, h, w, 0.02, 131 if u8 else 0)
    kk = (c // group) * k * k
    if u8:
        wq, ws, wz = rng.integers(0, 256, (oc, c // group, k, k)).astype(np.uint8), [0.004], 117
        so = 0.02 * 0.004 * np.sqrt(kk) * 74 * 74 / 100
    else:
        wq, ws, wz = rng.integers(-127, 128, (oc, c // group, k, k)).astype(np.int8), rng.uniform(0.001, 0.01, oc), 0
        so = 0.02 * 0.0055 * np.sqrt(kk) * 73 * 73 / 100
    b = rng.integers(-2000, 2000, oc).astype(np.int32)
    y = g.conv(x, wq, b, ws, so, 110 if u8 else 0, stride=s, pad=p, dilation=dilation, group=group, activation=act, recipe=recipe,
               weight_zero=wz)
    g.mark_output(y)
    xin = rng.integers(0, 256, (n, c, h, w)).astype(np.uint8) if u8 else rng.integers(-127, 128, (n, c, h, w)).astype(np.int8)
    return g, xin


def _check(rc):
    from tengine_b200 import runtime as rt

    assert rc == 0, rt.lib().tb200_last_error().decode()


def _finish(out_dev, g, want):
    import torch

    torch.cuda.synchronize()
    n, oc, oh, ow = g.dims(g.outputs[0])
    got = out_dev.cpu().numpy().reshape(n, oh, ow, -1)
    assert np.array_equal(got[..., :oc].transpose(0, 3, 1, 2), want)
    assert not got[..., oc:].any(), "pad lanes must hold 0"


@pytest.mark.parametrize("dtype", [abi.DT_INT8, abi.DT_UINT8], ids=["int8", "uint8"])
@pytest.mark.parametrize("case", [(2, 20, 9, 11, 24, 3, 1, 1, 1, 1), (1, 16, 12, 12, 16, 3, 2, 2, 1, 2), (2, 8, 9, 9, 16, 3, 1, 1, 2, 1),
                                  (1, 32, 13, 13, 40, 3, 1, 3, 1, 3)],
                         ids=["3x3", "3x3_s2_dil2", "grouped", "dil3"])
def test_tb200k_conv_direct(ctx, oracle, dtype, case):
    import torch
    from tengine_b200 import runtime as rt

    n, c, h, w, oc, k, s, p, group, dil = case
    rng = np.random.default_rng(sum(case) + dtype)
    g, x = _conv_graph(rng, dtype, n, c, h, w, oc, k, s, p, group, 0, abi.RECIPE_REF, dilation=dil)
    L = g.layers[0]
    want = oracle.run(g, [x], uint8_mode=0)[g.outputs[0]]
    cp, ocp, cg = _cpad(c), _cpad(oc), c // group
    cgp = cp if group == 1 else cg
    wp = np.zeros((ocp, k, k, cgp), g.np_dtype)
    wp[:oc, :, :, :cg] = L["weight"].transpose(0, 2, 3, 1)
    keep = []
    e = _epilogue(g, L, keep)
    sh = _shape(g, L)
    din, dwt = _dev(_nhwc(x, cp)), _dev(wp)
    out = torch.zeros(n * sh.oh * sh.ow * ocp, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()  # torch's allocations / copies ran on its own stream
    _check(rt.lib().tb200k_conv_direct(C.c_void_p(din.data_ptr()), C.c_void_p(dwt.data_ptr()), C.c_void_p(out.data_ptr()), C.byref(sh),
                                       C.byref(e), C.c_void_p(ctx.stream)))
    _finish(out.view(torch.int8 if dtype == abi.DT_INT8 else torch.uint8), g, want)


@pytest.mark.parametrize("dtype", [abi.DT_INT8, abi.DT_UINT8], ids=["int8", "uint8"])
@pytest.mark.parametrize("stride", [1, 2])
def test_tb200k_conv_dw3x3(ctx, oracle, dtype, stride):
    import torch
    from tengine_b200 import runtime as rt

    n, c, h, w = 2, 40, 13, 15
    rng = np.random.default_rng(90 + stride + dtype)
    g, x = _conv_graph(rng, dtype, n, c, h, w, c, 3, stride, 1, c, 6, abi.RECIPE_REF)
    L = g.layers[0]
    want = oracle.run(g, [x], uint8_mode=0)[g.outputs[0]]
    cp = _cpad(c)
    wp = np.zeros((3, 3, cp), g.np_dtype)
    wp[:, :, :c] = L["weight"].reshape(c, 3, 3).transpose(1, 2, 0)
    keep = []
    e = _epilogue(g, L, keep)
    sh = _shape(g, L)
    din, dwt = _dev(_nhwc(x, cp)), _dev(wp)
    out = torch.zeros(n * sh.oh * sh.ow * cp, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()  # torch's allocations / copies ran on its own stream
    _check(rt.lib().tb200k_conv_dw3x3(C.c_void_p(din.data_ptr()), C.c_void_p(dwt.data_ptr()), C.c_void_p(out.data_ptr()), C.byref(sh),
                                      C.byref(e), C.c_void_p(ctx.stream)))
    _finish(out.view(torch.int8 if dtype == abi.DT_INT8 else torch.uint8), g, want)


@pytest.mark.parametrize("dtype", [abi.DT_INT8, abi.DT_UINT8], ids=["int8", "uint8"])
def test_tb200k_conv_stem_nchw(ctx, oracle, dtype):
    import torch
    from tengine_b200 import runtime as rt

    n, c, h, w, oc, k = 2, 3, 30, 34, 24, 3
    rng = np.random.default_rng(5 + dtype)
    g, x = _conv_graph(rng, dtype, n, c, h, w, oc, k, 2, 1, 1, 0, abi.RECIPE_HCL)
    L = g.layers[0]
    want = oracle.run(g, [x], uint8_mode=0)[g.outputs[0]]
    ocp = _cpad(oc)
    wp = np.zeros((ocp, k, k, 4), g.np_dtype)
    wp[:oc, :, :, :c] = L["weight"].transpose(0, 2, 3, 1)
    keep = []
    e = _epilogue(g, L, keep)
    sh = _shape(g, L)
    din, dwt = _dev(x), _dev(wp)
    out = torch.zeros(n * sh.oh * sh.ow * ocp, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()  # torch's allocations / copies ran on its own stream
    _check(rt.lib().tb200k_conv_stem_nchw(C.c_void_p(din.data_ptr()), C.c_void_p(dwt.data_ptr()), C.c_void_p(out.data_ptr()), C.byref(sh),
                                          C.byref(e), C.c_void_p(ctx.stream)))
    _finish(out.view(torch.int8 if dtype == abi.DT_INT8 else torch.uint8), g, want)


@pytest.mark.parametrize("case", [(2, 48, 9, 9, 40), (1, 256, 14, 14, 512), (3, 32, 20, 20, 64)], ids=lambda c: "n%d_k%d_%dx%d_oc%d" % c)
def test_tb200k_gemm_i8(ctx, oracle, case):
    import torch
    from tengine_b200 import runtime as rt

    n, c, h, w, oc = case
    rng = np.random.default_rng(sum(case))
    g, x = _conv_graph(rng, abi.DT_INT8, n, c, h, w, oc, 1, 1, 0, 1, 0, abi.RECIPE_HCL)
    L = g.layers[0]
    want = oracle.run(g, [x])[g.outputs[0]]
    cp, ocp = _cpad(c), _cpad(oc)
    wp = np.zeros((ocp, cp), np.int8)
    wp[:oc, :c] = L["weight"].reshape(oc, c)
    keep = []
    e = _epilogue(g, L, keep)
    din, dwt = _dev(_nhwc(x, cp)), _dev(wp)
    out = torch.zeros(n * h * w * ocp, dtype=torch.int8, device="cuda")
    lib = rt.lib()
    lib.tb200k_gemm_i8.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
    torch.cuda.synchronize()
    _check(lib.tb200k_gemm_i8(din.data_ptr(), dwt.data_ptr(), out.data_ptr(), n * h * w, cp, oc, C.addressof(e), ctx.stream))
    _finish(out, g, want)


@pytest.mark.parametrize("shape", [(2, 3, 17, 19), (1, 40, 7, 9), (3, 16, 8, 8)], ids=str)
def test_tb200k_layout_round_trip(ctx, shape):
    import torch
    from tengine_b200 import runtime as rt

    n, c, h, w = shape
    x = np.random.default_rng(1).integers(-128, 128, shape).astype(np.int8)
    cp = _cpad(c)
    din = _dev(x)
    mid = torch.full((n * h * w * cp,), 77, dtype=torch.int8, device="cuda")
    back = torch.zeros(n * c * h * w, dtype=torch.int8, device="cuda")
    lib = rt.lib()
    torch.cuda.synchronize()
    _check(lib.tb200k_nchw_to_nhwc(C.c_void_p(din.data_ptr()), C.c_void_p(mid.data_ptr()), n, c, h, w, C.c_void_p(ctx.stream)))
    _check(lib.tb200k_nhwc_to_nchw(C.c_void_p(mid.data_ptr()), C.c_void_p(back.data_ptr()), n, c, h, w, C.c_void_p(ctx.stream)))
    torch.cuda.synchronize()
    assert np.array_equal(mid.cpu().numpy().reshape(n, h, w, cp), _nhwc(x, cp))  # pad lanes are written as 0
    assert np.array_equal(back.cpu().numpy().reshape(shape), x)
    assert lib.tb200k_cpad(c) == cp
