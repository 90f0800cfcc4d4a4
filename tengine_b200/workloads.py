"""Synthetic quantised workloads of BASELINE.json: MobileNet-v1, ResNet-50, YOLOv3-tiny as GraphDef tables.

Layer shapes follow the reference's own benchmark/models/{mobilenet,resnet50,yolov3_tiny}_benchmark.tmfile
(SURVEY.md 8(a) layer tables).  Weights are seeded random (there is no network for checkpoints); they are
quantised the way the reference's tools/quantize/quant_save_graph.cpp does it:
  int8  : symmetric per-output-channel weight scales max|w|/127 (:519-555), activations per tensor max|a|/127
  uint8 : asymmetric per-tensor (scale, zero_point) for weights and activations (:246-263)
  bias  : int32 with scale s_in*s_w[oc] (:579-600, :283-298)
Activation scales come from one fp32 calibration pass (torch CPU) over a small seeded batch, so that the
quantised activations neither saturate nor collapse.
Test/bench infrastructure; not part of the device library.
"""
import numpy as np

from . import abi
from .graphdef import GraphDef


class _H:  # handle: tensor id + fp32 calibration activation
    def __init__(self, tid, act):
        self.tid, self.act = tid, act


class QuantBuilder:
    def __init__(self, data_type, batch, c, h, w, seed=1234, calib_batch=2, inplace=False):
        """inplace: give max-pooling and ReLU (slope 0) outputs the quantisation of their input, as the reference's
        quantisation tool does by default (tools/quantize/quant_tool_int8.cpp:67 `inplace = true`,
        quant_save_graph.cpp:136-200 passes the scale through Pooling(max) / ReLU / Flatten / Reshape / Clip)."""
        import torch

        self.inplace = inplace

        self.torch = torch
        self.g = GraphDef(data_type)
        self.u8 = data_type == abi.DT_UINT8
        self.rng = np.random.default_rng(seed)
        self.batch = batch
        # network input: int8 in [-127,127] with scale 1/127; uint8 in [0,255] with scale 2/255, zp 128
        if self.u8:
            self.in_scale, self.in_zero = np.float32(2.0 / 255.0), 128
            q = self.rng.integers(0, 256, (calib_batch, c, h, w))
        else:
            self.in_scale, self.in_zero = np.float32(1.0 / 127.0), 0
            q = self.rng.integers(-127, 128, (calib_batch, c, h, w))
        act = torch.from_numpy(((q - self.in_zero) * float(self.in_scale)).astype(np.float32))
        self.input = _H(self.g.input(batch, c, h, w, self.in_scale, self.in_zero), act)

    # ---- quantisation helpers ----
    def _act_q(self, a):
        a = a.detach().numpy()
        if self.u8:
            lo, hi = min(float(a.min()), 0.0), max(float(a.max()), 0.0)
            scale = np.float32(max(hi - lo, 1e-6) / 255.0)
            zp = int(np.clip(np.round(-lo / scale), 0, 255))
            return scale, zp
        return np.float32(max(float(np.abs(a).max()), 1e-6) / 127.0), 0

    def _weight_q(self, w):
        oc = w.shape[0]
        if self.u8:
            lo, hi = min(float(w.min()), 0.0), max(float(w.max()), 0.0)
            scale = np.float32(max(hi - lo, 1e-8) / 255.0)
            zp = int(np.clip(np.round(-lo / scale), 0, 255))
            q = np.clip(np.round(w / scale + zp), 0, 255).astype(np.uint8)
            return q, np.array([scale], np.float32), zp
        s = (np.abs(w.reshape(oc, -1)).max(axis=1) / 127.0).astype(np.float32)
        s = np.maximum(s, np.float32(1e-8))
        q = np.clip(np.round(w / s.reshape((-1,) + (1,) * (w.ndim - 1))), -127, 127).astype(np.int8)
        return q, s, 0

    def _bias_q(self, b, s_in, ws):
        return np.round(b / (np.float32(s_in) * ws.astype(np.float32))).astype(np.int64).clip(-2**31, 2**31 - 1).astype(np.int32)

    def _fp_weight(self, oc, cg, kh, kw):
        fan_in = cg * kh * kw
        return (self.rng.standard_normal((oc, cg, kh, kw)) * np.sqrt(2.0 / fan_in)).astype(np.float32)

    # ---- ops ----
    def conv(self, x, oc, k, stride=1, pad=0, group=1, activation=-1, bias=True, recipe=abi.RECIPE_HCL, dilation=1):
        F = self.torch.nn.functional
        c = self.g.dims(x.tid)[1]
        w = self._fp_weight(oc, c // group, k, k)
        b = (self.rng.standard_normal(oc) * 0.05).astype(np.float32) if bias else None
        wq, ws, wz = self._weight_q(w)
        s_in = self.g.tensors[x.tid]["scale"]
        wdq = (wq.astype(np.float32) - wz) * ws.reshape((-1, 1, 1, 1) if not self.u8 else (1, 1, 1, 1))
        bq = self._bias_q(b, s_in, ws if not self.u8 else ws[0] * np.ones(oc, np.float32)) if bias else None
        bdq = None if bq is None else (bq.astype(np.float32) * np.float32(s_in) * (ws if not self.u8 else ws[0]))
        a = F.conv2d(x.act, self.torch.from_numpy(wdq), None if bdq is None else self.torch.from_numpy(bdq.astype(np.float32)),
                     stride=stride, padding=pad, dilation=dilation, groups=group)
        if activation == 0:
            a = F.relu(a)
        elif activation == 6:
            a = a.clamp(0, 6)
        so, zo = self._act_q(a)
        tid = self.g.conv(x.tid, wq, bq, ws, so, zo, stride=stride, pad=pad, dilation=dilation, group=group,
                          activation=activation, recipe=recipe, weight_zero=wz)
        return _H(tid, a)

    def fc(self, x, oc, bias=True):
        n, c, h, w = self.g.dims(x.tid)
        k = c * h * w
        wf = (self.rng.standard_normal((oc, k)) * np.sqrt(1.0 / k)).astype(np.float32)
        b = (self.rng.standard_normal(oc) * 0.05).astype(np.float32) if bias else None
        wq, ws, wz = self._weight_q(wf)
        s_in = self.g.tensors[x.tid]["scale"]
        wdq = (wq.astype(np.float32) - wz) * (ws.reshape(-1, 1) if not self.u8 else ws[0])
        bq = self._bias_q(b, s_in, ws if not self.u8 else ws[0] * np.ones(oc, np.float32)) if bias else None
        a = x.act.reshape(x.act.shape[0], -1) @ self.torch.from_numpy(wdq).T
        if bq is not None:
            a = a + self.torch.from_numpy((bq.astype(np.float32) * np.float32(s_in) * (ws if not self.u8 else ws[0])).astype(np.float32))
        a = a.reshape(a.shape[0], oc, 1, 1)
        so, zo = self._act_q(a)
        return _H(self.g.fc(x.tid, wq, bq, ws, so, zo, weight_zero=wz), a)

    def pool(self, x, method, kernel, stride, pad=0, global_pool=False, caffe_flavor=0):
        """`pad` is the reference's pad_*_org; the real (possibly asymmetric) pads come from GraphDef.pool()."""
        F = self.torch.nn.functional
        tid = self.g.pool(x.tid, method, kernel, stride, pad, out_scale=1.0, out_zero=0, global_pool=global_pool,
                          caffe_flavor=caffe_flavor)
        L = self.g.layers[-1]
        if L["pool_global"]:
            a = x.act.mean(dim=(2, 3), keepdim=True) if method == abi.POOL_AVG else x.act.amax(dim=(2, 3), keepdim=True)
        else:
            # calibration only: windows clipped to the image (extra bottom/right rows never contribute)
            ph0, ph1, pw0, pw1 = L["pad_h0"], L["pad_h1"], L["pad_w0"], L["pad_w1"]
            if method == abi.POOL_MAX:
                xp = F.pad(x.act, (pw0, pw1, ph0, ph1), value=float("-inf"))
                a = F.max_pool2d(xp, kernel, stride)
            else:
                xp = F.pad(x.act, (pw0, pw1, ph0, ph1), value=0.0)
                ones = F.pad(self.torch.ones_like(x.act), (pw0, pw1, ph0, ph1), value=0.0)
                a = F.avg_pool2d(xp, kernel, stride) / F.avg_pool2d(ones, kernel, stride).clamp_min(1e-9)
            oh, ow = self.g.dims(tid)[2:]
            a = a[:, :, :oh, :ow]
        so, zo = self._act_q(a)
        if self.inplace and method == abi.POOL_MAX:
            src = self.g.tensors[x.tid]
            so, zo = src["scale"], src["zero_point"]
        self.g.tensors[tid]["scale"], self.g.tensors[tid]["zero_point"] = float(so), int(zo)
        return _H(tid, a)

    def relu(self, x, negative_slope=0.0):
        a = self.torch.nn.functional.leaky_relu(x.act, negative_slope) if negative_slope else self.torch.relu(x.act)
        so, zo = self._act_q(a)
        if self.inplace and not negative_slope:
            src = self.g.tensors[x.tid]
            so, zo = src["scale"], src["zero_point"]
        return _H(self.g.relu(x.tid, so, zo, negative_slope), a)

    def add(self, x, y):
        a = x.act + y.act
        so, zo = self._act_q(a)
        return _H(self.g.eltwise(x.tid, y.tid, so, zo, abi.ELT_SUM), a)

    def concat(self, xs):
        a = self.torch.cat([x.act for x in xs], dim=1)
        so, zo = self._act_q(a)
        return _H(self.g.concat([x.tid for x in xs], so, zo), a)

    def concat1(self, x):
        """Single-input Concat (a plain copy in the reference, concat_kernel_ref_int8.c:45-56)."""
        src = self.g.tensors[x.tid]
        return _H(self.g.concat([x.tid], src["scale"], src["zero_point"]), x.act)

    def upsample(self, x, scale):
        a = self.torch.nn.functional.interpolate(x.act, scale_factor=scale, mode="nearest")
        return _H(self.g.upsample(x.tid, scale), a)

    def identity(self, x):
        return _H(self.g.identity(x.tid), x.act)

    def flatten(self, x):
        return _H(self.g.flatten(x.tid), x.act.reshape(x.act.shape[0], -1, 1, 1))

    def softmax(self, x):
        a = self.torch.softmax(x.act, dim=1)
        # probabilities: the reference's models quantise them over [0, 1]
        so, zo = (np.float32(1.0 / 255.0), 0) if self.u8 else (np.float32(1.0 / 127.0), 0)
        return _H(self.g.softmax(x.tid, so, zo), a)

    def sigmoid(self, x):
        a = self.torch.sigmoid(x.act)
        so, zo = (np.float32(1.0 / 255.0), 0) if self.u8 else (np.float32(1.0 / 127.0), 0)
        return _H(self.g.sigmoid(x.tid, so, zo), a)

    def hardswish(self, x):
        a = self.torch.nn.functional.hardswish(x.act)
        so, zo = self._act_q(a)
        return _H(self.g.hardswish(x.tid, so, zo), a)

    def mul(self, x, y):
        a = x.act * y.act
        so, zo = self._act_q(a)
        return _H(self.g.eltwise(x.tid, y.tid, so, zo, abi.ELT_PROD), a)

    def finish(self, outs):
        for o in outs:
            self.g.mark_output(o.tid)
        return self.g

    def random_input(self, seed=42):
        rng = np.random.default_rng(seed)
        n, c, h, w = self.g.dims(self.input.tid)
        if self.u8:
            return rng.integers(0, 256, (n, c, h, w)).astype(np.uint8)
        return rng.integers(-127, 128, (n, c, h, w)).astype(np.int8)


def random_input(g, seed=42):
    rng = np.random.default_rng(seed)
    shape = g.dims(g.inputs[0])
    if g.data_type == abi.DT_UINT8:
        return rng.integers(0, 256, shape).astype(np.uint8)
    return rng.integers(-127, 128, shape).astype(np.int8)


def mobilenet_v1(data_type=abi.DT_INT8, batch=1, res=224, seed=1234, width=1.0, classes=1000, dw_recipe=None):
    """MobileNet-v1 224 as in benchmark/models/mobilenet_benchmark.tmfile: 3x3 s2 stem, 13 x (dw3x3 + pw1x1), global
    average pool, 1x1 classifier conv [1000,1024,1,1].  All convs carry a fused ReLU (activation = 0).
    dw_recipe: recipe of the depthwise layers; the reference uses its HCL kernel at batch 1 and conv_ref
    (conv_dw_hcl_x86.c:536 -> conv_ref.c) at batch > 1 for int8, conv_ref always for uint8."""
    b = QuantBuilder(data_type, batch, 3, res, res, seed)
    if dw_recipe is None:
        dw_recipe = abi.RECIPE_REF if (batch > 1 or data_type == abi.DT_UINT8) else abi.RECIPE_HCL
    ch = lambda c: max(8, int(c * width))
    x = b.conv(b.input, ch(32), 3, stride=2, pad=1, activation=0)
    cfg = [(64, 1), (128, 2), (128, 1), (256, 2), (256, 1), (512, 2), (512, 1), (512, 1), (512, 1), (512, 1), (512, 1),
           (1024, 2), (1024, 1)]
    for oc, s in cfg:
        c = b.g.dims(x.tid)[1]
        x = b.conv(x, c, 3, stride=s, pad=1, group=c, activation=0, recipe=dw_recipe)
        x = b.conv(x, ch(oc), 1, activation=0)
    x = b.pool(x, abi.POOL_AVG, 7, 1, global_pool=True)
    x = b.conv(x, classes, 1, activation=-1)
    return b.finish([x]), b


def tail_net(data_type=abi.DT_INT8, batch=2, seed=11):
    """The glue ops around the classifier / detection heads: sigmoid * x (SiLU as the reference's int8 graphs spell it),
    hardswish (uint8 only, like the reference), flatten of a C x H x W tensor, FC, softmax."""
    b = QuantBuilder(data_type, batch, 3, 12, 12, seed)
    x = b.conv(b.input, 24, 3, stride=1, pad=1, activation=-1)
    x = b.mul(x, b.sigmoid(x))
    if b.u8:
        x = b.hardswish(b.conv(x, 20, 1, activation=-1))
    x = b.pool(x, abi.POOL_MAX, 2, 2)
    f = b.flatten(x)
    y = b.softmax(b.fc(f, 10))
    return b.finish([y, x]), b


def tiny_net(data_type=abi.DT_INT8, batch=2, seed=7):
    """A few layers of every kind for fast CPU-oracle parity tests."""
    b = QuantBuilder(data_type, batch, 3, 20, 20, seed)
    x = b.conv(b.input, 16, 3, stride=2, pad=1, activation=0)
    x = b.conv(x, 16, 3, stride=1, pad=1, group=16, activation=0, recipe=abi.RECIPE_REF)
    x = b.conv(x, 32, 1, activation=0)
    y = b.conv(x, 32, 3, pad=1, activation=6)
    z = b.add(x, y)
    z = b.relu(z)
    p = b.pool(z, abi.POOL_MAX, 2, 2)
    q = b.conv(p, 24, 1, activation=-1)
    u = b.upsample(q, 2)
    cat = b.concat([u, z])
    l = b.relu(b.conv(cat, 40, 3, pad=1, activation=-1), negative_slope=0.1)
    g = b.pool(l, abi.POOL_AVG, 10, 1, global_pool=True)
    f = b.fc(g, 10)
    return b.finish([f, l]), b


def resnet50(data_type=abi.DT_UINT8, batch=1, res=224, seed=1234, width=1.0, classes=1000, blocks=(3, 4, 6, 3), softmax=False):
    """ResNet-50 as in benchmark/models/resnet50_benchmark.tmfile (Caffe layout): 7x7 s2 stem with fused ReLU, 3x3 s2
    max-pool (caffe_flavor 1 -> real pads 0,1,0,1), bottlenecks whose projection / first 1x1 carry the stride, Eltwise-sum
    followed by a STANDALONE ReLU node (each re-quantises), global average pool, FC 2048->1000 (Softmax stays on the CPU)."""
    b = QuantBuilder(data_type, batch, 3, res, res, seed, inplace=True)  # as the reference's quantisation tool writes its models
    ch = lambda c: max(16, int(c * width))
    x = b.conv(b.input, ch(64), 7, stride=2, pad=3, activation=0)
    x = b.pool(x, abi.POOL_MAX, 3, 2, 0, caffe_flavor=1)
    for stage, nblk in enumerate(blocks):
        mid, out = ch(64 << stage), ch(256 << stage)
        for i in range(nblk):
            stride = 2 if (i == 0 and stage > 0) else 1
            y = b.conv(x, mid, 1, stride=stride, activation=0)
            y = b.conv(y, mid, 3, pad=1, activation=0)
            y = b.conv(y, out, 1, activation=-1)
            if i == 0:
                sc = b.conv(x, out, 1, stride=stride, activation=-1)
                x = b.add(sc, y)
            else:
                x = b.add(x, y)
            x = b.relu(x)
    x = b.pool(x, abi.POOL_AVG, b.g.dims(x.tid)[2], 1, global_pool=True, caffe_flavor=1)
    x = b.fc(x, classes)
    if softmax:  # the tail of the reference's resnet50 tmfile (fc1000 -> prob)
        x = b.softmax(x)
    return b.finish([x]), b


def yolov3_tiny(data_type=abi.DT_UINT8, batch=1, res=416, seed=1234, width=1.0, head=255):
    """YOLOv3-tiny as in benchmark/models/yolov3_tiny_benchmark.tmfile: 3x3 convs WITHOUT fused activation, each followed
    by a standalone leaky ReLU (slope 0.1), 2x2 max-pools with caffe_flavor 2 / pad_org 1 (real pads 0,1,0,1; the last one
    has stride 1), a single-input Concat, nearest Upsample x2, a two-input Concat (384 ch) and two 1x1 heads of 255 channels
    followed by Dropout (identity)."""
    b = QuantBuilder(data_type, batch, 3, res, res, seed, inplace=True)  # as the reference's quantisation tool writes its models
    ch = lambda c: max(8, int(c * width))

    def cbl(x, oc, k):
        return b.relu(b.conv(x, oc, k, pad=k // 2, activation=-1), negative_slope=0.1)

    x = b.input
    feats = []
    for i, oc in enumerate((16, 32, 64, 128, 256, 512)):
        x = cbl(x, ch(oc), 3)
        feats.append(x)
        x = b.pool(x, abi.POOL_MAX, 2, 2 if i < 5 else 1, 1, caffe_flavor=2)
    x = cbl(x, ch(1024), 3)
    x = cbl(x, ch(256), 1)
    route = b.concat1(x)
    y = cbl(route, ch(128), 1)
    y = b.upsample(y, 2)
    y = b.concat([y, feats[4]])
    y = cbl(y, ch(256), 3)
    big = cbl(x, ch(512), 3)
    o26 = b.identity(b.conv(y, head, 1, activation=-1))
    o13 = b.identity(b.conv(big, head, 1, activation=-1))
    return b.finish([o26, o13]), b


def yolov5s(data_type=abi.DT_INT8, batch=1, res=640, seed=1234, width=1.0, head=255):
    """YOLOv5s (v5.0 graph: Focus, C3 x {1,3,3,1}, SPP 5/9/13, PANet head, three 1x1 detection convs) in the form the reference
    pipeline runs it (tools/optimize/yolov5s-opt.py:114-156 cuts the Focus slicing and the post-processing out of the graph, so
    the network input is the app-sliced [N, 12, res/2, res/2] tensor -- examples/tm_yolov5s.cpp -- and the outputs are the three
    raw head tensors).  Activation: the script rewrites SiLU to HardSwish, which the reference CPU device only has for fp32 /
    uint8 (hardswish_ref.c:59-66); an int8 model therefore keeps SiLU as Sigmoid + Eltwise-PROD (sigmoid_ref.c:84,
    eltwise_ref.c:589).  BatchNorm is folded into the convolutions' bias as in every reference tmfile."""
    b = QuantBuilder(data_type, batch, 12, res // 2, res // 2, seed, inplace=True)
    ch = lambda c: max(8, int(c * width))

    def act(x):
        return b.hardswish(x) if b.u8 else b.mul(x, b.sigmoid(x))

    def conv(x, oc, k=1, s=1):
        return act(b.conv(x, oc, k, stride=s, pad=k // 2, activation=-1))

    def c3(x, oc, n, shortcut=True):
        c_ = oc // 2
        y = conv(x, c_, 1)
        for _ in range(n):
            z = conv(conv(y, c_, 1), c_, 3)
            y = b.add(y, z) if shortcut else z
        return conv(b.concat([y, conv(x, c_, 1)]), oc, 1)

    x = conv(b.input, ch(32), 3)                    # Focus convolution
    x = conv(x, ch(64), 3, 2)
    x = c3(x, ch(64), 1)
    x = conv(x, ch(128), 3, 2)
    p3 = c3(x, ch(128), 3)
    x = conv(p3, ch(256), 3, 2)
    p4 = c3(x, ch(256), 3)
    x = conv(p4, ch(512), 3, 2)
    y = conv(x, ch(256), 1)                          # SPP
    x = conv(b.concat([y] + [b.pool(y, abi.POOL_MAX, k, 1, k // 2) for k in (5, 9, 13)]), ch(512), 1)
    x = c3(x, ch(512), 1, shortcut=False)
    h10 = conv(x, ch(256), 1)
    x = c3(b.concat([b.upsample(h10, 2), p4]), ch(256), 1, shortcut=False)
    h14 = conv(x, ch(128), 1)
    o17 = c3(b.concat([b.upsample(h14, 2), p3]), ch(128), 1, shortcut=False)
    o20 = c3(b.concat([conv(o17, ch(128), 3, 2), h14]), ch(256), 1, shortcut=False)
    o23 = c3(b.concat([conv(o20, ch(256), 3, 2), h10]), ch(512), 1, shortcut=False)
    outs = [b.conv(o, head, 1, activation=-1) for o in (o17, o20, o23)]
    return b.finish(outs), b
