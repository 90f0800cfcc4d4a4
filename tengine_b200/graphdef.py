"""Host-side description of a quantised subgraph: the tb200_tensor_desc / tb200_layer_desc tables of
include/tengine_b200.h, built from numpy arrays.

The field names are the reference's (conv_param, pool_param, ... -- operator/prototype/*_param.h); shapes are
inferred with the reference's rules (operator/prototype/convolution.c:35-145, pooling.c:34-105).
Used by tests/ and bench.py; the production producer of these tables is the C++ Tengine device glue.
"""
import ctypes as C

import numpy as np

from . import abi


def conv_out_dim(i, k, s, p0, p1, d):
    return (i + p0 + p1 - (d * (k - 1) + 1)) // s + 1


class GraphDef:
    def __init__(self, data_type):
        assert data_type in (abi.DT_INT8, abi.DT_UINT8)
        self.data_type = data_type
        self.np_dtype = np.int8 if data_type == abi.DT_INT8 else np.uint8
        self.tensors = []  # dicts: dims, scale, zero_point
        self.layers = []  # dicts mirroring tb200_layer_desc + numpy constants
        self.inputs = []
        self.outputs = []
        self._keep = []

    # ---- tensors -------------------------------------------------------------------------------
    def add_tensor(self, dims, scale, zero_point=0):
        assert len(dims) == 4
        self.tensors.append(dict(dims=tuple(int(d) for d in dims), scale=float(np.float32(scale)),
                                 zero_point=int(zero_point)))
        return len(self.tensors) - 1

    def input(self, n, c, h, w, scale, zero_point=0):
        t = self.add_tensor((n, c, h, w), scale, zero_point)
        self.inputs.append(t)
        return t

    def mark_output(self, t):
        self.outputs.append(t)

    def dims(self, t):
        return self.tensors[t]["dims"]

    def _layer(self, op, inputs, out, **kw):
        L = dict(op=op, inputs=list(inputs), output=out, kernel_h=1, kernel_w=1, stride_h=1, stride_w=1, pad_h0=0,
                 pad_h1=0, pad_w0=0, pad_w1=0, dilation_h=1, dilation_w=1, group=1, activation=-1,
                 recipe=abi.RECIPE_HCL, pool_method=0, pool_global=0, caffe_flavor=0, negative_slope=0.0,
                 elt_type=abi.ELT_SUM, axis=1, up_scale=1, weight=None, bias=None, weight_scales=None,
                 weight_zero=0, bias_scale=0.0)
        L.update(kw)
        self.layers.append(L)
        return out

    # ---- ops -----------------------------------------------------------------------------------
    def conv(self, x, weight, bias, weight_scales, out_scale, out_zero=0, stride=1, pad=0, dilation=1, group=1,
             activation=-1, recipe=abi.RECIPE_HCL, weight_zero=0):
        n, c, h, w = self.dims(x)
        oc, cg, kh, kw = weight.shape
        assert cg * group == c, (cg, group, c)
        sh, sw = (stride, stride) if np.isscalar(stride) else stride
        dh, dw = (dilation, dilation) if np.isscalar(dilation) else dilation
        if np.isscalar(pad):
            p = (pad, pad, pad, pad)
        else:
            p = tuple(pad)  # h0, h1, w0, w1
        oh = conv_out_dim(h, kh, sh, p[0], p[1], dh)
        ow = conv_out_dim(w, kw, sw, p[2], p[3], dw)
        out = self.add_tensor((n, oc, oh, ow), out_scale, out_zero)
        weight = np.ascontiguousarray(weight, dtype=self.np_dtype)
        ws = np.ascontiguousarray(weight_scales, dtype=np.float32).reshape(-1)
        assert ws.size == (oc if self.data_type == abi.DT_INT8 else 1)
        b = None if bias is None else np.ascontiguousarray(bias, dtype=np.int32)
        return self._layer(abi.OP_CONV, [x], out, kernel_h=kh, kernel_w=kw, stride_h=sh, stride_w=sw, pad_h0=p[0],
                           pad_h1=p[1], pad_w0=p[2], pad_w1=p[3], dilation_h=dh, dilation_w=dw, group=group,
                           activation=activation, recipe=recipe, weight=weight, bias=b, weight_scales=ws,
                           weight_zero=weight_zero,
                           bias_scale=float(np.float32(self.tensors[x]["scale"]) * np.float32(ws[0])))

    def fc(self, x, weight, bias, weight_scales, out_scale, out_zero=0, weight_zero=0):
        n, c, h, w = self.dims(x)
        oc, k = weight.shape
        assert k == c * h * w
        out = self.add_tensor((n, oc, 1, 1), out_scale, out_zero)
        ws = np.ascontiguousarray(weight_scales, dtype=np.float32).reshape(-1)
        b = None if bias is None else np.ascontiguousarray(bias, dtype=np.int32)
        return self._layer(abi.OP_FC, [x], out, weight=np.ascontiguousarray(weight, dtype=self.np_dtype), bias=b,
                           weight_scales=ws, weight_zero=weight_zero,
                           bias_scale=float(np.float32(self.tensors[x]["scale"]) * np.float32(ws[0])))

    def pool(self, x, method, kernel, stride, pad=0, out_scale=None, out_zero=None, global_pool=False,
             caffe_flavor=0):
        n, c, h, w = self.dims(x)
        kh, kw = (kernel, kernel) if np.isscalar(kernel) else kernel
        sh, sw = (stride, stride) if np.isscalar(stride) else stride
        p = (pad, pad, pad, pad) if np.isscalar(pad) else tuple(pad)
        if global_pool or (kh == h and kw == w and not any(p)):
            kh, kw, sh, sw, p, oh, ow, global_pool = h, w, 1, 1, (0, 0, 0, 0), 1, 1, True
        else:
            # operator/prototype/pooling_param.h:59-105 calc_output_size + calc_real_pads; `pad` is pad_*_org
            def out_dim(i, k, s_, po):
                if caffe_flavor == 1:
                    o = 2 + (i - k + 2 * po - 1) // s_
                    if po > 0 and (o - 1) * s_ >= i + po:
                        o -= 1
                    return o
                if caffe_flavor == 2:
                    return 1 + (i - k + po) // s_
                return 1 + (i - k + 2 * po) // s_

            oh, ow = out_dim(h, kh, sh, p[0]), out_dim(w, kw, sw, p[2])
            if caffe_flavor == 2:
                p = (p[0] // 2, p[0] - p[0] // 2, p[2] // 2, p[2] - p[2] // 2)
            else:
                p = (p[0], max((oh - 1) * sh + kh - h, 0) - p[0], p[2], max((ow - 1) * sw + kw - w, 0) - p[2])
        src = self.tensors[x]
        out = self.add_tensor((n, c, oh, ow), src["scale"] if out_scale is None else out_scale,
                              src["zero_point"] if out_zero is None else out_zero)
        return self._layer(abi.OP_POOL, [x], out, pool_method=method, kernel_h=kh, kernel_w=kw, stride_h=sh,
                           stride_w=sw, pad_h0=p[0], pad_h1=p[1], pad_w0=p[2], pad_w1=p[3],
                           pool_global=int(global_pool), caffe_flavor=caffe_flavor)

    def relu(self, x, out_scale=None, out_zero=None, negative_slope=0.0):
        src = self.tensors[x]
        out = self.add_tensor(src["dims"], src["scale"] if out_scale is None else out_scale,
                              src["zero_point"] if out_zero is None else out_zero)
        return self._layer(abi.OP_RELU, [x], out, negative_slope=float(negative_slope))

    def eltwise(self, a, b, out_scale, out_zero=0, elt_type=abi.ELT_SUM):
        assert self.dims(a) == self.dims(b)
        out = self.add_tensor(self.dims(a), out_scale, out_zero)
        return self._layer(abi.OP_ELTWISE, [a, b], out, elt_type=elt_type)

    def concat(self, xs, out_scale, out_zero=0):
        n, _, h, w = self.dims(xs[0])
        c = sum(self.dims(x)[1] for x in xs)
        out = self.add_tensor((n, c, h, w), out_scale, out_zero)
        return self._layer(abi.OP_CONCAT, xs, out, axis=1)

    def upsample(self, x, scale):
        n, c, h, w = self.dims(x)
        src = self.tensors[x]
        out = self.add_tensor((n, c, h * scale, w * scale), src["scale"], src["zero_point"])
        return self._layer(abi.OP_UPSAMPLE, [x], out, up_scale=int(scale))

    def identity(self, x):
        src = self.tensors[x]
        out = self.add_tensor(src["dims"], src["scale"], src["zero_point"])
        return self._layer(abi.OP_IDENTITY, [x], out)

    def softmax(self, x, out_scale, out_zero=0, axis=1):
        return self._layer(abi.OP_SOFTMAX, [x], self.add_tensor(self.dims(x), out_scale, out_zero), axis=int(axis))

    def sigmoid(self, x, out_scale, out_zero=0):
        return self._layer(abi.OP_SIGMOID, [x], self.add_tensor(self.dims(x), out_scale, out_zero))

    def hardswish(self, x, out_scale, out_zero=0):
        return self._layer(abi.OP_HARDSWISH, [x], self.add_tensor(self.dims(x), out_scale, out_zero))

    def flatten(self, x):
        """OP_FLATTEN (axis 1..3) / OP_RESHAPE to [N, C*H*W, 1, 1]: the same bytes in NCHW order, same quantisation."""
        n, c, h, w = self.dims(x)
        src = self.tensors[x]
        return self._layer(abi.OP_RESHAPE, [x], self.add_tensor((n, c * h * w, 1, 1), src["scale"], src["zero_point"]))

    # ---- C tables ------------------------------------------------------------------------------
    def c_tables(self):
        """(TensorDesc[], LayerDesc[]) ctypes arrays; numpy constants stay referenced by self."""
        T = (abi.TensorDesc * len(self.tensors))()
        for i, t in enumerate(self.tensors):
            T[i].data_type = self.data_type
            for k in range(4):
                T[i].dims[k] = t["dims"][k]
            T[i].scale = t["scale"]
            T[i].zero_point = t["zero_point"]
        Ls = (abi.LayerDesc * len(self.layers))()
        for i, L in enumerate(self.layers):
            d = Ls[i]
            d.op = L["op"]
            d.num_inputs = len(L["inputs"])
            for k, v in enumerate(L["inputs"]):
                d.inputs[k] = v
            d.output = L["output"]
            for f in ("kernel_h", "kernel_w", "stride_h", "stride_w", "pad_h0", "pad_h1", "pad_w0", "pad_w1",
                      "dilation_h", "dilation_w", "group", "activation", "recipe", "pool_method", "pool_global",
                      "caffe_flavor", "negative_slope", "elt_type", "axis", "up_scale", "weight_zero", "bias_scale"):
                setattr(d, f, L[f])
            for f in ("weight", "bias", "weight_scales"):
                a = L[f]
                setattr(d, f, None if a is None else a.ctypes.data)
        self._keep = [T, Ls]
        return T, Ls

    def id_array(self, ids):
        return (C.c_int32 * len(ids))(*ids)

    def numel(self, t):
        return int(np.prod(self.dims(t)))

    # algorithmic work, SURVEY.md 8(d): ops = 2*MACs over conv+fc; bytes = their in+out activations (1 B/elem)
    # + weights (1 B) + bias (4 B), per launch of the whole graph
    def work(self):
        ops = 0
        byts = 0
        for L in self.layers:
            if L["op"] not in (abi.OP_CONV, abi.OP_FC):
                continue
            oc = self.dims(L["output"])[1]
            k = L["weight"].size // oc
            ops += 2 * self.numel(L["output"]) * k
            byts += self.numel(L["inputs"][0]) + self.numel(L["output"]) + L["weight"].size
            if L["bias"] is not None:
                byts += 4 * oc
        return ops, byts

    # ---- (de)serialisation for committed golden fixtures --------------------------------------------
    _SCALARS = ("op", "output", "kernel_h", "kernel_w", "stride_h", "stride_w", "pad_h0", "pad_h1", "pad_w0", "pad_w1",
                "dilation_h", "dilation_w", "group", "activation", "recipe", "pool_method", "pool_global", "caffe_flavor",
                "negative_slope", "elt_type", "axis", "up_scale", "weight_zero", "bias_scale")

    def to_dict(self):
        import json

        d = {"meta": np.frombuffer(json.dumps({
            "data_type": self.data_type, "tensors": self.tensors, "inputs": self.inputs, "outputs": self.outputs,
            "layers": [{**{k: L[k] for k in self._SCALARS}, "inputs": L["inputs"]} for L in self.layers]}).encode(),
            dtype=np.uint8)}
        for i, L in enumerate(self.layers):
            for f in ("weight", "bias", "weight_scales"):
                if L[f] is not None:
                    d[f"L{i}_{f}"] = L[f]
        return d

    @classmethod
    def from_dict(cls, d):
        import json

        meta = json.loads(bytes(d["meta"]).decode())
        g = cls(meta["data_type"])
        g.tensors = [dict(dims=tuple(t["dims"]), scale=t["scale"], zero_point=t["zero_point"]) for t in meta["tensors"]]
        g.inputs, g.outputs = list(meta["inputs"]), list(meta["outputs"])
        for i, L in enumerate(meta["layers"]):
            L = dict(L)
            for f in ("weight", "bias", "weight_scales"):
                k = f"L{i}_{f}"
                L[f] = np.ascontiguousarray(d[k]) if k in d else None
            g.layers.append(L)
        return g
