#!/usr/bin/env python3
"""Debug: ResNet-50 uint8 batch 4 (the tmfile test's case) through the direct ABI, default plan and per-layer plan, with the
GraphDef's own tensor numbering and with the device glue's numbering (outputs numbered before inputs)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.pyoracle import Oracle  # noqa: E402
from tengine_b200 import abi, workloads  # noqa: E402
from tengine_b200 import runtime as rt  # noqa: E402
from tengine_b200.graphdef import GraphDef  # noqa: E402


def renumber(g):
    """tensor ids in the order the device glue assigns them: per layer output first, then inputs"""
    order = []
    for L in g.layers:
        for t in [L["output"]] + L["inputs"]:
            if t not in order:
                order.append(t)
    for t in g.inputs + g.outputs:
        if t not in order:
            order.append(t)
    m = {old: new for new, old in enumerate(order)}
    h = GraphDef(g.data_type)
    h.tensors = [g.tensors[old] for old in order]
    for L in g.layers:
        M = dict(L)
        M["inputs"] = [m[t] for t in L["inputs"]]
        M["output"] = m[L["output"]]
        h.layers.append(M)
    h.inputs = [m[t] for t in g.inputs]
    h.outputs = [m[t] for t in g.outputs]
    return h, m


net = sys.argv[1] if len(sys.argv) > 1 else "resnet50"
if net == "resnet50":
    g, b = workloads.resnet50(abi.DT_UINT8, batch=4, softmax=True)
    g.mark_output(g.layers[-1]["inputs"][0])
else:
    g, b = workloads.yolov3_tiny(abi.DT_UINT8, batch=4)
x = b.random_input(21)
want = Oracle().run(g, [x], uint8_mode=0)
ctx = rt.Context(0)
for name, gg, mp in (("graphdef numbering", g, None),) + tuple((("glue numbering",) + renumber(g),)):
    for flags, fl in ((abi.PRERUN_DEFAULT, "default"), (abi.PRERUN_NO_GRAPH, "no_graph")):
        gr = rt.Graph(ctx, gg, flags)
        outs = gr.run([x])
        res = []
        for o, t in zip(outs, g.outputs):
            d = np.abs(o.astype(int) - want[t].astype(int))
            res.append((int(d.max()), int((d > 0).sum())))
        print(name, fl, "outputs (max diff, count):", res)
        if flags == abi.PRERUN_NO_GRAPH:
            ks = gr.layer_kernels()
            bad = 0
            for li, L in enumerate(g.layers):
                tid = L["output"] if mp is None else mp[L["output"]]
                got = gr.read_tensor(tid)
                d = np.abs(got.astype(int) - want[L["output"]].astype(int))
                if d.max() > 0:
                    print("   first differing layer", li, abi.OP_NAMES[L["op"]], ks[li], "max", int(d.max()), "count", int((d > 0).sum()), "dims", g.dims(L["output"]))
                    bad += 1
                    if bad >= 3:
                        break
        gr.close()
