#!/usr/bin/env python3
"""Debug helper: run a model file of oracle/_ref/models through run_graph() on device B200 and compare with the oracle."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.pyoracle import Oracle, Reference, run_tmfile  # noqa: E402
from tengine_b200 import abi, workloads  # noqa: E402

g, b = workloads.resnet50(abi.DT_UINT8, batch=4, softmax=True)
outs = [g.outputs[0], g.layers[-1]["inputs"][0]]
x = b.random_input(21)
ref = Reference(libdir=os.path.join(ROOT, "build", "tengine"))
want = Oracle().run(g, [x], uint8_mode=0)
for dev in ("B200", "CPU"):
    got, ms = run_tmfile(ref, os.path.join(ROOT, "oracle/_ref/models/resnet50_uint8.tmfile"), x, [g.dims(t) for t in outs], device=dev)
    for o, t, nm in zip(got, outs, ("prob", "fc")):
        d = np.abs(o.astype(int) - want[t].astype(int))
        print(dev, nm, "max diff", d.max(), "count", int((d > 0).sum()), "where", np.argwhere(d > 0)[:6].tolist())
    fc = got[1].reshape(4, -1)
    for n in range(4):
        top = np.argsort(-fc[n].astype(int))[:3]
        print(dev, n, top, fc[n][top], got[0].reshape(4, -1)[n][top], "oracle", want[outs[1]].reshape(4, -1)[n][top], want[outs[0]].reshape(4, -1)[n][top])
