#!/usr/bin/env python3
"""Run the first (stem) layer of MobileNet-v1 alone and compare with the CPU oracle (debug aid)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from oracle.pyoracle import Oracle  # noqa: E402
from tengine_b200 import abi, workloads  # noqa: E402
from tengine_b200 import runtime as rt  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 2
g, b = workloads.mobilenet_v1(abi.DT_INT8, batch=batch, res=224)
x = b.random_input(3)
ctx = rt.Context(0)
graph = rt.Graph(ctx, g, abi.PRERUN_NO_GRAPH)
y = graph.run([x])[0]
t = graph.read_tensor(g.layers[0]["output"])
want = Oracle().run(g, [x])
w0 = want[g.layers[0]["output"]]
print("kernel", graph.layer_kernels()[0], "stem equal:", np.array_equal(t, w0), "mismatches", int((t != w0).sum()), "of", t.size)
print("final equal:", np.array_equal(y, want[g.outputs[0]]))
