#!/usr/bin/env python3
"""Per-layer kernel times of a workload (events around every launch, no CUDA graph)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from tengine_b200 import abi, workloads  # noqa: E402
from tengine_b200 import runtime as rt  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 256
net = sys.argv[2] if len(sys.argv) > 2 else "mobilenet_v1"
dt = abi.DT_UINT8 if (len(sys.argv) > 3 and sys.argv[3] == "uint8") else abi.DT_INT8
res = 416 if net == "yolov3_tiny" else (640 if net == "yolov5s" else 224)
g, b = getattr(workloads, net)(dt, batch=batch, res=res)
ctx = rt.Context(0)
graph = rt.Graph(ctx, g, abi.PRERUN_NO_GRAPH)
graph.upload(0, b.random_input(1))
graph.sync()
for _ in range(3):
    graph.profile()
ms = np.mean([graph.profile() for _ in range(5)], axis=0)
for i, (k, t) in enumerate(zip(graph.layer_kernels(), ms)):
    L = g.layers[i]
    print(f"layer {i:2d} {k:26s} {t * 1000:8.1f} us   out {tuple(g.dims(L['output']))}")
print("total", ms.sum() * 1000)
