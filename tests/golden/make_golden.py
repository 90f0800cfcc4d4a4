#!/usr/bin/env python3
"""Generate tests/golden/ref_*.npz by running the UNMODIFIED reference (oracle/_ref, built by oracle/build_ref.py from
/root/reference) on seeded inputs.  Run in the container that has /root/reference; the fixtures are committed so that
the GPU box (which has neither) can still check the device against reference-produced bytes.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.pyoracle import Reference  # noqa: E402
from tengine_b200 import abi, workloads  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def save(name, g, x, ref, **env):
    want = [L["output"] for L in g.layers]
    r, _ = ref.run(g, [x], want=want, env=env or None)
    d = g.to_dict()
    d["input"] = x
    for t in want:
        d[f"ref_t{t}"] = r[t]
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    print(name, {k: v.shape for k, v in d.items() if k.startswith("ref_")}.__len__(), "tensors")


def main():
    ref = Reference()
    for dt, tag in ((abi.DT_INT8, "int8"), (abi.DT_UINT8, "uint8")):
        g, b = workloads.tiny_net(dt, batch=2, seed=7)
        save(f"ref_tiny_{tag}", g, b.random_input(3), ref)
        # MobileNet-v1 at reduced width/resolution (same layer kinds and order as the benchmark model) at batch 1:
        # exercises the reference's HCL selections (conv_hcl 1x1, conv_dw_hcl, conv_direct_hcl_int8 3x3)
        g, b = workloads.mobilenet_v1(dt, batch=1, res=64, seed=11, width=0.25, classes=32)
        save(f"ref_mobilenet025_{tag}", g, b.random_input(5), ref)
    # ResNet-50 / YOLOv3-tiny at reduced width (BASELINE.json C3 / C4 structure), int8, batch 1, quantised with the in-place scales of
    # the reference's tool: 7x7 stem, bottlenecks with eltwise + standalone ReLU, max pooling, leaky ReLU, concat, upsample
    g, b = workloads.resnet50(abi.DT_INT8, batch=1, res=96, width=0.25, classes=40, seed=5)
    save("ref_resnet50_small_int8", g, b.random_input(3), ref)
    g, b = workloads.yolov3_tiny(abi.DT_INT8, batch=1, res=96, width=0.25, head=27, seed=6)
    save("ref_yolov3_tiny_small_int8", g, b.random_input(3), ref)


if __name__ == "__main__":
    main()
