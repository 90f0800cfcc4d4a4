#!/usr/bin/env python3
"""Write the quantised tmfiles the integration tests feed to the UNMODIFIED apps (tm_classification_int8/uint8,
tm_benchmark) into oracle/_ref/models/, with the reference's own tmfile writer (tools/save_graph/save_graph.cpp through
oracle/ref_shim_save.cpp).  The graphs are this repo's synthetic workloads (seeded weights; there is no network for
checkpoints), quantised as tools/quantize/quant_save_graph.cpp does.  TEST INFRASTRUCTURE; needs oracle/_ref (i.e.
/root/reference at build time).  The files are build outputs: git-ignored, shipped to the GPU box by gpurun.
usage: make_models.py [--force]"""
import os
import struct
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
OUT = os.path.join(HERE, "_ref", "models")


def write_bmp(path, h=256, w=256, seed=7):
    """A smooth synthetic RGB image (24-bit BMP, the format stb_image in the examples decodes)."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    img = np.zeros((h, w, 3), np.float32)
    for c in range(3):
        for _ in range(6):
            fx, fy, ph = rng.uniform(0.5, 6), rng.uniform(0.5, 6), rng.uniform(0, 6.28)
            img[..., c] += rng.uniform(10, 40) * np.sin(xx * fx * 6.28 / w + yy * fy * 6.28 / h + ph)
    img = np.clip(img + 128, 0, 255).astype(np.uint8)
    row = (w * 3 + 3) // 4 * 4
    data = bytearray()
    for y in range(h - 1, -1, -1):
        line = img[y, :, ::-1].tobytes()
        data += line + b"\0" * (row - len(line))
    hdr = b"BM" + struct.pack("<IHHI", 54 + len(data), 0, 0, 54) + struct.pack("<IiiHHIIiiII", 40, w, h, 1, 24, 0, len(data), 2835, 2835, 0, 0)
    open(path, "wb").write(hdr + bytes(data))


def resnet50_two_outputs(workloads, abi):
    """ResNet-50 uint8 with the reference model's tail (fc1000 -> Softmax); the FC tensor is a second graph output so that tests
    can compare logits as well (a softmax over random-weight logits turns a 1-LSB difference into a different arg-max)."""
    g, b = workloads.resnet50(abi.DT_UINT8, batch=1, softmax=True)
    g.mark_output(g.layers[-1]["inputs"][0])
    return g, b


def main(force=False):
    from oracle.pyoracle import Reference, save_tmfile
    from tengine_b200 import abi, workloads

    if not Reference.available():
        print("[models] oracle/_ref not built: skipped")
        return
    os.makedirs(OUT, exist_ok=True)
    ref = Reference()
    jobs = [("mobilenet_v1_int8", lambda: workloads.mobilenet_v1(abi.DT_INT8, batch=1)),
            ("mobilenet_v1_uint8", lambda: workloads.mobilenet_v1(abi.DT_UINT8, batch=1)),
            ("resnet50_uint8", lambda: resnet50_two_outputs(workloads, abi)),
            ("yolov3_tiny_uint8", lambda: workloads.yolov3_tiny(abi.DT_UINT8, batch=1))]
    for name, build in jobs:
        path = os.path.join(OUT, name + ".tmfile")
        if os.path.exists(path) and not force:
            continue
        g, _ = build()
        save_tmfile(ref, g, path)
        print(f"[models] {path} ({os.path.getsize(path) / 1e6:.1f} MB)")
    bmp = os.path.join(OUT, "test.bmp")
    if not os.path.exists(bmp) or force:
        write_bmp(bmp)


if __name__ == "__main__":
    main("--force" in sys.argv)
