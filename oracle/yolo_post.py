"""CPU restatement of the detection post-processing of examples/tm_yolov3_tiny_uint8.cpp (TEST INFRASTRUCTURE; checker of
tb200_graph_yolo_detect):  dequantisation (:464-478), generate_proposals (:176-250), qsort_descent_inplace (:57-100),
nms_sorted_bboxes (:102-132), cv::Rect_<float> intersection / area.

PINNED (round 2): tests/test_yolo_post_pinned.py compares this file, box for box and bit for bit, with the example's own functions
compiled from the unmodified source (oracle/yolo_example_shim.cpp + oracle/cvstub, built into oracle/_ref/libyolo_example.so) and
with a committed fixture of their output.  What the pin corrected: `exp()` on a float argument resolves to the FLOAT overload in
that translation unit (its <cmath> / <math.h> bring std::exp(float) into scope), so sigmoid and exp(dw) * anchor are float
arithmetic throughout -- not double, as a reading of the text suggested.  expf comes from the C library (ctypes), the same function
the compiled example and the device's host-side table builder call.
"""
import ctypes
import ctypes.util

import numpy as np

f32 = np.float32
_libm = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
_libm.expf.restype = ctypes.c_float
_libm.expf.argtypes = [ctypes.c_float]


def _expf(x):
    return f32(_libm.expf(float(f32(x))))


def _sigmoid(x):
    return f32(f32(1.0) / f32(f32(1.0) + _expf(-f32(x))))  # static_cast<float>(1.f / (1.f + exp(-x))), exp = the float overload


def generate_proposals(stride, feat, anchors6, num_classes, prob_threshold):
    """feat: dequantised [3*(5+cls), H, W] float32 of ONE image.  Returns [(x, y, w, h, prob, label)] in the example's order."""
    _, fh, fw = feat.shape
    out = []
    per = num_classes + 5
    for h in range(fh):
        for w in range(fw):
            for a in range(3):
                scores = feat[a * per + 5:a * per + 5 + num_classes, h, w]
                cls = int(np.argmax(scores))  # first maximum == the example's strict `>` scan
                final = f32(_sigmoid(feat[a * per + 4, h, w]) * _sigmoid(scores[cls]))
                if final >= f32(prob_threshold):
                    dx, dy = _sigmoid(feat[a * per + 0, h, w]), _sigmoid(feat[a * per + 1, h, w])
                    dw, dh = feat[a * per + 2, h, w], feat[a * per + 3, h, w]
                    pred_x = f32(f32(f32(w) + dx) * f32(stride))
                    pred_y = f32(f32(f32(h) + dy) * f32(stride))
                    pred_w = f32(_expf(dw) * f32(anchors6[2 * a]))      # float exp(dw) * anchor_w
                    pred_h = f32(_expf(dh) * f32(anchors6[2 * a + 1]))
                    x0, y0 = f32(pred_x - f32(pred_w * f32(0.5))), f32(pred_y - f32(pred_h * f32(0.5)))
                    x1, y1 = f32(pred_x + f32(pred_w * f32(0.5))), f32(pred_y + f32(pred_h * f32(0.5)))
                    out.append([x0, y0, f32(x1 - x0), f32(y1 - y0), final, cls])
    return out


def qsort_descent_inplace(v, left, right):
    i, j = left, right
    p = v[(left + right) // 2][4]
    while i <= j:
        while v[i][4] > p:
            i += 1
        while v[j][4] < p:
            j -= 1
        if i <= j:
            v[i], v[j] = v[j], v[i]
            i += 1
            j -= 1
    if left < j:
        qsort_descent_inplace(v, left, j)
    if i < right:
        qsort_descent_inplace(v, i, right)


def _inter_area(a, b):
    x1, y1 = max(a[0], b[0]), max(a[1], b[1])
    w = f32(min(f32(a[0] + a[2]), f32(b[0] + b[2])) - x1)
    h = f32(min(f32(a[1] + a[3]), f32(b[1] + b[3])) - y1)
    return f32(0) if (w <= 0 or h <= 0) else f32(w * h)


def nms_sorted_bboxes(v, nms_threshold):
    picked = []
    areas = [f32(o[2] * o[3]) for o in v]
    for i, a in enumerate(v):
        keep = True
        for j in picked:
            inter = _inter_area(a, v[j])
            union = f32(f32(areas[i] + areas[j]) - inter)
            with np.errstate(divide="ignore", invalid="ignore"):
                if f32(inter / union) > f32(nms_threshold):
                    keep = False
        if keep:
            picked.append(i)
    return picked


def detect(outputs, scales, zeros, heads, num_classes, prob_threshold, nms_threshold):
    """outputs: per head a quantised NCHW array [N, 3*(5+cls), H, W] (uint8 or int8); heads: [(output index, stride, anchors6)] in
    proposal order.  Returns per image the kept boxes [(x, y, w, h, prob, label)]."""
    n = outputs[0].shape[0]
    res = []
    for img in range(n):
        props = []
        for (oi, stride, anchors6) in heads:
            q = outputs[oi][img].astype(np.float32)
            feat = ((q - f32(zeros[oi])) * f32(scales[oi])).astype(np.float32)
            props += generate_proposals(stride, feat, anchors6, num_classes, prob_threshold)
        if props:
            qsort_descent_inplace(props, 0, len(props) - 1)
        res.append([props[i] for i in nms_sorted_bboxes(props, nms_threshold)])
    return res
