"""ctypes front-ends of the TEST INFRASTRUCTURE libraries (never imported by tengine_b200/):

  Oracle        oracle/libtb200_oracle.so   the CPU restatement (tb200_oracle.c)
  Reference     oracle/_ref/libref_shim.so  the unmodified reference driven through ref_shim.c

Both consume a tengine_b200.graphdef.GraphDef and host NCHW numpy arrays.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def build(quiet=True):
    """Compile the C restatement (and, when /root/reference exists, oracle/_ref). Building the checker is not
    using it."""
    r = subprocess.run(["make", "-C", HERE], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + r.stdout[-2000:] + r.stderr[-2000:])
    if not quiet:
        print(r.stdout[-800:])


class Oracle:
    def __init__(self):
        path = os.path.join(HERE, "libtb200_oracle.so")
        if not os.path.exists(path):
            build()
        self.lib = C.CDLL(path)
        self.lib.tb200_oracle_run.restype = C.c_int
        self.lib.tb200_oracle_run.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int]

    def run(self, g, inputs, uint8_mode=0):
        """Run every layer; returns the list of all tensors (numpy NCHW), index = tensor id."""
        T, L = g.c_tables()
        bufs = [np.zeros(g.dims(i), dtype=g.np_dtype) for i in range(len(g.tensors))]
        for t, x in zip(g.inputs, inputs):
            assert x.shape == g.dims(t) and x.dtype == g.np_dtype, (x.shape, g.dims(t), x.dtype)
            bufs[t] = np.ascontiguousarray(x)
        ptrs = (C.c_void_p * len(bufs))(*[b.ctypes.data for b in bufs])
        rc = self.lib.tb200_oracle_run(T, len(g.tensors), L, len(g.layers), ptrs, uint8_mode)
        if rc != 0:
            raise RuntimeError(f"oracle failed at layer {-rc - 1}")
        return bufs


class Reference:
    """The unmodified reference (OAID/Tengine CPU device, or any device registered in the loaded library)."""

    def __init__(self, libdir=None):
        libdir = libdir or os.path.join(HERE, "_ref")
        shim = os.path.join(libdir, "libref_shim.so")
        if not os.path.exists(shim):
            raise FileNotFoundError(f"{shim} missing: run `make -C oracle` where /root/reference exists")
        C.CDLL(os.path.join(libdir, "libtengine-lite.so"), mode=C.RTLD_GLOBAL)
        self.lib = C.CDLL(shim)
        self.lib.ref_shim_run.restype = C.c_int
        self.lib.ref_shim_version.restype = C.c_char_p

    @staticmethod
    def available():
        return os.path.exists(os.path.join(HERE, "_ref", "libref_shim.so"))

    def version(self):
        return self.lib.ref_shim_version().decode()

    def run(self, g, inputs, want=None, device=None, threads=None, warmup=0, loops=1, env=None):
        """Returns ({tensor_id: array}, (min_ms, avg_ms)).  `want` defaults to g.outputs.
        env e.g. {"TG_DEBUG_REF": "1"} (cpu_module.c:158-166: force the naive reference kernels)."""
        want = list(g.outputs if want is None else want)
        T, L = g.c_tables()
        ins = [np.ascontiguousarray(x) for x in inputs]
        outs = [np.zeros(g.dims(t), dtype=g.np_dtype) for t in want]
        inp = (C.c_void_p * len(ins))(*[a.ctypes.data for a in ins])
        outp = (C.c_void_p * len(outs))(*[a.ctypes.data for a in outs])
        stats = (C.c_double * 4)()
        saved = {}
        for k, v in (env or {}).items():
            saved[k] = os.environ.get(k)
            os.environ[k] = v
        try:
            rc = self.lib.ref_shim_run(T, len(g.tensors), L, len(g.layers), g.id_array(g.inputs), len(g.inputs),
                                       g.id_array(want), len(want), inp, outp,
                                       device.encode() if device else None, threads or (os.cpu_count() or 1),
                                       warmup, loops, stats)
        finally:
            for k, v in saved.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v
        if rc != 0:
            raise RuntimeError(f"reference run failed rc={rc}")
        self.last_stats = tuple(stats)  # (min ms, avg ms, completed runs, elapsed ms)
        return dict(zip(want, outs)), (stats[0], stats[1])


def run_tmfile(ref, path, x, out_shapes, device=None, precision=3, threads=8, warmup=0, loops=1, num_gpus=0, env=None):
    """Run a tmfile through the library's serializer + run_graph on `device` (tm_benchmark's call sequence).  x: NCHW batch;
    out_shapes: shapes of the graph outputs for this batch.  precision: TENGINE_MODE_UINT8 = 3, INT8 = 4.
    num_gpus > 0: pass a tb200_device_option blob through set_context_device.  Returns (outputs, (min_ms, avg_ms))."""
    x = np.ascontiguousarray(x)
    outs = [np.zeros(s, dtype=x.dtype) for s in out_shapes]
    nbytes = (C.c_int64 * len(outs))(*[o.nbytes for o in outs])
    outp = (C.c_void_p * len(outs))(*[o.ctypes.data for o in outs])
    dims = (C.c_int * 4)(*x.shape)
    stats = (C.c_double * 2)()

    class Opt(C.Structure):
        _fields_ = [("dev_name", C.c_char_p), ("num_gpus", C.c_int32), ("first_gpu", C.c_int32)]

    blob, blob_size = None, 0
    if num_gpus and device:
        blob = Opt(device.encode(), int(num_gpus), 0)
        blob_size = C.sizeof(Opt)
    saved = {}
    for k, v in (env or {}).items():
        saved[k] = os.environ.get(k)
        os.environ[k] = v
    try:
        ref.lib.ref_shim_run_tmfile.restype = C.c_int
        rc = ref.lib.ref_shim_run_tmfile(path.encode(), device.encode() if device else None, C.byref(blob) if blob else None, blob_size,
                                         int(precision), dims, C.c_void_p(x.ctypes.data), len(outs), outp, nbytes, int(threads), int(warmup),
                                         int(loops), stats)
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    if rc != 0:
        raise RuntimeError(f"run_tmfile failed rc={rc}")
    return outs, (stats[0], stats[1])


def conv_f32(ref, x, w, bias, stride=1, pad=0, group=1, activation=-1, threads=8, env=None):
    """One fp32 convolution on the reference CPU device (Winograd F(4,3) where winograd_support() admits it, im2col + sgemm
    or the fp32 depthwise kernels otherwise).  x [n,c,h,w], w [oc,c/group,kh,kw] float32.  env e.g. {"TG_DEBUG_REF": "1"}."""
    x, w = np.ascontiguousarray(x, np.float32), np.ascontiguousarray(w, np.float32)
    n, c, h, wd = x.shape
    oc, _, kh, kw = w.shape
    oh, ow = (h + 2 * pad - kh) // stride + 1, (wd + 2 * pad - kw) // stride + 1
    y = np.zeros((n, oc, oh, ow), np.float32)
    b = None if bias is None else np.ascontiguousarray(bias, np.float32)
    saved = {}
    for k, v in (env or {}).items():
        saved[k] = os.environ.get(k)
        os.environ[k] = v
    try:
        ref.lib.ref_shim_conv_f32.restype = C.c_int
        rc = ref.lib.ref_shim_conv_f32(n, c, h, wd, oc, kh, kw, int(stride), int(pad), int(group), int(activation), C.c_void_p(x.ctypes.data),
                                       C.c_void_p(w.ctypes.data), C.c_void_p(b.ctypes.data) if b is not None else None, C.c_void_p(y.ctypes.data),
                                       int(threads))
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    if rc != 0:
        raise RuntimeError(f"reference fp32 conv failed rc={rc}")
    return y


def save_tmfile(ref, g, path):
    """Write GraphDef `g` as a Tengine tmfile with the reference's own writer (tools/save_graph/save_graph.cpp)."""
    T, L = g.c_tables()
    ref.lib.ref_shim_save_tmfile.restype = C.c_int
    rc = ref.lib.ref_shim_save_tmfile(T, len(g.tensors), L, len(g.layers), g.id_array(g.inputs), len(g.inputs),
                                      g.id_array(g.outputs), len(g.outputs), path.encode())
    if rc != 0:
        raise RuntimeError(f"save_tmfile failed rc={rc}")
