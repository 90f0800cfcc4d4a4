"""bench.py's product arm executed end to end WITHOUT a GPU: the CUDA library and torch.cuda are replaced by stand-ins that
follow the real interfaces (tengine_b200.runtime.Context / Graph), so that the control flow the driver launches -- single GPU and
the never-otherwise-exercised N > 1 path (shard verification, per-GPU events, leader barriers) -- cannot hide a Python error
until it runs on a multi-GPU box.  Nothing here measures anything; the assertions are about the JSON contract."""
import contextlib
import json
import os
import sys
import types

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class FakeContext:
    def __init__(self, device=0, devices=None):
        self.devices = [int(device)] if devices is None else [int(d) for d in devices]
        self.device = self.devices[0]
        self.broadcast_kind = "nccl" if len(self.devices) > 1 else "none"

    def stream_of(self, i):
        return 1000 + i

    def probe_int8_tops(self):
        return 4500.0

    def close(self):
        pass


class FakeGraph:
    """Every output element of image i is a function of image i alone (like the real graphs: images are independent units)."""

    def __init__(self, ctx, gdef, flags=0):
        from tengine_b200 import runtime as real

        self.ctx, self.g = ctx, gdef
        n = gdef.dims(gdef.inputs[0])[0]
        r = len(ctx.devices)
        self._shards = [(ctx.devices[k],) + tuple(real.shard_range(n, r, k)) for k in range(r)]

    def shards(self):
        return list(self._shards)

    def run(self, inputs, outputs=None):
        x = inputs[0]
        key = x.reshape(x.shape[0], -1).astype(np.int64).sum(axis=1)
        outs = outputs if outputs is not None else [np.empty(self.g.dims(o), self.g.np_dtype) for o in self.g.outputs]
        for o in outs:
            o[...] = (key % 251).astype(o.dtype).reshape((-1,) + (1,) * (o.ndim - 1))
        return outs

    def upload(self, i, x):
        pass

    def launch(self):
        pass

    def sync(self):
        pass

    def layer_kernels(self):
        names = ["conv_stem_nchw_tcgen05", "conv_dw3x3_tma_dp4a", "gemm_i8_tcgen05", "pool"]
        return [names[min(i, 3) if i < 2 else (1 + i % 2 if i < len(self.g.layers) - 2 else 3)] for i in range(len(self.g.layers))]

    def num_launches(self):
        return len(self.g.layers) + 1

    def profile(self):
        return [0.01 + 0.001 * i for i in range(len(self.g.layers))]

    def work(self):
        return 291.2e9, 2.612e9

    def arena_bytes(self):
        return 600 << 20, 1500 << 20, 4 << 20

    def close(self):
        pass


class FakeEvent:
    clock = [0.0]

    def __init__(self, enable_timing=False):
        self.t = None

    def record(self, stream=None):
        FakeEvent.clock[0] += 0.5
        self.t = FakeEvent.clock[0]

    def elapsed_time(self, other):
        return other.t - self.t


@pytest.mark.parametrize("ngpu,extra", [(1, []), (2, []), (4, ["--workload", "yolov3_tiny_uint8", "--global-batch", "16"])])
def test_product_arm_control_flow_and_json_contract(monkeypatch, capsys, ngpu, extra):
    import torch

    import bench
    from tengine_b200 import runtime as real

    fake = types.SimpleNamespace(Context=FakeContext, Graph=FakeGraph, PinnedBuffer=None, shard_range=real.shard_range)
    monkeypatch.setitem(sys.modules, "tengine_b200.runtime", fake)
    import tengine_b200

    monkeypatch.setattr(tengine_b200, "runtime", fake, raising=False)
    monkeypatch.setattr(torch.cuda, "set_device", lambda d: None)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda d=None: None)
    monkeypatch.setattr(torch.cuda, "is_available", lambda: False)
    monkeypatch.setattr(torch.cuda, "ExternalStream", lambda ptr, device=None: object())
    monkeypatch.setattr(torch.cuda, "Event", FakeEvent)
    monkeypatch.setattr(torch.cuda, "device", lambda d: contextlib.nullcontext())
    monkeypatch.setattr(torch.cuda, "stream", lambda s: contextlib.nullcontext())
    real_empty = torch.empty
    monkeypatch.setattr(torch, "empty", lambda *a, **k: real_empty(*(16,), dtype=k.get("dtype", torch.uint8)))
    monkeypatch.setattr(sys, "argv", ["bench.py", "--gpus", str(ngpu), "--steps", "3", "--warmup", "1", "--cpu-window", "0", "--watchdog", "0"] + extra)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        monkeypatch.delenv(k, raising=False)
    bench.main()
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
                "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert key in line, key
    assert line["n_gpus"] == ngpu and line["steps"] == 3 and line["unit"] == "images/s"
    assert line["scaling"] == ("strong" if extra else "weak")
    assert set(line["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} and line["e2e"]["h2d_bytes_per_step"] > 0
    assert set(line["roofline"]) >= {"bound", "achieved", "peak", "unit", "frac", "traffic"}
    if ngpu > 1:
        assert "shards checked" in line["config"]["verified"] and str(ngpu) in line["config"]["parallelism"]
