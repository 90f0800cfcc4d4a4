"""The CPU oracle against (1) the known-answer vectors embedded in the reference's own op tests and (2) fixtures
produced by running the unmodified reference (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from tests.helpers import dequant, kat_graphs, layer_outputs, load_golden


@pytest.mark.parametrize("kat", kat_graphs(), ids=lambda k: k[0])
def test_oracle_reproduces_reference_kats(oracle, kat):
    name, g, xins, expected, tol = kat
    for mode in (0, 1):
        out = oracle.run(g, xins, uint8_mode=mode)[g.outputs[0]]
        real = dequant(g, g.outputs[0], out)
        assert np.abs(real - expected).max() <= tol + 1e-6, (name, mode, real.ravel(), expected.ravel())


@pytest.mark.parametrize("name", ["ref_tiny_int8", "ref_mobilenet025_int8", "ref_resnet50_small_int8", "ref_yolov3_tiny_small_int8"])
def test_oracle_bit_exact_vs_reference_fixture_int8(oracle, name):
    g, x, ref = load_golden(name)
    out = oracle.run(g, [x])
    for t in layer_outputs(g):
        assert np.array_equal(out[t], ref[t]), f"{name}: tensor {t} differs from the reference"


@pytest.mark.parametrize("name", ["ref_tiny_uint8", "ref_mobilenet025_uint8"])
def test_oracle_vs_reference_fixture_uint8(oracle, name):
    """uint8: the reference accumulates in fp32 (conv_kernel_x86.c:1703-1794); the exact-integer restatement
    (mode 0, what the device computes) may differ by 1 LSB per layer; errors then propagate, so compare each layer
    on the REFERENCE's input to that layer."""
    g, x, ref = load_golden(name)
    bufs = {g.inputs[0]: x, **ref}
    from tengine_b200.graphdef import GraphDef

    for li, L in enumerate(g.layers):
        sub = GraphDef(g.data_type)
        sub.tensors = g.tensors
        sub.layers = [L]
        sub.inputs = list(L["inputs"])
        sub.outputs = [L["output"]]
        for mode in (0, 1):
            out = oracle.run(sub, [bufs[i] for i in L["inputs"]], uint8_mode=mode)[L["output"]]
            d = np.abs(out.astype(int) - ref[L["output"]].astype(int))
            assert d.max() <= 1, (name, li, mode, int(d.max()))
            assert (d > 0).mean() < 0.02, (name, li, mode, float((d > 0).mean()))
