"""bench.py with the GPU stand-ins of tests/fake_gpu.py installed: what tests/test_bench_flow.py launches under torchrun."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
import bench  # noqa: E402
import fake_gpu  # noqa: E402

fake_gpu.install()
bench.main()
