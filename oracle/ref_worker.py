"""One reference CPU process for bench.py's `--impl reference` / `cpu_baseline` legs.  TEST / MEASUREMENT INFRASTRUCTURE.

    python -m oracle.ref_worker <graph.npz> <images> <threads>

Runs batch-1 run_graph() calls of the UNMODIFIED reference (oracle/_ref) on `threads` OpenMP threads -- `images` of them,
or, when the parent sets REF_SHIM_WINDOW="<start ms> <end ms>" (fleet mode), as many as complete inside that common
wall-clock window -- and prints one JSON line {"images", "loop_s", "min_ms", "avg_ms"}.  The parent sets OMP_NUM_THREADS (read by libgomp at load time; the
reference derives its core count from omp_get_max_threads(), source/system/cpu.c:108-110) and REF_SHIM_CPUS (the CPU
list this process is pinned to after prerun, see ref_shim.c).  No torch import: the quantised graph arrives as a file.
"""
import json
import sys

import numpy as np


def main(argv):
    graph_npz, images, threads = argv[0], int(argv[1]), int(argv[2])
    from oracle.pyoracle import Reference
    from tengine_b200.graphdef import GraphDef

    d = dict(np.load(graph_npz))
    g = GraphDef.from_dict(d)
    x = d["input"]
    ref = Reference()
    _, (mn, avg) = ref.run(g, [x], threads=threads, warmup=1, loops=images)
    done = int(ref.last_stats[2])  # fleet mode (REF_SHIM_WINDOW): runs completed inside the common window
    print(json.dumps({"images": done, "loop_s": ref.last_stats[3] / 1000.0, "min_ms": mn, "avg_ms": avg}), flush=True)


if __name__ == "__main__":
    main(sys.argv[1:])
