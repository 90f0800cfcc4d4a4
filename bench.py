#!/usr/bin/env python3
"""bench.py -- images/sec of the B200 int8 convolution + GEMM backend on BASELINE.json's headline workload.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload mobilenet_v1_int8] [--batch B]

One "step" = one pass of the hot path (the whole MobileNet-v1 int8 224x224 graph: 28 convolutions + global pool,
through tb200_graph_*) over one synthetic batch of 256 images per GPU.
  value   whole-job images/s with the batch already resident in HBM (device timing, CUDA events, max over ranks)
  e2e     the same through the reference-facing call tb200_graph_run() with HOST buffers (pinned), H2D + D2H inside
  roofline  dominant kernel: algorithmic bytes / its event-timed duration vs the measured HBM peak
  cpu_baseline / --impl reference : the UNMODIFIED reference CPU backend (oracle/_ref) on this box's host cores
Multi-GPU (torchrun, one rank per GPU): the batch dimension is sharded, 256 images per rank (weak scaling); rank 0
packs the weights and ONE NCCL broadcast fills the other ranks' weight arenas at prerun; no collective afterwards.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "images/sec int8 CNN inference (MobileNet-v1 224x224)"
UNIT = "images/s"
# name -> (builder, data type, resolution, default images per GPU, metric label)
WORKLOADS = {
    "mobilenet_v1_int8": ("mobilenet_v1", "int8", 224, 256, "images/sec int8 CNN inference (MobileNet-v1 224x224)"),
    "mobilenet_v1_uint8": ("mobilenet_v1", "uint8", 224, 256, "images/sec uint8 CNN inference (MobileNet-v1 224x224)"),
    "resnet50_uint8": ("resnet50", "uint8", 224, 512, "images/sec uint8 CNN inference (ResNet-50 224x224)"),
    "resnet50_int8": ("resnet50", "int8", 224, 512, "images/sec int8 CNN inference (ResNet-50 224x224)"),
    "yolov3_tiny_uint8": ("yolov3_tiny", "uint8", 416, 16, "images/sec uint8 CNN inference (YOLOv3-tiny 416x416)"),
    "yolov3_tiny_int8": ("yolov3_tiny", "int8", 416, 16, "images/sec int8 CNN inference (YOLOv3-tiny 416x416)"),
}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        super().__init__(daemon=True)
        self.gpu, self.rows, self.stop_flag = gpu, [], False

    def run(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                       "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.p.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
                if self.stop_flag:
                    break
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        try:
            self.p.terminate()
        except Exception:
            pass
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 8:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        busy = sorted(sm)[len(sm) // 2:] if sm else []
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def ncu_traffic(workload, batch, family):
    """dram__bytes_read.sum + dram__bytes_write.sum of the kernel family, summed over its launches of one step, from the
    committed `ncu --set full` capture (profiles/r01_ncu_traffic.json; not measured live: ncu replays every kernel ~40x).
    None when the capture is of another workload / batch."""
    p = os.path.join(ROOT, "profiles", "r01_ncu_traffic.json")
    if not os.path.exists(p):
        return None
    d = json.load(open(p))
    if d.get("workload") != workload or d.get("batch") != batch:
        return None
    f = d["families"].get(family)
    return float(f["dram_bytes_per_step"]) if f else None


def build_workload(name, batch):
    from tengine_b200 import abi, workloads

    if name not in WORKLOADS:
        raise SystemExit(f"unknown workload {name}; choose from {sorted(WORKLOADS)}")
    builder, dt, res, _, _ = WORKLOADS[name]
    return getattr(workloads, builder)(abi.DT_INT8 if dt == "int8" else abi.DT_UINT8, batch=batch, res=res)


def reference_cpu_rate(workload, images, threads_per_proc=int(os.environ.get("TB200_REF_THREADS", "1")), budget_s=150.0):
    """images/s of the UNMODIFIED reference CPU backend (oracle/_ref) using every host core this process may run on:
    P = cores/T independent processes x T OpenMP threads (T = TB200_REF_THREADS, default 1), each pinned to its own
    CPUs and looping batch-1 run_graph() -- the reference's best case for THROUGHPUT: its HCL kernels scale poorly with
    threads on these layer sizes (measured on 8 cores, MobileNet-v1 int8: 8x1 threads 60.7 img/s, 4x2 58.1, 2x4 42.1,
    1x8 22.5), its batched int8 path is slower and, for 3x3, wrong, and its cluster mask cannot
    describe more than 64 CPUs (source/system/cpu.c:120-121,269).  Throughput = images / wall time of the slowest worker's timed loop.
    Workers are plain subprocesses (python -m oracle.ref_worker) with a hard time budget; a worker that does not report
    in time is killed by PID and the measurement fails loudly."""
    import subprocess
    import tempfile

    cpus = sorted(os.sched_getaffinity(0))
    t = max(1, min(threads_per_proc, len(cpus)))
    procs = max(1, len(cpus) // t)
    per = max(4, images // procs)
    g, b = build_workload(workload, 1)
    d = g.to_dict()
    d["input"] = b.random_input(1)
    tmp = tempfile.NamedTemporaryFile(suffix=".npz", delete=False)
    tmp.close()
    np.savez(tmp.name, **d)
    ps = []
    try:
        for i in range(procs):
            env = dict(os.environ)
            env["OMP_NUM_THREADS"] = str(t)
            env["REF_SHIM_CPUS"] = ",".join(str(c) for c in cpus[i * t:(i + 1) * t])
            env["PYTHONPATH"] = ROOT + os.pathsep + env.get("PYTHONPATH", "")
            ps.append(subprocess.Popen([sys.executable, "-m", "oracle.ref_worker", tmp.name, str(per), str(t)], cwd=ROOT,
                                       env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
        deadline = time.time() + budget_s
        res = []
        for p in ps:
            try:
                out, err = p.communicate(timeout=max(1.0, deadline - time.time()))
            except subprocess.TimeoutExpired:
                raise RuntimeError(f"reference worker pid {p.pid} did not finish within {budget_s:.0f} s")
            if p.returncode != 0:
                raise RuntimeError(f"reference worker failed rc={p.returncode}: {err[-500:]}")
            res.append(json.loads(out.strip().splitlines()[-1]))
    finally:
        for p in ps:
            if p.poll() is None:
                p.kill()
                p.wait()
        os.unlink(tmp.name)
    loop_s = max(r["loop_s"] for r in res)
    rate = per * procs / loop_s
    sample = (f"{procs} processes x {t} pinned threads, {per} batch-1 run_graph() each of {workload} "
              f"(slowest loop {loop_s:.2f} s, best single-image latency {min(r['min_ms'] for r in res):.1f} ms)")
    return rate, procs * t, sample


def run_reference_arm(args, rank):
    if rank != 0:
        return
    t0 = time.time()
    per_step = max(2, args.ref_images)
    rates = []
    for s in range(args.warmup + args.steps):
        r, cores, sample = reference_cpu_rate(args.workload, per_step)
        if s >= args.warmup:
            rates.append(r)
        if time.time() - t0 > 150 and rates:
            break
    v = float(np.mean(rates))
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": len(rates),
            "warmup": args.warmup, "ms_per_step": 1000.0 * per_step / v, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": WORKLOADS[args.workload][1], "data": "synthetic",
            "config": {"workload": f"{args.workload}, reference CPU backend, {per_step} images per step (batch-1 runs)"},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "reference", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--workload", default="mobilenet_v1_int8")
    ap.add_argument("--batch", type=int, default=0, help="images per GPU per step (0: the workload's BASELINE.json batch)")
    ap.add_argument("--ref-images", type=int, default=256, help="reference arm: images per step (split over the worker processes)")
    ap.add_argument("--no-tensorcore", action="store_true", help="route convs through the CUDA-core cross-check kernels")
    ap.add_argument("--cpu-images", type=int, default=2048, help="cpu_baseline sample size (images over all workers); 0 disables")
    args = ap.parse_args()

    if args.workload not in WORKLOADS:
        raise SystemExit(f"unknown workload {args.workload}; choose from {sorted(WORKLOADS)}")
    if args.batch <= 0:
        args.batch = WORKLOADS[args.workload][3]
    global METRIC
    METRIC = WORKLOADS[args.workload][4]
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference_arm(args, rank)
        return

    import torch
    import torch.distributed as dist
    from tengine_b200 import abi, sharding
    from tengine_b200 import runtime as rt

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    g, b = build_workload(args.workload, args.batch)
    ctx = rt.Context(local_rank)
    # prerun: rank 0 packs the weights; the others allocate the arena and receive it by ONE NCCL broadcast
    graph = rt.Graph(ctx, g, (abi.PRERUN_DEFAULT if rank == 0 else abi.PRERUN_NO_WEIGHTS) | (abi.PRERUN_NO_TENSORCORE if args.no_tensorcore else 0))
    if world > 1:
        ptr, nbytes = graph.weight_arena()

        class _Arena:
            __cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}

        arena = torch.as_tensor(_Arena(), device=f"cuda:{local_rank}")
        sharding.broadcast_arena(arena, src=0)  # the ONLY collective on the data path (prerun, not steady state)
        torch.cuda.synchronize()

    stream = torch.cuda.ExternalStream(ctx.stream, device=local_rank)
    x = rt.PinnedBuffer(g.dims(g.inputs[0]), g.np_dtype)
    x.array[...] = b.random_input(42 + rank)
    ys = [rt.PinnedBuffer(g.dims(o), g.np_dtype) for o in g.outputs]  # every graph output comes back to the host
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{local_rank}")  # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident: `value` ----------------
    graph.upload(0, x.array)
    graph.sync()
    for _ in range(args.warmup):
        graph.launch()
    graph.sync()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    wall0 = time.time()
    with torch.cuda.stream(stream):
        for s in range(args.steps):
            flush.zero_()  # L2 flush between timed iterations (outside the event pair)
            ev[s][0].record(stream)
            graph.launch()
            ev[s][1].record(stream)
    barrier()
    wall_dev = time.time() - wall0
    dev_ms = sum(a.elapsed_time(b_) for a, b_ in ev)
    clocks = sampler.finish() if sampler else None

    # ---------------- end to end through the reference-facing call with HOST buffers: `e2e` ----------------
    for _ in range(2):
        graph.run([x.array], [y.array for y in ys])
    barrier()
    t0 = time.perf_counter()
    for s in range(args.steps):
        graph.run([x.array], [y.array for y in ys])
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0

    # ---------------- per-kernel profile (events around every launch) for the roofline ----------------
    prof = np.zeros(len(g.layers))
    nprof = min(args.steps, 5)
    for _ in range(nprof):
        prof += np.array(graph.profile())
    prof /= nprof
    kernels = graph.layer_kernels()

    dev_ms, e2e_ms = sharding.max_over_ranks([dev_ms, e2e_s * 1000.0], device=f"cuda:{local_rank}")

    if rank == 0:
        images = args.batch * world * args.steps
        value = images / (dev_ms / 1000.0)
        e2e_value = images / (e2e_ms / 1000.0)
        # dominant kernel family by share of the step
        fam = {}
        for li, k in enumerate(kernels):
            fam.setdefault(k, []).append(li)
        dom = max(fam, key=lambda k: prof[fam[k]].sum())
        dom_ms = float(prof[fam[dom]].sum())
        dom_bytes = 0.0
        for li in fam[dom]:
            L = g.layers[li]
            dom_bytes += g.numel(L["inputs"][0]) + g.numel(L["output"])
            if L["weight"] is not None:
                dom_bytes += L["weight"].size + (4 * g.dims(L["output"])[1] if L["bias"] is not None else 0)
        peak, peak_src = measured_peaks()
        achieved = dom_bytes / (dom_ms / 1000.0) / 1e9
        ops, byts = graph.work()
        step_ms = dev_ms / args.steps
        cpu = None
        if args.cpu_images > 0 and world == 1:
            try:
                v, cores, sample = reference_cpu_rate(args.workload, args.cpu_images)
                cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "reference", "sample": sample}
            except Exception as e:  # oracle/_ref absent
                cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": f"unavailable: {e}"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": WORKLOADS[args.workload][1],
            "data": "synthetic",
            "config": {"workload": f"{args.workload} {WORKLOADS[args.workload][2]}x{WORKLOADS[args.workload][2]} batch={args.batch} per GPU" + (" (BASELINE.json configs[1])" if args.workload == "mobilenet_v1_int8" else ""),
                       "global_batch": args.batch * world, "parallelism": f"batch-sharded x{world}, weights broadcast once (NCCL) at prerun",
                       "l2": "256 MiB L2 flush between timed iterations; per-step activations >> 126 MB L2",
                       "layout": "NHWC int8 in HBM, channels padded to 16"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(x.nbytes), "d2h_bytes_per_step": int(sum(y.nbytes for y in ys)),
                    "ms_per_step": e2e_ms / args.steps, "api": "tb200_graph_run(host NCHW in, host NCHW out), pinned buffers"},
            "gpu_launches": graph.num_launches() * args.steps,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": dom, "launches_per_step": len(fam[dom]), "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_traffic(args.workload, args.batch, dom), "peak_source": peak_src,
                         "kernel_ms_per_step": dom_ms, "share_of_step": dom_ms / float(prof.sum()),
                         "algorithmic_bytes_per_step": dom_bytes},
            "whole_graph": {"algorithmic_gop_per_step": ops / 1e9, "algorithmic_gb_per_step": byts / 1e9,
                            "achieved_tops": ops / (step_ms / 1000.0) / 1e12, "achieved_gbs": byts / (step_ms / 1000.0) / 1e9,
                            "hbm_frac": byts / (step_ms / 1000.0) / 1e9 / peak,
                            "kernel_ms": {k: float(prof[v].sum()) for k, v in fam.items()}},
            "cpu_baseline": cpu,
            "wall_s_device_region": wall_dev,
        }
        print(json.dumps(line), flush=True)
    graph.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
