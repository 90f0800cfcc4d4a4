#!/usr/bin/env python3
"""Build the UNMODIFIED reference (OAID/Tengine, CPU device only) into oracle/_ref/.

TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is on the product path; the
outputs are used (a) to pin oracle/tb200_oracle.c against the real reference,
(b) as the CPU baseline of bench.py (`--impl reference`, cpu_baseline.kind
"reference").

What this does (and does not do):
  * It does NOT run the reference's CMake build system.  It compiles the C
    sources where they lie under /root/reference with /usr/bin/gcc (-O3 -mfma
    -mf16c -fopenmp, the flags source/device/cpu/CMakeLists.txt:268-271 and
    cmake/libraries/openmp.cmake use on x86) and links them into
    oracle/_ref/libtengine-lite.so.
  * The five tiny registry headers that CMake would generate
    (cmake/registry.cmake:2-40 applied to source/device/register.h.in,
    source/device/cpu/cpu_ops.h.in, source/operator/prototype.h.in,
    source/serializer/register.h.in, source/serializer/tmfile/tm2_ops.h.in) and
    defines.h are expanded here from the templates, at build time, into
    oracle/_ref/gen/ -- they are lists of `extern int register_xxx();` calls.
  * Symbols are left visible (the reference's TENGINE_ENABLE_ALL_SYMBOL=ON
    behaviour, source/CMakeLists.txt:324-327) so that tests can call internal
    kernels (ref_conv_int8, ...) and an out-of-tree device can call
    register_device().
  * No reference SOURCE is copied into this repository: outputs go only to
    oracle/_ref/ (git-ignored, but shipped to the GPU box by gpurun).

Usage: python oracle/build_ref.py [--ref /root/reference] [--jobs 8] [--extra-device SRC...]
"""
import argparse
import concurrent.futures as cf
import glob
import hashlib
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
CC = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"

CFLAGS = ["-O3", "-mfma", "-mf16c", "-fopenmp", "-fPIC", "-fdata-sections", "-ffunction-sections",
          "-w", "-std=gnu99"]


def registry_text(lead_reg, lead_del, back, files):
    """Python restatement of cmake/registry.cmake:2-40 (GENERATE_REGISTER_HEADER_FILE)."""
    bgn, end = "// code generation start\n", "// code generation finish\n"
    reg_def = reg_cal = del_def = del_cal = bgn
    for f in files:
        name = os.path.splitext(os.path.basename(f))[0]
        rf = f"{lead_reg}{name}{back}()"
        df = f"{lead_del}{name}{back}()"
        reg_def += f"extern int {rf};\n"
        reg_cal += (f"    ret = {rf};\n    if(0 != ret)\n    {{\n"
                    f"        TLOG_ERR(\"Tengine FATAL: Call %s failed(%d).\\n\", \"{rf}\", ret);\n    }}\n")
        del_def += f"extern int {df};\n"
        del_cal += (f"    ret = {df};\n    if(0 != ret)\n    {{\n"
                    f"        TLOG_ERR(\"Tengine FATAL: Call %s failed(%d).\\n\", \"{rf}\", ret);\n    }}\n")
    return {"_GEN_REG_DEF_STR": reg_def + end, "_GEN_REG_CAL_STR": reg_cal + "    " + end,
            "_GEN_DEL_DEF_STR": del_def + end, "_GEN_DEL_CAL_STR": del_cal + "    " + end}


def configure(template, target, subst):
    txt = open(template).read()
    for k, v in subst.items():
        txt = txt.replace("@" + k + "@", v)
    os.makedirs(os.path.dirname(target), exist_ok=True)
    if not os.path.exists(target) or open(target).read() != txt:
        open(target, "w").write(txt)


def source_lists(ref):
    S = os.path.join(ref, "source")
    g = lambda p: sorted(glob.glob(os.path.join(S, p)))
    ops = sorted(d for d in os.listdir(os.path.join(S, "device/cpu/op"))
                 if os.path.isdir(os.path.join(S, "device/cpu/op", d)))
    cpu_ref, cpu_x86, cpu_reg = [], [], []
    for op in ops:
        cpu_ref += g(f"device/cpu/op/{op}/*.c")
        cpu_x86 += g(f"device/cpu/op/{op}/x86/*.c")
        cpu_reg += g(f"device/cpu/op/{op}/{op}_ref.c")
        cpu_reg += g(f"device/cpu/op/{op}/x86/*_hcl_x86.c")
    proto = g("operator/prototype/*.c")
    srl_tm2 = g("serializer/tmfile/*.c")
    srl_ops = g("serializer/tmfile/op/*.c")
    srcs = (g("api/*.c") + g("executer/*.c") + g("graph/*.c") + g("module/*.c") + g("optimizer/*.c")
            + g("system/*.c") + g("utility/*.c") + g("scheduler/*.c") + g("operator/*.c") + proto
            + g("serializer/*.c") + srl_tm2 + srl_ops + g("device/*.c") + g("device/cpu/*.c")
            + cpu_ref + cpu_x86)
    return dict(srcs=srcs, cpu_reg=cpu_reg, proto=proto, srl_tm2=srl_tm2, srl_ops=srl_ops)


def gen_headers(ref, gen, lists, device_files):
    S = os.path.join(ref, "source")
    defines = open(os.path.join(S, "defines.h.in")).read()
    defines = defines.replace("#cmakedefine TENGINE_HAS_LIB_POSIX_THREAD", "#define TENGINE_HAS_LIB_POSIX_THREAD")
    defines = defines.replace("#cmakedefine TENGINE_HAS_INC_SYSLOG", "/* #undef TENGINE_HAS_INC_SYSLOG */")
    defines = defines.replace("#cmakedefine TENGINE_ENABLE_ENV_VAR", "#define TENGINE_ENABLE_ENV_VAR")
    os.makedirs(os.path.join(gen, "source"), exist_ok=True)
    p = os.path.join(gen, "source/defines.h")
    if not os.path.exists(p) or open(p).read() != defines:
        open(p, "w").write(defines)
    configure(os.path.join(S, "device/register.h.in"), os.path.join(gen, "source/device/register.h"),
              registry_text("register_", "unregister_", "", device_files))
    configure(os.path.join(S, "device/cpu/cpu_ops.h.in"), os.path.join(gen, "source/device/cpu/cpu_ops.h"),
              registry_text("register_", "unregister_", "_op", lists["cpu_reg"]))
    configure(os.path.join(S, "operator/prototype.h.in"), os.path.join(gen, "source/operator/prototype.h"),
              registry_text("register_", "unregister_", "_op", lists["proto"]))
    configure(os.path.join(S, "serializer/register.h.in"), os.path.join(gen, "source/serializer/register.h"),
              registry_text("register_", "unregister_", "", lists["srl_tm2"]))
    configure(os.path.join(S, "serializer/tmfile/tm2_ops.h.in"),
              os.path.join(gen, "source/serializer/tmfile/tm2_ops.h"),
              registry_text("register_", "unregister_", "_op", lists["srl_ops"]))


def include_flags(ref, gen):
    S = os.path.join(ref, "source")
    inc = [S, os.path.join(gen, "source"), os.path.join(S, "device"), os.path.join(gen, "source/device/cpu"),
           os.path.join(S, "device/cpu"), os.path.join(S, "operator/prototype"), os.path.join(S, "serializer"),
           os.path.join(gen, "source/serializer"), os.path.join(gen, "source/device"),
           os.path.join(gen, "source/operator")]
    return [f"-I{d}" for d in inc]


def compile_one(args):
    src, obj, cmd = args
    if os.path.exists(obj) and os.path.getmtime(obj) >= os.path.getmtime(src):
        return None
    os.makedirs(os.path.dirname(obj), exist_ok=True)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        return f"FAILED {src}\n{r.stderr[-2000:]}"
    return None


def obj_path(objdir, ref, src):
    rel = os.path.relpath(src, ref) if src.startswith(ref) else "ext/" + hashlib.md5(src.encode()).hexdigest()[:8] + "_" + os.path.basename(src)
    return os.path.join(objdir, rel + ".o")


def build(ref="/root/reference", jobs=8, verbose=True):
    """Build oracle/_ref/libtengine-lite.so (reference, CPU device only). Returns its path."""
    if not os.path.isdir(os.path.join(ref, "source")):
        raise RuntimeError(f"reference tree not found at {ref}")
    gen = os.path.join(OUT, "gen")
    objdir = os.path.join(OUT, "obj")
    lists = source_lists(ref)
    gen_headers(ref, gen, lists, [os.path.join(ref, "source/device/cpu/cpu_device.c")])
    inc = include_flags(ref, gen)
    tasks = []
    objs = []
    for s in lists["srcs"]:
        o = obj_path(objdir, ref, s)
        objs.append(o)
        tasks.append((s, o, [CC] + CFLAGS + inc + ["-c", s, "-o", o]))
    errs = []
    with cf.ThreadPoolExecutor(max_workers=jobs) as ex:
        for e in ex.map(compile_one, tasks):
            if e:
                errs.append(e)
    if errs:
        raise RuntimeError("\n".join(errs[:5]))
    lib = os.path.join(OUT, "libtengine-lite.so")
    newest = max(os.path.getmtime(o) for o in objs)
    if not os.path.exists(lib) or os.path.getmtime(lib) < newest:
        rsp = os.path.join(OUT, "objs.rsp")
        open(rsp, "w").write("\n".join(objs))
        subprocess.check_call([CC, "-shared", "-fopenmp", "-o", lib, "@" + rsp, "-lm", "-ldl", "-lpthread"])
    if verbose:
        print(f"[oracle] reference library: {lib} ({len(objs)} objects)")
    return lib


def build_apps(ref="/root/reference", verbose=True):
    """Compile the UNMODIFIED example / benchmark programs of the north star against the library."""
    lib = os.path.join(OUT, "libtengine-lite.so")
    # the apps include "tengine/c_api.h" (CMake installs source/api there): expose it through a symlink
    incdir = os.path.join(OUT, "gen/include")
    os.makedirs(incdir, exist_ok=True)
    link = os.path.join(incdir, "tengine")
    if not os.path.islink(link):
        os.symlink(os.path.join(ref, "source/api"), link)
    inc = [f"-I{incdir}", f"-I{ref}/source", f"-I{os.path.join(OUT, 'gen/source')}", f"-I{ref}/examples/common",
           f"-I{ref}/benchmark/common", f"-I{ref}/examples"]
    apps = {
        "tm_classification_int8": ([f"{ref}/examples/tm_classification_int8.c", f"{ref}/examples/common/tengine_operations.c"], CC, ["-std=gnu99"]),
        "tm_classification_uint8": ([f"{ref}/examples/tm_classification_uint8.c", f"{ref}/examples/common/tengine_operations.c"], CC, ["-std=gnu99"]),
        "tm_benchmark": ([f"{ref}/benchmark/tm_benchmark.cc", f"{ref}/benchmark/common/timer.cc"], CXX, ["-std=c++11"]),
    }
    built = []
    for name, (srcs, comp, std) in apps.items():
        out = os.path.join(OUT, name)
        if os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(s) for s in srcs) \
                and os.path.getmtime(out) >= os.path.getmtime(lib):
            built.append(out)
            continue
        cmd = [comp, "-O2", "-w"] + std + inc + srcs + ["-o", out, f"-L{OUT}", "-ltengine-lite",
                                                          "-Wl,-rpath,$ORIGIN", "-lm", "-ldl", "-lpthread", "-fopenmp"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            print(f"[oracle] WARNING: could not build {name}:\n{r.stderr[-1500:]}", file=sys.stderr)
            continue
        built.append(out)
    # the detection post-processing of the unmodified YOLOv3-tiny example, callable (oracle/yolo_example_shim.cpp): its source needs
    # C++ OpenCV, replaced here by the minimal stand-in headers of oracle/cvstub (only cv::Rect_ carries semantics)
    here = os.path.dirname(os.path.abspath(__file__))
    shim, out = os.path.join(here, "yolo_example_shim.cpp"), os.path.join(OUT, "libyolo_example.so")
    example = f"{ref}/examples/tm_yolov3_tiny_uint8.cpp"
    if os.path.exists(shim) and os.path.exists(example):
        if not (os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(s) for s in (shim, example, os.path.join(here, "cvstub/opencv2/core/core.hpp")))):
            ops_o = os.path.join(OUT, "yolo_example_tengine_operations.o")  # helpers the example's main() calls (C source of the reference)
            r = subprocess.run([CC, "-O2", "-w", "-std=gnu99", "-fPIC", "-c"] + inc + [f"{ref}/examples/common/tengine_operations.c", "-o", ops_o], capture_output=True, text=True)
            cmd = [CXX, "-O2", "-w", "-std=c++11", "-fPIC", "-shared", f"-I{os.path.join(here, 'cvstub')}"] + inc + [shim, ops_o, "-o", out, f"-L{OUT}", "-ltengine-lite",
                                                                                                                   "-Wl,-rpath,$ORIGIN", "-lm"]
            if r.returncode == 0:
                r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                print(f"[oracle] WARNING: could not build libyolo_example.so:\n{r.stderr[-1500:]}", file=sys.stderr)
        if os.path.exists(out):
            built.append(out)
    if verbose:
        print("[oracle] apps:", ", ".join(os.path.basename(b) for b in built))
    return built


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--jobs", type=int, default=os.cpu_count() or 8)
    a = ap.parse_args()
    build(a.ref, a.jobs)
    build_apps(a.ref)
