#!/usr/bin/env python3
"""Compile oracle/ref_shim.c (+ the bridge to the reference's tmfile writer, tools/save_graph/*.cpp compiled where they
lie) into <libdir>/libref_shim.so, linked against <libdir>/libtengine-lite.so.  TEST INFRASTRUCTURE.
usage: build_shim.py <reference root> <libdir>"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CC = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"


def main(ref, libdir):
    gen = os.path.join(libdir, "gen")
    inc = [f"-I{ref}/source", f"-I{gen}/source", f"-I{gen}/include", f"-I{ref}/source/operator/prototype",
           f"-I{ref}/tools/save_graph", f"-I{ref}/source/serializer/tmfile"]
    objd = os.path.join(libdir, "shim_obj")
    os.makedirs(objd, exist_ok=True)
    objs = []
    jobs = [(os.path.join(HERE, "ref_shim.c"), [CC, "-std=gnu99", "-fopenmp"]),
            (os.path.join(HERE, "ref_shim_save.cpp"), [CXX, "-std=c++11", "-include", "cstdint"]),
            (f"{ref}/tools/save_graph/save_graph.cpp", [CXX, "-std=c++11", "-include", "cstdint"]),
            (f"{ref}/tools/save_graph/tm2_op_save.cpp", [CXX, "-std=c++11", "-include", "cstdint"]),
            (f"{ref}/tools/save_graph/tm2_generate.c", [CC, "-std=gnu99"])]
    for src, comp in jobs:
        o = os.path.join(objd, os.path.basename(src) + ".o")
        subprocess.check_call(comp + ["-O2", "-fPIC", "-w"] + inc + ["-c", src, "-o", o])
        objs.append(o)
    subprocess.check_call([CXX, "-shared", "-o", os.path.join(libdir, "libref_shim.so")] + objs +
                          [f"-L{libdir}", "-ltengine-lite", "-fopenmp", "-Wl,-rpath,$ORIGIN"])


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
