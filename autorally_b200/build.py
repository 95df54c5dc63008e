"""In-tree build of libmppi_b200.so (sm_100a only) and of the C++ host-layer test driver.

``nvcc`` cross-compiles without a GPU, so this runs in the CPU container; the built ``.so`` travels
to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
CSRC = os.path.join(_HERE, "csrc")
LIB_DIR = os.path.join(_HERE, "lib")
LIB = os.path.join(LIB_DIR, "libmppi_b200.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC"]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found")
    return exe


def _newest(paths):
    return max(os.path.getmtime(p) for p in paths)


def _sources():
    srcs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh"))]
    srcs.append(os.path.join(ROOT, "include", "mppi_b200.h"))
    return srcs


def build_library(force=False, verbose=False):
    os.makedirs(LIB_DIR, exist_ok=True)
    srcs = _sources()
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= _newest(srcs):
        return LIB
    cus = [s for s in srcs if s.endswith(".cu")]
    objs = []
    procs = []
    for cu in cus:
        obj = os.path.join(LIB_DIR, os.path.basename(cu)[:-3] + ".o")
        objs.append(obj)
        cmd = [_nvcc()] + NVCC_FLAGS + ["-I", os.path.join(ROOT, "include"), "-c", cu, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose and out:
            print(out)
        if p.returncode != 0:
            raise RuntimeError("nvcc failed: %s\n%s" % (" ".join(cmd), out))
    link = [_nvcc(), "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-ldl"]
    subprocess.check_call(link)
    return LIB


def build_host_tests(force=False):
    """Compiles the C++ host-layer test drivers (include/autorally_control/path_integral/*)."""
    outs = [_build_driver(name, force) for name in ("host_api_driver", "control_loop_driver")]
    return outs[0]


def _build_driver(name, force=False):
    src = os.path.join(ROOT, "tests", "cpp", name + ".cu")
    if not os.path.exists(src):
        return None
    out = os.path.join(LIB_DIR, name)
    deps = [src, LIB]
    inc = os.path.join(ROOT, "include")
    for d, _, fs in os.walk(inc):
        deps += [os.path.join(d, f) for f in fs]
    if not force and os.path.exists(out) and os.path.getmtime(out) >= _newest(deps):
        return out
    cmd = [_nvcc()] + NVCC_FLAGS + ["-I", inc, "-I", os.path.join(inc, "compat"), src, "-o", out, "-L", LIB_DIR, "-lmppi_b200",
                                    "-Xlinker", "-rpath", "-Xlinker", "$ORIGIN"]
    subprocess.check_call(cmd)
    return out


if __name__ == "__main__":
    import sys
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
