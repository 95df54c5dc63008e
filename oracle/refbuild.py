"""Builds oracle/_ref/libautorally_ref.so: the REFERENCE's own MPPI controller (rdesc/autorally), compiled
for sm_100a from the sources where they lie under /root/reference, against stand-in headers for the
libraries this container lacks (oracle/ref_shim/: Eigen, cnpy, ROS/XmlRpc, OpenCV, Boost, DDP).

TEST INFRASTRUCTURE ONLY.  The library is the checker the parity tests compare against and the
"reference" arm of bench.py; the product never loads it.  /root/reference exists only in the build
container: the built .so travels to the GPU box (git-ignored, not gpurun-ignored); no reference source
is copied into this repo.

Not the reference's build system: one nvcc command.  Differences from SRC/CMakeLists.txt:27-37:
-arch=sm_52 -> sm_100a (the only GPU here), no -maxrregcount=32 (a Maxwell occupancy knob; it does not
change arithmetic), same default -fmad=true, no --use_fast_math.
"""
from __future__ import annotations

import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_INCLUDE = "/root/reference/autorally_control/include"
OUT_DIR = os.path.join(_HERE, "_ref")
LIB = os.path.join(OUT_DIR, "libautorally_ref.so")
SRC = os.path.join(_HERE, "ref_harness.cu")
SHIM = os.path.join(_HERE, "ref_shim")


def reference_present():
    return os.path.exists(os.path.join(REF_INCLUDE, "autorally_control", "path_integral", "mppi_controller.cu"))


def _newest_dep():
    deps = [SRC]
    for d, _, fs in os.walk(SHIM):
        deps += [os.path.join(d, f) for f in fs]
    pi = os.path.join(REF_INCLUDE, "autorally_control", "path_integral")
    deps += [os.path.join(pi, f) for f in os.listdir(pi)]
    return max(os.path.getmtime(p) for p in deps)


def build(force=False):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    os.makedirs(OUT_DIR, exist_ok=True)
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= _newest_dep():
        return LIB
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++14", "-w", "-lineinfo",
           "-Xcompiler", "-fPIC", "-shared", "-I", SHIM, "-I", REF_INCLUDE, SRC, "-o", LIB,
           "-lcurand", "-Xlinker", "-rpath=/usr/local/cuda/lib64"]
    subprocess.check_call(cmd)
    return LIB


def build_if_possible(force=False):
    """Builds when /root/reference is present (this container); elsewhere the prebuilt file is used."""
    if reference_present():
        return build(force)
    return LIB if os.path.exists(LIB) else None


if __name__ == "__main__":
    print(build_if_possible(force=True))
