"""In-tree build of the CUDA shared library (sm_100a only).  No JIT cache: the .so lives next to
this file so that it travels with the repo snapshot to the GPU box."""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libvsmpc.so")
STAMP = os.path.join(HERE, ".libvsmpc.stamp")

SOURCES = ["vsmpc_api.cu", "vsmpc_linearise.cu", "vsmpc_qp_generic.cu", "vsmpc_qp_structured.cu", "vsmpc_qp_condensed.cu", "vsmpc_qp_condensed_wide.cu", "vsmpc_qp_fallback.cu",
           "vsmpc_plant.cu", "vsmpc_kinematics.cu",
           "vsmpc_microbench.cu"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "--shared", "-Xptxas", "-v"]
if os.environ.get("VSMPC_PHASE_CLOCKS"):      # development build: per-phase clock stamps in the condensed kernel
    NVCC_FLAGS.append("-DVSMPC_PHASE_CLOCKS")
if os.environ.get("VSMPC_FB_CLOCKS"):         # development build: the fallback kernel prints its phase times
    NVCC_FLAGS.append("-DVSMPC_FB_CLOCKS")


def _digest() -> str:
    h = hashlib.sha256()
    files = sorted(os.listdir(CSRC)) + [os.path.join(ROOT, "include", "vsmpc.h")]
    for f in files:
        p = f if os.path.isabs(f) else os.path.join(CSRC, f)
        with open(p, "rb") as fh:
            h.update(f.encode())
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def nvcc_path() -> str:
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def build(force: bool = False, verbose: bool = False) -> str:
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(STAMP):
        with open(STAMP) as f:
            if f.read().strip() == dig:
                return LIB
    cmd = [nvcc_path()] + NVCC_FLAGS + ["-I", os.path.join(ROOT, "include"), "-I", CSRC, "-o", LIB] \
        + [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = res.stdout + res.stderr
    with open(os.path.join(HERE, ".libvsmpc.buildlog"), "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + log[-6000:])
    if verbose:
        print(log)
    with open(STAMP, "w") as f:
        f.write(dig)
    return LIB


def build_examples(force: bool = False) -> str:
    """C++ host examples over include/vsmpc_adapter.hpp (g++ only; link libvsmpc.so by relative rpath): the single-instance
    controller loop, the one-process multi-GPU batch and the end-to-end bench of the C++ host layer (pack + upload + solve).  Returns the path of the first."""
    out_dir = os.path.join(ROOT, "examples", "bin")
    os.makedirs(out_dir, exist_ok=True)
    exes = []
    for name in ("cpp_controller", "cpp_multi_gpu", "cpp_host_bench"):
        src = os.path.join(ROOT, "examples", name + ".cpp")
        exe = os.path.join(out_dir, name)
        exes.append(exe)
        deps = [src, os.path.join(ROOT, "include", "vsmpc_adapter.hpp"), os.path.join(ROOT, "include", "vsmpc.h"), LIB]
        if not force and os.path.exists(exe) and all(os.path.getmtime(exe) >= os.path.getmtime(d) for d in deps):
            continue
        cmd = [shutil.which("g++") or "g++", "-O2", "-std=c++17", "-Wall", "-I", os.path.join(ROOT, "include"), src,
               "-L", HERE, "-lvsmpc", "-pthread", "-Wl,-rpath,$ORIGIN/../../" + os.path.basename(HERE), "-o", exe]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("g++ failed:\n" + res.stdout + res.stderr)
    return exes[0]


if __name__ == "__main__":
    print(build(force=True, verbose=True))
