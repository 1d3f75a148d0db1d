"""Build recipe of the oracle's C restatement (oracle/c/vsmpc_ref.c) -> oracle/_build/libvsmpc_ref.so.

TEST INFRASTRUCTURE.  The reference's tick cannot be compiled here (its hot-path translation units
include <OsqpEigen/OsqpEigen.h>, Eigen, iDynTree, YARP and BLF headers, none of which exist in this
image — DESIGN.md "Oracle"; only its scalar jet model builds, oracle/build_ref.py -> oracle/_ref/);
this library is the CPU "port" of the path.
It is compiled with -O3 -march=native on the machine that uses it (rebuilt when the CPU differs).
"""
import hashlib
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "c", "vsmpc_ref.c")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libvsmpc_ref.so")
STAMP = os.path.join(OUT_DIR, "libvsmpc_ref.stamp")
FLAGS = ["-O3", "-march=native", "-fopenmp", "-fPIC", "-shared", "-std=gnu11"]


def _stamp() -> str:
    h = hashlib.sha256(open(SRC, "rb").read())
    h.update(" ".join(FLAGS).encode())
    try:
        cpu = subprocess.run(["gcc", "-march=native", "-Q", "--help=target"], capture_output=True, text=True).stdout
        h.update(cpu.encode())
    except Exception:
        pass
    return h.hexdigest()


def build(force: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    st = _stamp()
    if not force and os.path.exists(LIB) and os.path.exists(STAMP) and open(STAMP).read().strip() == st:
        return LIB
    cmd = ["gcc"] + FLAGS + ["-o", LIB, SRC, "-lm"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("gcc failed:\n" + res.stdout + res.stderr)
    open(STAMP, "w").write(st)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
