"""Build recipes of oracle/_ref/: the reference's OWN sources, compiled from where they lie under /root/reference.

TEST INFRASTRUCTURE.  The reference's build system (CMake + pixi) and its third-party libraries (Eigen, OsqpEigen / OSQP,
iDynTree, YARP, BLF, matio, boost) are not available here, but the hot path's translation units only need the *headers* they
include to exist.  oracle/ref_stubs/ provides stand-ins for those headers (see each file's first lines), and

* ``build()``      -> oracle/_ref/libjetmodel_ref.so: ``utils/src/JetModel.cpp`` (rows a6 / a17 of SURVEY §8) + oracle/ref_shim.cpp;
* ``build_mpc()``  -> oracle/_ref/libvsmpc_reference.so: the 13 translation units of ``VariableSamplingMPC`` (MPC_SOURCES
  below: rows a1-a15, a17 and the call sequence of a16) + oracle/ref_mpc_shim.cpp.

No reference source is copied: g++ reads the files under /root/reference; the only outputs are the two libraries
(git-ignored, they travel to the GPU box, where /root/reference does not exist and the prebuilt files are used).
Users: tests/test_reference_pinned.py, tests/test_oracle.py, tests/golden/make_reference_golden.py,
tests/golden/make_jet_model_golden.py.
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
REF_UT = "/root/reference/src/flight-controller/utils"
OUT_DIR = os.path.join(HERE, "_ref")
LIB = os.path.join(OUT_DIR, "libjetmodel_ref.so")


def available() -> bool:
    return os.path.exists(os.path.join(REF_UT, "src", "JetModel.cpp"))


def build(force: bool = False) -> str:
    """Compiles when /root/reference is present (the build container); elsewhere returns the prebuilt file or ''."""
    if not available():
        return LIB if os.path.exists(LIB) else ""
    if os.path.exists(LIB) and not force:
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    cmd = ["g++", "-O2", "-fPIC", "-shared", "-std=c++17", "-ffp-contract=off",
           "-I", os.path.join(HERE, "ref_stubs"), "-I", os.path.join(REF_UT, "include"),
           "-o", LIB, os.path.join(REF_UT, "src", "JetModel.cpp"), os.path.join(HERE, "ref_shim.cpp")]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("g++ failed:\n" + res.stdout + res.stderr)
    return LIB


REF_FC = "/root/reference/src/flight-controller"
MPC_LIB = os.path.join(OUT_DIR, "libvsmpc_reference.so")
MPC_SOURCES = ["momentum-based-linear-mpc-lib/src/variableSamplingMPC/variableSamplingMPC.cpp",
               "momentum-based-linear-mpc-lib/src/variableSamplingMPC/systemDynamicsVSMPC.cpp",
               "momentum-based-linear-mpc-lib/src/variableSamplingMPC/constraintsVSMPC.cpp",
               "momentum-based-linear-mpc-lib/src/variableSamplingMPC/costsVSMPC.cpp",
               "momentum-based-linear-mpc-lib/src/IMPCProblem/IMPCProblem.cpp",
               "momentum-based-linear-mpc-lib/src/IMPCProblem/IQPUtilsMPC.cpp",
               "momentum-based-linear-mpc-lib/src/IMPCProblem/systemDynamic.cpp",
               "utils/src/QPInput.cpp", "utils/src/IQPCost.cpp", "utils/src/IQPConstraint.cpp",
               "utils/src/FlightControlUtils.cpp", "utils/src/TrajectoryManager.cpp", "utils/src/JetModel.cpp"]


def build_mpc(force: bool = False) -> str:
    """oracle/_ref/libvsmpc_reference.so: the reference's VariableSamplingMPC (13 of its own translation units, compiled
    where they lie) + oracle/ref_mpc_shim.cpp, against the stand-in headers of oracle/ref_stubs/."""
    if not available():
        return MPC_LIB if os.path.exists(MPC_LIB) else ""
    shim = os.path.join(HERE, "ref_mpc_shim.cpp")
    stubs = os.path.join(HERE, "ref_stubs")
    newest = max([os.path.getmtime(shim)] + [os.path.getmtime(os.path.join(d, f)) for d, _, fs in os.walk(stubs) for f in fs])
    if os.path.exists(MPC_LIB) and not force and os.path.getmtime(MPC_LIB) >= newest:
        return MPC_LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    cmd = ["g++", "-O2", "-fPIC", "-shared", "-std=c++17", "-ffp-contract=off", "-I", stubs,
           "-I", os.path.join(REF_FC, "utils", "include"), "-I", os.path.join(REF_FC, "momentum-based-linear-mpc-lib", "include"),
           "-o", MPC_LIB] + [os.path.join(REF_FC, f) for f in MPC_SOURCES] + [shim]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("g++ failed:\n" + res.stdout + res.stderr)
    return MPC_LIB


GLUE_LIB = os.path.join(OUT_DIR, "libvsmpc_reference_glue.so")


def build_glue(force: bool = False) -> str:
    """oracle/_ref/libvsmpc_reference_glue.so: the same 13 reference translation units + oracle/ref_mpc_shim.cpp compiled with
    -DVSMPC_WITH_GLUE, i.e. together with the PRODUCT's reference-side binding include/vsmpc_reference_glue.hpp (the class a
    maintainer drops into the reference tree: configure(weak_ptr<IParametersHandler>, QPInput&) / update(QPInput&) / solveMPC /
    getters over libvsmpc.so), so that tests/test_reference_glue.py can drive the reference's VariableSamplingMPC and the
    GPU-backed class on ONE QPInput object.  Links libvsmpc.so by relative rpath."""
    root = os.path.dirname(HERE)
    pkg = os.path.join(root, "paper_gorbani_2025_humanoids_multi-rate-mpc-ironcub_b200")
    if not available():
        return GLUE_LIB if os.path.exists(GLUE_LIB) else ""
    shim = os.path.join(HERE, "ref_mpc_shim.cpp")
    stubs = os.path.join(HERE, "ref_stubs")
    inc = os.path.join(root, "include")
    deps = [shim, os.path.join(inc, "vsmpc_reference_glue.hpp"), os.path.join(inc, "vsmpc_adapter.hpp"), os.path.join(inc, "vsmpc.h")]
    newest = max([os.path.getmtime(d) for d in deps] + [os.path.getmtime(os.path.join(d, f)) for d, _, fs in os.walk(stubs) for f in fs])
    if os.path.exists(GLUE_LIB) and not force and os.path.getmtime(GLUE_LIB) >= newest:
        return GLUE_LIB
    if not os.path.exists(os.path.join(pkg, "libvsmpc.so")):
        return ""
    os.makedirs(OUT_DIR, exist_ok=True)
    cmd = ["g++", "-O2", "-fPIC", "-shared", "-std=c++17", "-ffp-contract=off", "-DVSMPC_WITH_GLUE", "-I", stubs, "-I", inc,
           "-I", os.path.join(REF_FC, "utils", "include"), "-I", os.path.join(REF_FC, "momentum-based-linear-mpc-lib", "include"),
           "-o", GLUE_LIB] + [os.path.join(REF_FC, f) for f in MPC_SOURCES] + [shim] \
        + ["-L", pkg, "-lvsmpc", "-Wl,-rpath,$ORIGIN/../../" + os.path.basename(pkg)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("g++ failed:\n" + res.stdout + res.stderr)
    return GLUE_LIB


def load():
    import ctypes
    path = build()
    if not path:
        return None
    lib = ctypes.CDLL(path)
    for f, args in (("ref_jet_poly", [ctypes.c_int, ctypes.c_double, ctypes.c_double]),
                    ("ref_jet_scalar", [ctypes.c_int, ctypes.c_double])):
        getattr(lib, f).argtypes = args
        getattr(lib, f).restype = ctypes.c_double
    return lib


if __name__ == "__main__":
    print(build(force=True))
    print(build_mpc(force=True))
    print(build_glue(force=True))
