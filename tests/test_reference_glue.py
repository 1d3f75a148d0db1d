"""The product's reference-side binding (include/vsmpc_reference_glue.hpp) compiled for real: oracle/build_ref.build_glue()
builds it together with the reference's own QPInput.cpp / JetModel.cpp / VariableSamplingMPC (13 translation units read under
/root/reference) against the stand-in headers, linked to libvsmpc.so.  One QPInput object, two classes:
    reference  VariableSamplingMPC::configure(weak_ptr<IParametersHandler>, QPInput&) / update(QPInput&) / solveMPC()
    product    vsmpc::VariableSamplingMPCOnGpu — the same signatures over the C-ABI (fillPack -> vsmpc_set_state -> ...)
CPU: the binding compiles, loads, packs exactly what the Python pack builder packs, and fails loudly without a GPU.
GPU: 24 ticks driven like src/variable_sampling_mpc.py:106-131 (outputs fed back into QPInput), every getter per physical
quantity at 1e-6 and the four QPInput fields update() writes (costsVSMPC.cpp:155-160, systemDynamicsVSMPC.cpp:310) at 1e-12.
The library travels to the GPU box prebuilt (oracle/_ref/ is not gpurun-ignored)."""
import subprocess
import sys

import numpy as np
import pytest

from helpers import field_rel, load_trajectories, pkg

GLUE_PROBE = """
import sys
sys.path[:0] = [{root!r}, {tests!r}]
import numpy as np
from helpers import load_trajectories, pkg
import reference_driver as rd
if rd.lib(glue=True) is None:
    print("NOLIB"); raise SystemExit(0)
syn, P = pkg("synthetic"), pkg("pack")
nom = syn.make_states(3, seed=4, perturbed=True)
worst = 0.0
for i in range(3):
    r = rd.ReferenceInstance(nom, i, trajectories=load_trajectories(), glue=True)
    pk = r.glue_fill_pack(P.DEFAULT_JOINT_SELECTOR)
    worst = max(worst, float(np.abs(pk - P.build_pack(nom)[:, i]).max()))
    ok = r.glue_configure()
    r.close()
print("PACKDIFF", worst, "CONFIGURE", ok)
"""


def test_glue_compiles_packs_like_python_and_fails_loudly_without_gpu():
    """Runs in a subprocess: the glue library and the plain reference library define the same symbols."""
    import os
    import torch
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-c", GLUE_PROBE.format(root=root, tests=os.path.join(root, "tests"))],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    if "NOLIB" in out.stdout:
        pytest.skip("oracle/_ref/libvsmpc_reference_glue.so not built (no /root/reference here)")
    line = [l for l in out.stdout.splitlines() if l.startswith("PACKDIFF")][0].split()
    # vsmpc::fillPack == pack.build_pack; the only rounding is rpy = getRotation().asRPY(), which the binding (like the
    # reference) recomputes from the rotation matrix while the synthetic states carry the angles themselves
    assert float(line[1]) < 1e-15
    assert line[3] == str(torch.cuda.is_available())    # no CUDA device: configure() returns false (no CPU fallback)


GLUE_LOOP = """
import sys
sys.path[:0] = [{root!r}, {tests!r}]
import json
import numpy as np
from helpers import field_rel, load_trajectories, pkg
import reference_driver as rd
assert rd.lib(glue=True) is not None
syn, P = pkg("synthetic"), pkg("pack")
sel = P.DEFAULT_JOINT_SELECTOR
traj = load_trajectories()
B, ticks = 2, 24
worst = dict(getters=0.0, published=0.0, solution=0.0)
for i in range(B):
    nom = syn.make_states(B, seed=3, perturbed=False)
    r = rd.ReferenceInstance(nom, i, trajectories=traj, glue=True)      # configures the reference class
    pub_ref = r.output()["qp_input"].copy()
    r.clear_published()
    assert r.glue_configure()                                            # ... and the GPU-backed class on the same QPInput
    pub = r.glue_output()["qp_input"]
    worst["published"] = max(worst["published"], float(np.abs(pub - pub_ref).max() / max(1.0, np.abs(pub_ref).max())))
    assert (r.n_var, r.n_con) == (r.L.ref_glue_nvar(r.h), r.L.ref_glue_ncon(r.h))
    fb = None
    for t in range(ticks):
        st = syn.make_states(B, seed=900 + t, perturbed=True, near_bound_fraction=0.5)
        if fb is not None:      # src/variable_sampling_mpc.py:124-131: outputs of the last tick go back into QPInput
            st["throttle_prev"][i], st["thrust_des"][i], st["thrust_dot_des"][i] = fb["throttle"], fb["thrust"], fb["thrust_dot"]
            st["q_cmd"][i] = fb["joints"]
        r.update(st)
        zr = r.solve()
        o = r.output()
        assert o["status"] == 1
        r.clear_published()
        assert r.glue_update() and r.glue_solve()
        g = r.glue_output()
        assert g["ok"] and g["status"] == 0
        for k in ("throttle", "thrust", "thrust_dot"):
            worst["getters"] = max(worst["getters"], field_rel(g[k], o[k]))
        worst["getters"] = max(worst["getters"], field_rel(g["joints"][sel], o["joints"][sel]))
        assert np.array_equal(np.delete(g["joints"], sel), np.delete(o["joints"], sel))
        for f in range(4):
            worst["getters"] = max(worst["getters"], field_rel(g["final"][3 * f:3 * f + 3], o["final"][3 * f:3 * f + 3]))
        worst["solution"] = max(worst["solution"], float(np.abs(g["solution"] - zr).max() / max(1.0, np.abs(zr).max())))
        worst["published"] = max(worst["published"], float(np.abs(g["qp_input"] - o["qp_input"]).max() / max(1.0, np.abs(o["qp_input"]).max())))
        fb = o
    r.close()
print("RESULT", json.dumps(worst))
"""


@pytest.mark.gpu
def test_reference_class_and_gpu_class_on_one_qpinput_object():
    import json
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if not os.path.exists(os.path.join(root, "oracle", "_ref", "libvsmpc_reference_glue.so")):
        pytest.skip("oracle/_ref/libvsmpc_reference_glue.so not built")
    out = subprocess.run([sys.executable, "-c", GLUE_LOOP.format(root=root, tests=os.path.join(root, "tests"))],
                         capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, (out.stdout[-2000:], out.stderr[-3000:])
    w = json.loads([l for l in out.stdout.splitlines() if l.startswith("RESULT")][0][7:])
    assert w["getters"] < 1e-6 and w["solution"] < 1e-6, w       # north_star tolerance, per physical quantity
    assert w["published"] < 1e-12, w                             # the QPInput fields update() writes
