"""The header-only C++ adapter (include/vsmpc_adapter.hpp) mirrors the reference class
`VariableSamplingMPC` (variableSamplingMPC.h:15-41).  CPU: it compiles with g++ against the C-ABI and
fails loudly without a GPU.  GPU: examples/cpp_controller.cpp runs a closed loop of ticks with the
reference driver's feedback (src/variable_sampling_mpc.py:124-131) and must reproduce the Python host
layer's outputs on the same inputs."""
import os
import subprocess

import numpy as np
import pytest

from helpers import ROOT, pkg

N_TICKS = 25          # crosses the 20-tick throttle release


def _inputs(tmp_path):
    syn, pack = pkg("synthetic"), pkg("pack")
    d = np.load(os.path.join(ROOT, "tests", "golden", "trajectories.npz"))
    nom = syn.make_states(1, perturbed=False)
    packs = [pack.build_pack(syn.make_states(1, seed=100 + t, perturbed=True, near_bound_fraction=0.0))[:, 0]
             for t in range(N_TICKS)]
    sel = list(pack.DEFAULT_JOINT_SELECTOR)
    blob = np.concatenate([
        [d["alphaGravity"].size, d["positionCoM"].shape[1], N_TICKS, nom["joint_pos"].shape[1]],
        d["alphaGravity"].ravel(), d["positionCoM"].T.ravel(), d["velocityCoM"].T.ravel(), d["RPY"].T.ravel(),
        d["RPYDot"].T.ravel(), nom["joint_pos"][0], np.array(sel, dtype=np.float64),
        pack.build_pack(nom)[:, 0]] + packs).astype(np.float64)
    fin = str(tmp_path / "in.bin")
    blob.tofile(fin)
    return fin, nom, packs, sel, d


def test_adapter_compiles_and_fails_loudly_without_gpu(tmp_path):
    exe = pkg("_build").build_examples(force=True)
    assert os.path.exists(exe)
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu test")
    fin, *_ = _inputs(tmp_path)
    res = subprocess.run([exe, fin, str(tmp_path / "out.bin")], capture_output=True, text=True)
    assert res.returncode != 0                      # no CPU fallback
    assert "vsmpc" in res.stderr


@pytest.mark.gpu
def test_cpp_controller_matches_python_host(tmp_path):
    exe = pkg("_build").build_examples()
    fin, nom, packs, sel, d = _inputs(tmp_path)
    fout = str(tmp_path / "out.bin")
    res = subprocess.run([exe, fin, fout], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    nJ = nom["joint_pos"].shape[1]
    got = np.fromfile(fout).reshape(N_TICKS, 12 + nJ + 12 + 1)
    # the same loop through the Python mirror of the C-ABI
    bat, P, L = pkg("batched"), pkg("pack"), pkg("_lib")
    traj = pkg("config").load_trajectories_npz(os.path.join(ROOT, "tests", "golden", "trajectories.npz"))
    mpc = bat.BatchedVSMPC(1, None, traj, full_solution=True)
    mpc.configure(nom)
    off = P.PACK_OFFSETS
    jref = nom["joint_pos"][0].copy()
    out = None
    for t in range(N_TICKS):
        pk = packs[t].copy()[:, None]
        if t > 0:
            pk[off["throttle_prev"][0]:off["throttle_prev"][0] + 4, 0] = out[L.OUT_THROTTLE:L.OUT_THROTTLE + 4]
            pk[off["thrust_des"][0]:off["thrust_des"][0] + 4, 0] = out[L.OUT_THRUST:L.OUT_THRUST + 4]
            pk[off["thrust_dot_des"][0]:off["thrust_dot_des"][0] + 4, 0] = out[L.OUT_THRUST_DOT:L.OUT_THRUST_DOT + 4]
            pk[off["q_cmd"][0]:off["q_cmd"][0] + 8, 0] = jref[sel]
        mpc.update_pack(np.ascontiguousarray(pk))
        mpc.solveMPC()
        o, st = mpc.get_output()
        out = o[0]
        jref[sel] = out[L.OUT_JOINTS_REF:L.OUT_JOINTS_REF + 8]
        exp = np.concatenate([out[L.OUT_THROTTLE:L.OUT_THROTTLE + 12], jref,
                              out[L.OUT_FINAL_STATE:L.OUT_FINAL_STATE + 12], [float(st[0])]])
        np.testing.assert_allclose(got[t], exp, rtol=1e-12, atol=1e-12)
    mpc.close()
