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


def _batch_inputs(tmp_path, B, n_ticks):
    syn, pack = pkg("synthetic"), pkg("pack")
    d = np.load(os.path.join(ROOT, "tests", "golden", "trajectories.npz"))
    nom = syn.make_states(B, perturbed=False)
    sel = list(pack.DEFAULT_JOINT_SELECTOR)
    packs = [pack.build_pack(syn.make_states(B, seed=300 + t, perturbed=True, near_bound_fraction=0.3)) for t in range(n_ticks)]
    blob = np.concatenate([
        [d["alphaGravity"].size, d["positionCoM"].shape[1], n_ticks, nom["joint_pos"].shape[1], B],
        d["alphaGravity"].ravel(), d["positionCoM"].T.ravel(), d["velocityCoM"].T.ravel(), d["RPY"].T.ravel(),
        d["RPYDot"].T.ravel(), np.array(sel, dtype=np.float64), nom["joint_pos"].ravel(),
        pack.build_pack(nom).T.ravel()] + [p.T.ravel() for p in packs]).astype(np.float64)     # instance-major records
    fin = str(tmp_path / "in_batch.bin")
    blob.tofile(fin)
    return fin, nom, packs, sel


@pytest.mark.gpu
def test_cpp_one_process_multi_gpu_batch(tmp_path):
    """examples/cpp_multi_gpu.cpp: vsmpc::PackBatch (per-instance records -> SoA) + vsmpc::MultiGpuMPC (vsmpc_create_multi)
    in one process.  On a one-GPU box the shards share device 0 — the same host path, the same split of the SoA by column
    range —; with two or more GPUs they go to devices 0 and 1.  The program itself checks bit-identity with the single-device
    batch on every tick; here its rows are compared with the Python host layer."""
    import torch
    pkg("_build").build_examples()
    exe = os.path.join(ROOT, "examples", "bin", "cpp_multi_gpu")
    B, n_ticks = 37, 3           # ragged: shards of 12 / 12 / 13
    fin, nom, packs, sel = _batch_inputs(tmp_path, B, n_ticks)
    for devices in (["0", "0", "0"], ["0", "1"] if torch.cuda.device_count() >= 2 else ["0", "0"]):
        fout = str(tmp_path / "out_batch.bin")
        res = subprocess.run([exe, fin, fout, ",".join(devices)], capture_output=True, text=True)
        assert res.returncode == 0, (res.returncode, res.stderr)
        assert "bit-identical" in res.stdout
        got = np.fromfile(fout).reshape(n_ticks, B, 55)
        bat, P, L = pkg("batched"), pkg("pack"), pkg("_lib")
        traj = pkg("config").load_trajectories_npz(os.path.join(ROOT, "tests", "golden", "trajectories.npz"))
        mpc = bat.BatchedVSMPC(B, None, traj)
        mpc.configure(nom, phase0=(np.arange(B) % 20).astype(np.int32))
        off = P.PACK_OFFSETS
        out = None
        for t in range(n_ticks):
            pk = packs[t].copy()
            if t > 0:
                pk[off["throttle_prev"][0]:off["throttle_prev"][0] + 4] = out[:, L.OUT_THROTTLE:L.OUT_THROTTLE + 4].T
                pk[off["thrust_des"][0]:off["thrust_des"][0] + 4] = out[:, L.OUT_THRUST:L.OUT_THRUST + 4].T
                pk[off["thrust_dot_des"][0]:off["thrust_dot_des"][0] + 4] = out[:, L.OUT_THRUST_DOT:L.OUT_THRUST_DOT + 4].T
                pk[off["q_cmd"][0]:off["q_cmd"][0] + 8] = out[:, L.OUT_JOINTS_REF:L.OUT_JOINTS_REF + 8].T
            mpc.update_pack(np.ascontiguousarray(pk))
            mpc.solveMPC()
            out, st = mpc.get_output()
            assert np.array_equal(got[t, :, :54], out) and np.array_equal(got[t, :, 54], st.astype(float))
        mpc.close()


@pytest.mark.gpu
def test_python_multi_handle_matches_single(tmp_path):
    """vsmpc_multi_* through ctypes: ragged shards (incl. a shard without instances when B < G), per-instance parameters
    split by column range, full solution gathered — bit-identical to one handle."""
    import torch
    syn, bat = pkg("synthetic"), pkg("batched")
    traj = pkg("config").load_trajectories_npz(os.path.join(ROOT, "tests", "golden", "trajectories.npz"))
    ndev = torch.cuda.device_count()
    for B, G in ((50, 4), (3, 5)):
        nom = syn.make_states(B, perturbed=False)
        per = syn.make_states(B, seed=12, perturbed=True, near_bound_fraction=0.4)
        one = bat.BatchedVSMPC(B, None, traj, full_solution=True)
        multi = bat.MultiGpuVSMPC(B, None, traj, n_gpus=G, devices=[g % ndev for g in range(G)], full_solution=True)
        assert sum(c for _, c in multi.shards()) == B
        ph = (np.arange(B) % 20).astype(np.int32)
        one.configure(nom, ph); multi.configure(nom, ph)
        for _ in range(2):
            one.update(per); multi.update(per)
            one.solveMPC(); multi.solveMPC()
            (o1, s1), (o2, s2) = one.get_output(), multi.get_output()
            assert np.array_equal(o1, o2) and np.array_equal(s1, s2) and (s1 == 0).all()
            assert np.array_equal(one.getSolution(), multi.getSolution())
        one.close(); multi.close()


def test_pack_batch_blocked_scatter_equals_one_instance_at_a_time():
    """vsmpc::PackBatch::setMany (eight instances per cache line, several host threads, ranges off the 64-byte boundary) writes
    the same SoA as PackBatch::set — examples/cpp_host_bench --selftest, no GPU needed."""
    pkg("_build").build_examples()
    res = subprocess.run([os.path.join(ROOT, "examples", "bin", "cpp_host_bench"), "--selftest"], capture_output=True, text=True)
    assert res.returncode == 0 and "selftest ok" in res.stdout, (res.returncode, res.stdout, res.stderr)


@pytest.mark.gpu
def test_cpp_host_bench_runs_and_solves(tmp_path):
    """examples/cpp_host_bench.cpp: pack + upload + solve + read-back pipelined from C++ (page-locked PackBatch, two ticks in
    flight); every instance solved, and the pack is timed inside the loop."""
    import json
    pkg("_build").build_examples()
    exe = os.path.join(ROOT, "examples", "bin", "cpp_host_bench")
    fin, nom, packs, sel = _batch_inputs(tmp_path, 300, 3)
    res = subprocess.run([exe, fin, "12", "3", "2", "0"], capture_output=True, text=True)
    assert res.returncode == 0, (res.returncode, res.stderr)
    out = json.loads(res.stdout.strip().splitlines()[-1])
    assert out["instances"] == 300 and out["steps"] == 12 and out["pinned"] is True
    assert out["solved_fraction_last_step"] == 1.0 and out["value"] > 0 and 0 < out["pack_ms_per_step"] < out["ms_per_step"]
