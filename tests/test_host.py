"""CPU tests of the host layer: header <-> Python layout agreement, exported symbols, config reader,
argument errors of the C-ABI (no compute call is made without a GPU)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from helpers import ROOT, pkg

HEADER = open(os.path.join(ROOT, "include", "vsmpc.h")).read()


def test_pack_offsets_match_header():
    P = pkg("pack")
    defs = dict(re.findall(r"#define\s+(VSMPC_PK_\w+)\s+(\d+)", HEADER))
    names = {"wRb": "WRB", "omega_world": "OMEGA_WORLD", "rpy": "RPY", "mass": "MASS", "gravity": "GRAVITY",
             "M_b": "MB", "base_pos": "BASE_POS", "p_com": "P_COM", "momentum_body": "MOMENTUM_BODY",
             "A_mom_body": "AMOM_BODY", "jet_axes": "JET_AXES", "jet_arms": "JET_ARMS", "J_rel_ang": "J_REL_ANG",
             "J_jet_lin": "J_JET_LIN", "J_com": "J_COM", "thrust": "THRUST", "thrust_dot_est": "THRUST_DOT_EST",
             "thrust_des": "THRUST_DES", "thrust_dot_des": "THRUST_DOT_DES", "throttle_prev": "THROTTLE_PREV",
             "q_cmd": "Q_CMD"}
    for f, (off, size) in P.PACK_OFFSETS.items():
        assert int(defs["VSMPC_PK_" + names[f]]) == off, f
    m = re.search(r"#define\s+VSMPC_PACK_DOUBLES\s+(\d+)", HEADER)
    assert int(m.group(1)) == P.PACK_DOUBLES == 359
    L = pkg("_lib")
    for k in ("DELTA_Q", "THROTTLE", "THRUST", "THRUST_DOT", "FINAL_STATE", "JOINTS_REF", "DOUBLES"):
        v = int(re.search(rf"#define\s+VSMPC_OUT_{k}\s+(\d+)", HEADER).group(1))
        assert v == getattr(L, "OUT_" + k)


def test_library_exports_every_declared_symbol():
    L = pkg("_lib")
    lib = L.load()
    declared = set(re.findall(r"\b(vsmpc_[a-z0-9_]+)\s*\(", HEADER))
    declared -= {"vsmpc_handle", "vsmpc_config"}
    assert declared, "no functions parsed from the header"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/vsmpc.h but not exported"
    assert set(L.EXPORTS) == declared


def test_config_struct_layout_matches_c():
    # the C side of the same struct, compiled here with gcc
    import subprocess, tempfile
    src = '#include <stdio.h>\n#include <stddef.h>\n#include "vsmpc.h"\nint main(){printf("%zu %zu %zu\\n", sizeof(vsmpc_config), offsetof(vsmpc_config, solver), offsetof(vsmpc_config, jet_coeff));return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "t.c"), "-o", os.path.join(d, "t")])
        sz, off_solver, off_jet = [int(x) for x in subprocess.check_output([os.path.join(d, "t")]).split()]
    L = pkg("_lib")
    assert C.sizeof(L.VsmpcConfig) == sz
    assert L.VsmpcConfig.solver.offset == off_solver
    assert L.VsmpcConfig.jet_coeff.offset == off_jet


def test_plant_model_struct_layout_matches_c():
    import subprocess, tempfile
    src = '#include <stdio.h>\n#include <stddef.h>\n#include "vsmpc.h"\nint main(){printf("%zu %zu %zu\\n", sizeof(vsmpc_plant_model), offsetof(vsmpc_plant_model, q0), offsetof(vsmpc_plant_model, n_sub));return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "t.c"), "-o", os.path.join(d, "t")])
        sz, off_q0, off_nsub = [int(x) for x in subprocess.check_output([os.path.join(d, "t")]).split()]
    L = pkg("_lib")
    assert C.sizeof(L.VsmpcPlantModel) == sz
    assert L.VsmpcPlantModel.q0.offset == off_q0 and L.VsmpcPlantModel.n_sub.offset == off_nsub
    # row offsets of the plant state / parameter / instance-parameter tables
    for name, val in (("PS_P_COM", L.PS_P_COM), ("PS_Q_CMD", L.PS_Q_CMD), ("PLANT_STATE_DOUBLES", L.PLANT_STATE_DOUBLES),
                      ("PP_THRUST_DISTURBANCE", L.PP_THRUST_DISTURBANCE), ("PLANT_PARAM_DOUBLES", L.PLANT_PARAM_DOUBLES),
                      ("IP_THROTTLE_MAX", L.IP_THROTTLE_MAX), ("INSTANCE_PARAM_DOUBLES", L.INSTANCE_PARAM_DOUBLES),
                      ("ROLLOUT_REC_DOUBLES", L.ROLLOUT_REC_DOUBLES)):
        assert int(re.search(rf"#define\s+VSMPC_{name}\s+(\d+)", HEADER).group(1)) == val, name


def test_create_argument_errors_without_gpu():
    bat, L, cfg = pkg("batched"), pkg("_lib"), pkg("config")
    traj = cfg.hover_trajectories()
    with pytest.raises(bat.VsmpcError, match="nIterSmall"):
        bat.BatchedVSMPC(4, dict(nIter=17, nIterSmall=7, controlHorizon=5), traj)
    with pytest.raises(bat.VsmpcError, match="constant"):
        bat.BatchedVSMPC(4, dict(jointsLambdaOption="constant"), traj)
    with pytest.raises(bat.VsmpcError, match="unfiltered"):
        bat.BatchedVSMPC(4, dict(jointsLambdaOption="bogus"), traj)
    with pytest.raises(bat.VsmpcError, match="joint deltas"):
        bat.BatchedVSMPC(4, dict(weightDeltaJoint=[1.0] * 7), traj)
    # rates / periods the resampling loops and the 20-tick ratio would loop or divide by zero on
    with pytest.raises(bat.VsmpcError, match="rates"):
        bat.BatchedVSMPC(4, None, dict(traj, alpha_fps=0))
    with pytest.raises(bat.VsmpcError, match="rates"):
        bat.BatchedVSMPC(4, None, dict(traj, traj_fps=-10))
    with pytest.raises(bat.VsmpcError, match="ratio"):
        bat.BatchedVSMPC(4, dict(periodMPCLargeSteps=0.001, periodMPCSmallSteps=0.005), traj)
    with pytest.raises(bat.VsmpcError, match="1 Hz"):
        bat.BatchedVSMPC(4, dict(periodMPC=2.0), traj)
    lib = L.load()
    assert lib.vsmpc_create(None, 4, 0, None) == L.ERR_ARG
    assert lib.vsmpc_get_references(None, None) == L.ERR_ARG and lib.vsmpc_get_hessian(None, 0, None) == L.ERR_ARG
    assert lib.vsmpc_get_constraint_matrix(None, 0, None) == L.ERR_ARG and lib.vsmpc_get_pivot_counts(None, None) == L.ERR_ARG
    assert lib.vsmpc_n_var(None) == -1
    assert lib.vsmpc_solve(None) == L.ERR_ARG


def test_product_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    bat, cfg = pkg("batched"), pkg("config")
    with pytest.raises(bat.VsmpcError, match="cuda|CUDA|device"):
        bat.BatchedVSMPC(4, None, cfg.hover_trajectories())


def test_xml_config_reader(tmp_path):
    cfg = pkg("config")
    xml = tmp_path / "c.xml"
    xml.write_text('''<?xml version="1.0" encoding="UTF-8" ?>
<robot name="x"><device name="d" type="dummy"><group name="VS_MPC_CONFIG">
<param name="useJetDynamic">true</param><param name="periodMPC">0.005</param><param name="nIter">17</param>
<param name="controlledJoints">("l_shoulder_pitch", "l_elbow")</param>
<param name="weightCoMPos">(500.0 500.0 5000.0)</param><param name="jointsLambdaOption">"unfiltered"</param>
<group name="TRAJECTORY_MANAGER"><param name="trajectoryFile">"a.mat"</param></group>
</group></device></robot>''')
    p = cfg.read_xml_config(str(xml))
    assert p["useJetDynamic"] is True and p["periodMPC"] == 0.005 and p["nIter"] == 17
    assert p["controlledJoints"] == ["l_shoulder_pitch", "l_elbow"]
    assert p["weightCoMPos"] == [500.0, 500.0, 5000.0] and p["jointsLambdaOption"] == "unfiltered"
    assert p["TRAJECTORY_MANAGER"]["trajectoryFile"] == "a.mat"
    d = cfg.default_params()
    assert d["nIter"] == 17 and d["nIterSmall"] == 7 and d["controlHorizon"] == 12


def test_pack_builder_roundtrip():
    from helpers import state_from_pack
    syn, P = pkg("synthetic"), pkg("pack")
    st = syn.make_states(5, perturbed=True)
    pack = P.build_pack(st)
    assert pack.shape == (359, 5) and pack.dtype == np.float64
    back = state_from_pack(pack)
    assert np.array_equal(P.build_pack(back), pack)
    with pytest.raises(ValueError):
        bad = dict(st); bad["wRb"] = st["wRb"][:, :2]
        P.build_pack(bad)


def test_mat73_reader_on_npz_equivalent(tmp_path):
    """The MAT-v7.3 reader is exercised against the reference files in tests/golden/make_fixtures.py; here the
    committed fixture must at least be self-consistent with what the loader returns."""
    cfg = pkg("config")
    t = cfg.load_trajectories_npz(os.path.join(ROOT, "tests", "golden", "trajectories.npz"))
    assert t["alpha_fps"] == 10 and t["traj_fps"] == 10 and t["alphaGravity"].shape == (1, 351)


def test_mat73_reader_on_v73_files_generated_in_test(tmp_path):
    """mat73.loadmat73 / config.load_trajectories_mat on MAT-v7.3 (HDF5) files written by the test itself (tests/hdf5_min.py:
    h5py is not in this image): contiguous and chunked + deflate float64 datasets, the two files and variable names of the
    reference's TrajectoryManager (UT/src/TrajectoryManager.cpp:67-140: fps, alphaGravity; fps, positionCoM, velocityCoM,
    RPY, RPYDot), the shapes of the reference's own files (1 x 351, 3 x 1481)."""
    from hdf5_min import write_mat73
    m73, cfg = pkg("mat73"), pkg("config")
    g = np.random.default_rng(5)
    a = {"fps": np.array([[10.0]]), "alphaGravity": g.random((1, 351))}
    t = {"fps": np.array([[10.0]]), "positionCoM": g.normal(size=(3, 1481)), "velocityCoM": g.normal(size=(3, 1481)),
         "RPY": g.normal(size=(3, 1481)), "RPYDot": g.normal(size=(3, 1481))}
    fa, ft = str(tmp_path / "alphaGravity.mat"), str(tmp_path / "minimumJerkTrajectory.mat")
    write_mat73(fa, a)
    write_mat73(ft, t, chunked={"positionCoM": 500, "RPYDot": 1481, "velocityCoM": 7})
    d = m73.loadmat73(ft)
    assert set(d) == set(t)
    for k in t:
        assert d[k].shape == t[k].shape and np.array_equal(d[k], t[k]), k
    tr = cfg.load_trajectories_mat(fa, ft)
    assert tr["alpha_fps"] == 10 and tr["traj_fps"] == 10
    assert np.array_equal(tr["alphaGravity"], a["alphaGravity"]) and np.array_equal(tr["RPY"], t["RPY"])
    with pytest.raises(ValueError):
        m73.Mat73(os.path.join(ROOT, "tests", "golden", "trajectories.npz"))      # not an HDF5 file


def test_rollout_log_writer_roundtrip(tmp_path):
    """The rollout record is written with the reference driver's log keys (src/variable_sampling_mpc.py:163-194)."""
    import scipy.io
    ro = pkg("rollout")
    rec = np.random.default_rng(0).normal(size=(12, 3, 16))
    f = str(tmp_path / "log.mat")
    ro.save_log_mat(f, rec, 1, record_every=2)
    d = scipy.io.loadmat(f)
    assert np.array_equal(d["CoMPosition"], rec[:, 1, 0:3]) and np.array_equal(d["throttle"], rec[:, 1, 10:14])
    assert np.allclose(d["time_controller"].ravel(), 0.01 * (1 + np.arange(12)))
