"""GPU: the CUDA path (through the C-ABI) against the frozen golden vectors of tests/golden/: golden_*.npz (frozen from the
oracle) and reference_*.npz (frozen from the reference's own code compiled against stand-in headers,
tests/golden/make_reference_golden.py — the same scenarios, every number from the reference)."""
import numpy as np
import pytest

from helpers import assert_output_rows_close, assert_solution_close, field_rel, golden, load_trajectories, pkg
from oracle_driver import oracle_trajectories_to_product

pytestmark = pytest.mark.gpu
REL = 1e-6   # north_star tolerance (FP64)


def rel_err(a, b):
    return np.abs(a - b).max() / max(1.0, np.abs(b).max())


@pytest.mark.parametrize("source", ["golden_qp.npz", "reference_qp.npz"])
@pytest.mark.parametrize("solver", [0, 1, 2])
def test_single_tick_golden(solver, source):
    g = golden(source)
    bat = pkg("batched")
    B = g["nom_pack"].shape[1]
    for tag, free in (("pin", False), ("free", True)):
        if f"{tag}_z" not in g.files:       # the reference-made file holds the first tick after configure only
            continue
        mpc = bat.BatchedVSMPC(B, None, oracle_trajectories_to_product(load_trajectories()), solver=solver,
                               full_solution=True)
        mpc.configure_pack(g["nom_pack"], g["joint_pos_sel"])
        if free:
            mpc.debug_set_counters(-1, 19)
        mpc.update_pack(g["per_pack"])
        A, BJ, BT, c, _ = mpc.get_dynamics()
        q, l, u = mpc.get_qp_vectors()
        mpc.solveMPC()
        z = mpc.getSolution()
        out, status = mpc.get_output()
        assert (status == 0).all()
        if f"{tag}_A" in g.files:            # the continuous-time blocks are private members of the reference's classes
            assert rel_err(A, g[f"{tag}_A"]) < 1e-12 and rel_err(BJ, g[f"{tag}_BJ"]) < 1e-12
            assert rel_err(BT, g[f"{tag}_BT"]) < 1e-12 and rel_err(c, g[f"{tag}_c"]) < 1e-12
        assert rel_err(q, g[f"{tag}_q"]) < 1e-12 and rel_err(l, g[f"{tag}_l"]) < 1e-12 and rel_err(u, g[f"{tag}_u"]) < 1e-12
        for i in range(B):       # every physical quantity against its own magnitude
            assert_solution_close(z[i], g[f"{tag}_z"][i], REL, what=(tag, "z", i))
        assert_output_rows_close(out, g[f"{tag}_row"], REL, what=(tag, "row"))
        mpc.close()


@pytest.mark.parametrize("source", ["golden_ticks.npz", "reference_ticks.npz"])
@pytest.mark.parametrize("solver,full", [(0, False), (0, True), (1, True), (2, False)])
def test_tick_sequence_golden(solver, full, source):
    """24 consecutive ticks: reference-window shift and throttle release on tick 20, alpha_g cursor, RPY
    unwrapping through +-pi, joint accumulator, output hold — all device-resident state."""
    g = golden(source)
    bat = pkg("batched")
    B = g["nom_pack"].shape[1]
    mpc = bat.BatchedVSMPC(B, None, oracle_trajectories_to_product(load_trajectories()), solver=solver,
                           full_solution=full)
    mpc.configure_pack(g["nom_pack"], g["joint_pos_sel"])
    for t in range(g["packs"].shape[0]):
        mpc.update_pack(g["packs"][t])
        mpc.solveMPC()
        out, status = mpc.get_output()
        assert (status == 0).all()
        assert_output_rows_close(out, g["rows"][t], REL, what=("tick", t))
        if full:
            z = mpc.getSolution()
            for i in range(B):
                assert_solution_close(z[i], g["z"][t][i], REL, what=("tick", t, "z", i))
    mpc.close()


def test_status_gate_holds_outputs_on_bad_input():
    """Non-finite data -> status 2 and the previous outputs are held (variableSamplingMPC.cpp:91)."""
    g = golden("golden_qp.npz")
    bat = pkg("batched")
    B = g["nom_pack"].shape[1]
    mpc = bat.BatchedVSMPC(B, None, oracle_trajectories_to_product(load_trajectories()))
    mpc.configure_pack(g["nom_pack"], g["joint_pos_sel"])
    mpc.update_pack(g["per_pack"])
    mpc.solveMPC()
    out0, st0 = mpc.get_output()
    bad = g["per_pack"].copy()
    bad[331, 2] = np.nan          # thrust of instance 2
    mpc.update_pack(bad)
    mpc.solveMPC()
    out1, st1 = mpc.get_output()
    assert st1[2] == 2 and (np.delete(st1, 2) == 0).all()
    assert np.array_equal(out1[2], out0[2])            # held
    assert not np.array_equal(out1[0][46:54], out0[0][46:54])   # the others accumulated a second increment
    mpc.close()


def test_failed_first_tick_holds_the_configure_time_posture():
    """A tick that fails before any success must hold what configure() left: the joint reference is Robot::getJointPos()
    at configure time (variableSamplingMPC.cpp:60,91), the other getters are zero — not an uninitialised or stale row.  A
    re-configure resets the rows of a previous run."""
    g = golden("golden_qp.npz")
    bat, mp = pkg("batched"), pkg("mpc")
    B = g["nom_pack"].shape[1]
    mpc = bat.BatchedVSMPC(B, None, oracle_trajectories_to_product(load_trajectories()))
    for attempt in range(2):
        mpc.configure_pack(g["nom_pack"], g["joint_pos_sel"])
        out, st = mpc.get_output()
        assert (st == 0).all() and np.array_equal(out[:, 46:54], g["joint_pos_sel"].T) and not out[:, :46].any()
        bad = g["per_pack"].copy()
        bad[331, 1] = np.nan                       # thrust of instance 1
        mpc.update_pack(bad)
        mpc.solveMPC()
        out, st = mpc.get_output()
        assert st[1] == 2 and (np.delete(st, 1) == 0).all()
        assert np.array_equal(out[1, 46:54], g["joint_pos_sel"][:, 1]) and not out[1, :46].any()
        assert np.abs(out[0, 8:12]).max() > 0      # the others solved
        mpc.update_pack(g["per_pack"])             # run on, so that the second configure has something to reset
        mpc.solveMPC()
    mpc.close()


def test_single_instance_mirror_matches_batched():
    from helpers import state_from_pack
    g = golden("golden_qp.npz")
    mp = pkg("mpc")
    nom = state_from_pack(g["nom_pack"], g["joint_pos_sel"])
    per = state_from_pack(g["per_pack"])
    i = 3

    def robot(st):
        return dict(wRb=st["wRb"][i], base_pos=st["base_pos"][i], omega_world=st["omega_world"][i], rpy=st["rpy"][i],
                    M_b=st["M_b"][i], p_com=st["p_com"][i], momentum_body=st["momentum_body"][i],
                    A_mom_body=st["A_mom_body"][i], jet_axes=st["jet_axes"][i], jet_arms=st["jet_arms"][i],
                    J_rel_body=st["J_rel_body"][i], J_jet_lin=st["J_jet_lin"][i], J_com=st["J_com"][i],
                    thrust=st["thrust"][i], joint_pos=st["joint_pos"][i], gravity=st["gravity"][i])

    def fill(qp, st):
        qp.setThrottleMPC(st["throttle_prev"][i]); qp.setThrustDesMPC(st["thrust_des"][i])
        qp.setThrustDotDesMPC(st["thrust_dot_des"][i]); qp.setEstimatedThrustDot(st["thrust_dot_est"][i])
        qp.setOutputQPJointsPosition(st["q_cmd"][i])

    r = mp.RobotState(**robot(nom))
    qp = mp.QPInput(); qp.setRobot(r); qp.setRobotReference(r); qp.setEmptyJetModel(); fill(qp, nom)
    m = mp.VariableSamplingMPC()
    assert m.configure(pkg("config").default_params(), qp, oracle_trajectories_to_product(load_trajectories()))
    r.setState(**robot(per)); fill(qp, per)
    assert m.update(qp) and m.solveMPC()
    row = g["pin_row"][i]
    assert rel_err(m.getThrustReference(), row[12:16]) < REL
    assert rel_err(m.getThrottleReference(), row[8:12]) < REL
    assert rel_err(m.getThrustDotReference(), row[16:20]) < REL
    assert rel_err(m.getJointsReferencePosition()[3:11], row[46:54]) < REL
    assert m.getJointsReferencePosition().shape == (23,)
    assert rel_err(m.getSolution(), g["pin_z"][i]) < REL
    assert m.getNStatesMPC() == 26.0 and m.getNInputMPC() == 12.0 and m.getNOptimizationVariables() == 588


@pytest.mark.parametrize("params", [None, dict(nIter=34, nIterSmall=14, controlHorizon=24), dict(useEstimatedThrust=False)])
def test_cuda_against_the_live_compiled_reference(params):
    """The CUDA path and the reference's own code (oracle/_ref/libvsmpc_reference.so, compiled from the reference's sources —
    it travels to the GPU box) side by side on states no golden file holds: consecutive ticks with a new perturbed state
    every tick (throttle release and reference-window shift included), gradient and bounds at 1e-12 on every tick, minimiser
    and the getters at the north-star tolerance."""
    import reference_driver
    if reference_driver.lib() is None:
        pytest.skip("oracle/_ref/libvsmpc_reference.so not built (no /root/reference here)")
    syn, bat, L = pkg("synthetic"), pkg("batched"), pkg("_lib")
    traj = load_trajectories()
    B, ticks = 3, 24
    nom = syn.make_states(B, seed=11, perturbed=False)
    refs = [reference_driver.ReferenceInstance(nom, i, params=params, trajectories=traj) for i in range(B)]
    mpc = bat.BatchedVSMPC(B, params, oracle_trajectories_to_product(traj), full_solution=True)
    mpc.configure(nom)
    assert all((r.n_var, r.n_con) == (mpc.n_var, mpc.n_con) for r in refs)
    sel = pkg("pack").DEFAULT_JOINT_SELECTOR
    N, Nc = mpc.params["nIter"], mpc.params["controlHorizon"]
    nblk = Nc - mpc.params["nIterSmall"] + 1
    for t in range(ticks):
        st = syn.make_states(B, seed=7000 + t, perturbed=True, near_bound_fraction=0.5)
        mpc.update(st)
        q, l, u = mpc.get_qp_vectors()
        pub = mpc.get_references()      # the QPInput fields update() writes (costsVSMPC.cpp:155-160, systemDynamicsVSMPC.cpp:310)
        mpc.solveMPC()
        z = mpc.getSolution()
        out, status = mpc.get_output()
        assert (status == 0).all(), (t, status)
        for i, r in enumerate(refs):
            r.update(st)
            rP, rq, rA, rl, ru = r.qp()
            assert rel_err(q[i], rq) < 1e-12 and rel_err(l[i], rl) < 1e-12 and rel_err(u[i], ru) < 1e-12, (t, i)
            if t in (0, 19, 20) and i == 0:
                # IMPCProblem::getHessian / getLinearConstraintMatrix, entry by entry incl. the sparsity pattern
                P, A = mpc.getHessian(i), mpc.getLinearConstraintMatrix(i)
                assert np.array_equal(P != 0, rP != 0) and np.abs(P - rP).max() <= 1e-12 * np.abs(rP).max()
                assert np.array_equal(A != 0, rA != 0), (t, np.argwhere((A != 0) != (rA != 0))[:5])
                assert np.abs(A - rA).max() <= 1e-12 * max(1.0, np.abs(rA).max())
            assert_solution_close(z[i], r.solve(), REL, N=N, Nc=Nc, nblk=nblk, what=(t, i))
            o = r.output()
            assert o["status"] == 1
            assert field_rel(out[i, L.OUT_THROTTLE:L.OUT_THROTTLE + 4], o["throttle"]) < REL
            assert field_rel(out[i, L.OUT_THRUST:L.OUT_THRUST + 4], o["thrust"]) < REL
            assert field_rel(out[i, L.OUT_THRUST_DOT:L.OUT_THRUST_DOT + 4], o["thrust_dot"]) < REL
            assert field_rel(out[i, L.OUT_JOINTS_REF:L.OUT_JOINTS_REF + 8], o["joints"][sel]) < REL
            for f in range(4):      # getFinalCoMPosition / LinMom / RPY / AngMom
                assert field_rel(out[i, L.OUT_FINAL_STATE + 3 * f:L.OUT_FINAL_STATE + 3 * f + 3], o["final"][3 * f:3 * f + 3]) < REL
            # the 13 published values: alphaGravity, posCoMReference, RPYReference, momentumReference
            mine = np.concatenate([[pub["alphaGravity"][i]], pub["posCoMReference"][i], pub["RPYReference"][i],
                                   pub["momentumReference"][i]])
            assert np.abs(mine - o["qp_input"]).max() <= 1e-12 * max(1.0, np.abs(o["qp_input"]).max()), (t, i, mine, o["qp_input"])
    for r in refs:
        r.close()
    mpc.close()
