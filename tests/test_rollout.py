"""Closed-loop rollouts (BASELINE.json configs[2]): device-resident loop (surrogate plant + K1 + K2 on the GPU,
csrc/vsmpc_plant.cu) against the oracle MPC driven in closed loop over the oracle's own restatement of the same
surrogate plant (oracle/plant_surrogate.py), sequenced like src/variable_sampling_mpc.py:106-161.

Stated bound (SURVEY §8c asks for <= 1e-4 m / 1e-4 rad over 10 s): here <= 1e-7 m / rad, 1e-6 relative on thrust and
throttle, over 45 ticks (two throttle releases), FP64 on both sides."""
import numpy as np
import pytest

from helpers import load_trajectories, pkg, robot_data
from oracle_driver import oracle_trajectories_to_product

N_TICKS = 45


def geometry(rb):
    return dict(com_from_base_body=rb.com_from_base_body, jet_pos_body=rb.jet_pos_body, jet_axes_body=rb.jet_axes_body,
                J_rel_body=rb.J_rel_body, J_jet_lin_body=rb.J_jet_lin_body, J_com_body=rb.J_com_body,
                gravity=rb.gravity, joint_pos0=rb.joint_pos0)


def oracle_plant(rb, st, i, mass_scale, inertia_scale, dT):
    from oracle.plant_surrogate import SurrogatePlant
    R = st["wRb"][i]
    state = dict(p_com=st["p_com"][i], lin_mom_world=R @ st["momentum_body"][i, :3], rpy=st["rpy"][i],
                 ang_mom_body=st["momentum_body"][i, 3:], thrust=st["thrust"][i], thrust_dot=st["thrust_dot_est"][i],
                 throttle=st["throttle_prev"][i], thrust_des=st["thrust_des"][i], thrust_dot_des=st["thrust_dot_des"][i],
                 q_cmd=st["q_cmd"][i])
    return SurrogatePlant(geometry(rb), rb.mass * mass_scale, rb.I_body * inertia_scale, dT, state)


def make_case(B, seed=11):
    syn = pkg("synthetic")
    rb = syn.SyntheticRobot()
    g = np.random.default_rng(seed)
    ms, isc = g.uniform(0.95, 1.05, B), g.uniform(0.9, 1.1, B)
    dT = g.normal(0, 10.0, (B, 4))
    st = syn.make_states(B, seed=seed, perturbed=True, near_bound_fraction=0.2, mass_scale=ms, inertia_scale=isc)
    # a loop starts from a consistent state: thrusts near hover, small rates
    st["thrust"] = np.full((B, 4), rb.mass * 9.81 / 4.0) + g.normal(0, 8.0, (B, 4))
    st["thrust_des"] = st["thrust"].copy()
    st["thrust_dot_est"] = g.normal(0, 5.0, (B, 4))
    st["thrust_dot_des"] = np.zeros((B, 4))
    # the loop starts at the posture the frozen kinematics refer to (the jet frames then follow q_cmd - q0)
    st["q_cmd"] = np.tile(rb.joint_pos0, (B, 1))
    # the generator draws omega_B independently of the angular momentum; a plant state has w_B = I_B^-1 h_ang^B
    for i in range(B):
        wB = np.linalg.solve(rb.I_body * isc[i], st["momentum_body"][i, 3:])
        st["omega_world"][i] = st["wRb"][i] @ wB
    return rb, st, ms, isc, dT


def test_oracle_plant_pack_matches_host_pack():
    """Two independent constructions of the getter-level data from a plant state agree (no GPU)."""
    B = 5
    rb, st, ms, isc, dT = make_case(B)
    for i in range(B):
        pl = oracle_plant(rb, st, i, ms[i], isc[i], dT[i])
        a, b = pl.robot_data(), robot_data(st, i)
        for k, v in a.__dict__.items():
            if isinstance(v, np.ndarray):
                np.testing.assert_allclose(v, getattr(b, k), rtol=1e-13, atol=1e-13, err_msg=k)


def test_oracle_closed_loop_is_stable():
    """Sanity of the surrogate itself: in flight (alphaGravity = 1) from the hover thrust, the oracle loop keeps the CoM
    within 8 cm and the attitude within 80 mrad over 120 ticks while it trims the asymmetric jet torques."""
    from oracle.plant_surrogate import SurrogateLoop
    syn = pkg("synthetic")
    rb = syn.SyntheticRobot()
    st = syn.make_states(1, perturbed=False)
    st["thrust"][:] = rb.mass * 9.81 / 4.0
    st["thrust_des"][:] = rb.mass * 9.81 / 4.0
    st["throttle_prev"][:] = 76.0
    traj = load_trajectories()
    traj["TRAJECTORY_MANAGER"]["arrays"]["alphaGravity"] = np.ones_like(traj["TRAJECTORY_MANAGER"]["arrays"]["alphaGravity"])
    loop = SurrogateLoop(oracle_plant(rb, st, 0, 1.0, 1.0, np.zeros(4)), trajectories=traj)
    p0 = loop.plant.s["p_com"].copy()
    worst_p = worst_a = 0.0
    for _ in range(120):
        r = loop.tick()
        worst_p, worst_a = max(worst_p, np.abs(r[:3] - p0).max()), max(worst_a, np.abs(r[3:6]).max())
    assert np.all(np.isfinite(r)) and worst_p < 0.08 and worst_a < 0.08, (worst_p, worst_a)


@pytest.mark.gpu
@pytest.mark.parametrize("use_graph", [False, True])
def test_device_rollout_matches_oracle_loop(use_graph):
    from oracle.plant_surrogate import SurrogateLoop
    B = 6
    rb, st, ms, isc, dT = make_case(B)
    bat, ro = pkg("batched"), pkg("rollout")
    traj = load_trajectories()
    mpc = bat.BatchedVSMPC(B, None, oracle_trajectories_to_product(traj))
    loop = ro.BatchedRollout(mpc, rb)
    loop.init(st, mass_scale=ms, inertia_scale=isc, thrust_disturbance=dT)
    rec = loop.run(N_TICKS, record_every=1, use_graph=use_graph)
    assert rec.shape == (N_TICKS, B, 16)
    assert (rec[:, :, 14] == 0).all()
    worst_p = worst_a = worst_t = worst_u = 0.0
    for i in range(B):
        o = SurrogateLoop(oracle_plant(rb, st, i, ms[i], isc[i], dT[i]), trajectories=traj)
        for t in range(N_TICKS):
            r = o.tick()
            worst_p = max(worst_p, np.abs(rec[t, i, 0:3] - r[0:3]).max())
            worst_a = max(worst_a, np.abs(rec[t, i, 3:6] - r[3:6]).max())
            worst_t = max(worst_t, np.abs(rec[t, i, 6:10] - r[6:10]).max() / 100.0)
            worst_u = max(worst_u, np.abs(rec[t, i, 10:14] - r[10:14]).max() / 100.0)
    assert worst_p < 1e-7 and worst_a < 1e-7, (worst_p, worst_a)
    assert worst_t < 1e-6 and worst_u < 1e-6, (worst_t, worst_u)
    # state read-back agrees with the last record
    ps = loop.plant_state()
    np.testing.assert_allclose(ps[0:3].T, rec[-1, :, 0:3], rtol=0, atol=1e-15)
    mpc.close()


@pytest.mark.gpu
def test_rollout_first_pack_matches_host_pack():
    B = 8
    rb, st, ms, isc, dT = make_case(B, seed=5)
    bat, ro, P = pkg("batched"), pkg("rollout"), pkg("pack")
    mpc = bat.BatchedVSMPC(B, None, oracle_trajectories_to_product(load_trajectories()))
    loop = ro.BatchedRollout(mpc, rb)
    loop.init(st, mass_scale=ms, inertia_scale=isc, thrust_disturbance=dT)
    np.testing.assert_allclose(loop.pack(), P.build_pack(st), rtol=1e-13, atol=1e-13)
    mpc.close()


@pytest.mark.gpu
def test_device_rollout_long_horizon_matches_oracle_loop():
    """The device-resident loop on a 2x-knot horizon (default solver -> the condensed kernel with several column warps)
    against the oracle loop: same bounds as the reference horizon, 22 ticks (one throttle release)."""
    from oracle.plant_surrogate import SurrogateLoop
    B, n_ticks = 3, 22
    params = dict(nIter=34, nIterSmall=14, controlHorizon=24)
    rb, st, ms, isc, dT = make_case(B, seed=23)
    bat, ro = pkg("batched"), pkg("rollout")
    traj = load_trajectories()
    mpc = bat.BatchedVSMPC(B, params, oracle_trajectories_to_product(traj))
    loop = ro.BatchedRollout(mpc, rb)
    loop.init(st, mass_scale=ms, inertia_scale=isc, thrust_disturbance=dT)
    rec = loop.run(n_ticks, record_every=1, use_graph=True)
    assert (rec[:, :, 14] == 0).all()
    worst_p = worst_a = worst_t = worst_u = 0.0
    for i in range(B):
        o = SurrogateLoop(oracle_plant(rb, st, i, ms[i], isc[i], dT[i]), trajectories=traj, params=params)
        for t in range(n_ticks):
            r = o.tick()
            worst_p = max(worst_p, np.abs(rec[t, i, 0:3] - r[0:3]).max())
            worst_a = max(worst_a, np.abs(rec[t, i, 3:6] - r[3:6]).max())
            worst_t = max(worst_t, np.abs(rec[t, i, 6:10] - r[6:10]).max() / 100.0)
            worst_u = max(worst_u, np.abs(rec[t, i, 10:14] - r[10:14]).max() / 100.0)
    assert worst_p < 1e-7 and worst_a < 1e-7, (worst_p, worst_a)
    assert worst_t < 1e-6 and worst_u < 1e-6, (worst_t, worst_u)
    mpc.close()


@pytest.mark.gpu
def test_logged_rollout_writes_the_reference_log_keys(tmp_path):
    """rollout.run_logged + save_log_mat: the 23 series of the reference driver's end-of-run .mat file
    (src/variable_sampling_mpc.py:163-186), consistent with the device loop's own record of the same loop."""
    import scipy.io
    syn, bat, ro = pkg("synthetic"), pkg("batched"), pkg("rollout")
    B, n = 2, 25
    traj = dict(oracle_trajectories_to_product(load_trajectories()))
    traj["alphaGravity"] = np.ones_like(traj["alphaGravity"])
    st = syn.make_states(B, seed=77, perturbed=True, near_bound_fraction=0.0)
    runs = []
    for logged in (True, False):
        mpc = bat.BatchedVSMPC(B, None, traj)
        loop = ro.BatchedRollout(mpc, syn.SyntheticRobot())
        loop.init(st)
        runs.append(ro.run_logged(loop, n, instance=1) if logged else loop.run(n, record_every=1, use_graph=False))
        mpc.close()
    log, rec = runs
    f = str(tmp_path / "log.mat")
    ro.save_log_mat(f, log)
    d = scipy.io.loadmat(f)
    assert all(k in d for k in ro.LOG_KEYS) and len(ro.LOG_KEYS) == 23
    for k, cols, rows in (("CoMPosition", 3, n + 1), ("CoMPosition_desired", 3, n + 1), ("base_orientation_desired", 3, n + 1),
                          ("linear_momentum", 3, n + 1), ("angular_momentum", 3, n + 1), ("momentum_reference", 6, n + 1),
                          ("base_position", 3, n), ("base_orientation", 3, n), ("base_lin_vel", 3, n), ("base_ang_vel", 3, n),
                          ("joints_pos_meas", 23, n), ("joints_pos_ref", 23, n), ("estimated_thrust", 4, n), ("thrust_desired", 4, n),
                          ("thrust_desired_dot", 4, n), ("estimated_thrust_dot", 4, n), ("throttle", 4, n), ("mom_dot", 6, n)):
        assert d[k].shape == (rows, cols), (k, d[k].shape)
    assert d["alpha_gravity"].size == n and d["time_MPC"].size == n
    assert np.allclose(d["time_controller"].ravel(), 0.005 * (1 + np.arange(n)))
    # the same loop recorded by the device: CoM, attitude, thrust and throttle agree tick by tick
    assert np.allclose(d["CoMPosition"][1:], rec[:, 1, 0:3], rtol=0, atol=1e-12)
    assert np.allclose(d["base_orientation"], rec[:, 1, 3:6], rtol=0, atol=1e-12)
    assert np.allclose(d["estimated_thrust"], rec[:, 1, 6:10], rtol=0, atol=1e-9)
    # (the record's throttle is the one in effect during the plant steps, which the plant switches on its own schedule)
    assert np.isfinite(d["throttle"]).all() and (d["throttle"] >= 0).all() and (d["throttle"] <= 100).all()
