"""Neural jet plant + per-jet EKF.  Both are PINNED: tests/golden/jet_nn.npz holds outputs of the
reference's own torch module (tests/golden/make_jet_nn_golden.py imports src/mujoco_lib/nn_jet_model.py in the build
container), tests/golden/jet_ekf.npz outputs of the reference's own jet_kalman_filter.py executed over a CasADi stand-in
(tests/golden/make_jet_ekf_golden.py)."""
import numpy as np
import pytest

from helpers import golden, pkg


def weights():
    g = golden("jet_nn.npz")
    return g, {k: g[k] for k in ("w_ih", "b_ih", "b_hh", "fc_w", "fc_b", "norm")}


def test_oracle_nn_step_matches_reference_outputs():
    from oracle.jet_nn_oracle import nn_jet_step
    g, w = weights()
    dt = float(g["dt"])
    for k in range(g["T_in"].shape[0]):
        T, Td = nn_jet_step(g["T_in"][k], g["u_in"][k], w, dt)
        np.testing.assert_allclose(T, g["T_out"][k], rtol=2e-6, atol=2e-5)
        np.testing.assert_allclose(Td, g["Td_out"][k], rtol=2e-5, atol=2e-4)


def test_oracle_nn_sequence_matches_reference_outputs():
    from oracle.jet_nn_oracle import nn_jet_step
    g, w = weights()
    T = np.full(4, 10.0, np.float32)
    for k in range(g["seq_u"].shape[0]):
        T, Td = nn_jet_step(T, g["seq_u"][k], w, float(g["dt"]))
        np.testing.assert_allclose(T, g["seq_T"][k], rtol=1e-5, atol=1e-4)
    np.testing.assert_allclose(Td, g["seq_Td"][-1], rtol=1e-4, atol=1e-3)


def test_oracle_ekf_matches_reference_outputs():
    """The oracle's EKF against outputs of the reference's own jet_kalman_filter.py (executed unmodified over a CasADi
    stand-in, tests/golden/make_jet_ekf_golden.py): model step, Jacobian, and the four-jet filter over 400 steps with
    the simulator's covariances.  FP64 on both sides: 1e-10 relative."""
    from oracle.jet_nn_oracle import JetEKF
    g = golden("jet_ekf.npz")
    dt = float(g["dt"])
    e = JetEKF(g["R"], g["Q"], g["P0"], dt)
    for k in range(g["pts_x"].shape[0]):
        np.testing.assert_allclose(e.f(g["pts_x"][k], g["pts_u"][k]), g["pts_f"][k], rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(e.A(g["pts_x"][k], g["pts_u"][k]), g["pts_A"][k], rtol=1e-10, atol=1e-12)
    jets = [JetEKF(g["R"], g["Q"], g["P0"], dt) for _ in range(4)]
    T, Td = g["T0"].copy(), g["Td0"].copy()
    for k in range(g["seq_u"].shape[0]):
        for i in range(4):
            T[i], Td[i] = jets[i].update([T[i], Td[i]], g["seq_u"][k, i], [g["seq_zT"][k, i], g["seq_zTd"][k, i]])
        np.testing.assert_allclose(T, g["seq_T"][k], rtol=1e-10, atol=1e-9)
        np.testing.assert_allclose(Td, g["seq_Td"][k], rtol=1e-10, atol=1e-9)
    for i in range(4):
        np.testing.assert_allclose(jets[i].P, g["P_end"][i], rtol=1e-10, atol=1e-13)


def test_ekf_jacobian_is_the_derivative_of_its_model():
    from oracle.jet_nn_oracle import JetEKF
    e = JetEKF(np.eye(2) * 0.5, np.eye(2) * 0.1, np.eye(2) * 0.1, 0.001)
    rng = np.random.default_rng(1)
    for _ in range(10):
        x = np.array([rng.uniform(20, 200), rng.normal(0, 30)]); u = rng.uniform(0, 100)
        A = e.A(x, u)
        num = np.zeros((2, 2))
        for j in range(2):
            h = 1e-5 * max(1.0, abs(x[j])); d = np.zeros(2); d[j] = h
            num[:, j] = (e.f(x + d, u) - e.f(x - d, u)) / (2 * h)
        np.testing.assert_allclose(A, num, rtol=1e-6, atol=1e-8)


@pytest.mark.gpu
def test_device_rollout_with_neural_jet_plant_matches_oracle_loop():
    """Closed loop with the reference simulator's jet plant + estimator pair on the device (float32 network, float64 EKF)
    against the oracle loop.  Bound: 1e-5 m / rad, 1e-4 relative thrust / throttle over 30 ticks (the network is float32 on
    both sides; expf / tanhf differ in the last place between the two)."""
    from helpers import load_trajectories
    from oracle_driver import oracle_trajectories_to_product
    from oracle.plant_surrogate import SurrogateLoop, SurrogatePlant
    from test_rollout import geometry, make_case
    g, w = weights()
    B, n_ticks = 4, 30
    rb, st, ms, isc, dT = make_case(B, seed=31)
    bat, ro = pkg("batched"), pkg("rollout")
    traj = load_trajectories()
    mpc = bat.BatchedVSMPC(B, None, oracle_trajectories_to_product(traj))
    loop = ro.BatchedRollout(mpc, rb)
    loop.set_jet_nn(w)
    loop.init(st, mass_scale=ms, inertia_scale=isc, thrust_disturbance=dT)
    rec = loop.run(n_ticks, record_every=1)
    assert (rec[:, :, 14] == 0).all()
    worst = np.zeros(3)
    for i in range(B):
        R = st["wRb"][i]
        state = dict(p_com=st["p_com"][i], lin_mom_world=R @ st["momentum_body"][i, :3], rpy=st["rpy"][i],
                     ang_mom_body=st["momentum_body"][i, 3:], thrust=st["thrust"][i], thrust_dot=st["thrust_dot_est"][i],
                     throttle=st["throttle_prev"][i], thrust_des=st["thrust_des"][i],
                     thrust_dot_des=st["thrust_dot_des"][i], q_cmd=st["q_cmd"][i])
        plant = SurrogatePlant(geometry(rb), rb.mass * ms[i], rb.I_body * isc[i], dT[i], state, jet_nn=w)
        o = SurrogateLoop(plant, trajectories=traj)
        for t in range(n_ticks):
            r = o.tick()
            worst[0] = max(worst[0], np.abs(rec[t, i, 0:6] - r[0:6]).max())
            worst[1] = max(worst[1], np.abs(rec[t, i, 6:10] - r[6:10]).max() / 100.0)
            worst[2] = max(worst[2], np.abs(rec[t, i, 10:14] - r[10:14]).max() / 100.0)
    assert worst[0] < 1e-5 and worst[1] < 1e-4 and worst[2] < 1e-4, worst
    # the jets really follow the network: thrusts differ from a run with the second-order jet model
    mpc2 = bat.BatchedVSMPC(B, None, oracle_trajectories_to_product(traj))
    loop2 = ro.BatchedRollout(mpc2, rb)
    loop2.init(st, mass_scale=ms, inertia_scale=isc, thrust_disturbance=dT)
    rec2 = loop2.run(n_ticks, record_every=1)
    assert np.abs(rec2[-1, :, 6:10] - rec[-1, :, 6:10]).max() > 0.5
    mpc.close(); mpc2.close()


@pytest.mark.gpu
def test_device_nn_step_matches_reference_outputs():
    """The CUDA network step against outputs of the reference's torch module (single steps and the 400-step sequence)."""
    from helpers import load_trajectories
    from oracle_driver import oracle_trajectories_to_product
    g, w = weights()
    bat, ro = pkg("batched"), pkg("rollout")
    mpc = bat.BatchedVSMPC(1, None, oracle_trajectories_to_product(load_trajectories()))
    loop = ro.BatchedRollout(mpc)
    loop.set_jet_nn(w)
    T, Td = loop.jet_nn_eval(g["T_in"], g["u_in"], float(g["dt"]))
    np.testing.assert_allclose(T, g["T_out"], rtol=2e-6, atol=2e-5)
    np.testing.assert_allclose(Td, g["Td_out"], rtol=5e-5, atol=5e-4)
    Ts = np.full((1, 4), 10.0, np.float32)
    for k in range(g["seq_u"].shape[0]):
        Ts, Tds = loop.jet_nn_eval(Ts, g["seq_u"][k][None], float(g["dt"]))
        # a float32 recursion of 400 steps: last-place differences of expf / tanhf between the CUDA and the torch CPU
        # kernels accumulate; 1e-4 relative is the north_star tolerance for FP32 arithmetic
        np.testing.assert_allclose(Ts[0], g["seq_T"][k], rtol=1e-4, atol=1e-4)
    mpc.close()
