"""Batched reduced kinematics (UT/src/Robot.cpp:212-278,325-332).

CPU: the NumPy restatement (oracle/kinematics_oracle.py) against what does not depend on it — finite differences of its own
forward kinematics (free-floating, relative and CoM Jacobians; centroidal momentum from finite-difference link motion), the
kinetic energy of a base twist (mass-matrix base block), and a closed-form planar two-link chain.  iDynTree itself and the
iRonCub URDF are absent: parity against the library is UNPINNED, the 23-DoF tree is synthetic.
GPU: csrc/vsmpc_kinematics.cu against the oracle on random states (every kinematic row of the pack, 1e-12), and a closed
loop with live kinematics — the joints follow the MPC's commands, the base moves, the device builds every tick's pack from
the robot state — against the oracle MPC fed by the NumPy kinematics."""
import numpy as np
import pytest

from helpers import assert_output_rows_close, load_trajectories, pkg
from oracle import kinematics_oracle as KO


def random_robot_state(model, B, seed, vel=1.0):
    syn = pkg("synthetic")
    g = np.random.default_rng(seed)
    nd = model["n_dof"]
    rpy = g.normal(0, 0.3, (B, 3))
    q0 = syn.SyntheticRobot().joint_pos0
    return dict(wRb=syn.rpy_to_R(rpy), base_pos=np.array([0.0, 0.0, 1.0]) + g.normal(0, 0.1, (B, 3)),
                base_lin_vel=vel * g.normal(0, 0.3, (B, 3)), omega_world=vel * g.normal(0, 0.4, (B, 3)),
                q=q0[None, :nd] + g.normal(0, 0.3, (B, nd)), qd=vel * g.normal(0, 0.5, (B, nd)),
                thrust=g.uniform(60, 200, (B, 4)), thrust_dot_est=g.normal(0, 20, (B, 4)), thrust_des=g.uniform(60, 200, (B, 4)),
                thrust_dot_des=g.normal(0, 10, (B, 4)), throttle_prev=g.uniform(20, 90, (B, 4)),
                q_cmd=q0[None, :nd] + g.normal(0, 0.02, (B, nd)))


def oracle_rows(model, rs, i):
    return KO.robot_set_state(model, rs["wRb"][i], rs["base_pos"][i], rs["base_lin_vel"][i], rs["omega_world"][i],
                              rs["q"][i], rs["qd"][i])


def test_oracle_jacobians_against_finite_differences_cpu():
    kin = pkg("kinematics")
    model = kin.synthetic_humanoid()
    rs = random_robot_state(model, 1, 3)
    o = oracle_rows(model, rs, 0)
    wRb, bp, q = rs["wRb"][0], rs["base_pos"][0], rs["q"][0]
    h = 1e-6

    def frames(qv):
        wR, wp = KO.forward_kinematics(model, wRb, bp, qv)
        m = np.asarray(model["mass"])
        pc = sum(m[l] * (wp[l] + wR[l] @ model["com"][l]) for l in range(model["n_links"])) / m.sum()
        pj = [wp[model["jet_link"][i]] + wR[model["jet_link"][i]] @ model["jet_pos"][i] for i in range(4)]
        Rj = [wR[model["jet_link"][i]] for i in range(4)]
        return pc, pj, Rj

    for j in range(model["n_dof"]):
        dq = np.zeros(model["n_dof"])
        dq[j] = h
        pc1, pj1, Rj1 = frames(q + dq)
        pc0, pj0, Rj0 = frames(q - dq)
        assert np.allclose((pc1 - pc0) / (2 * h), o["J_com_full"][:, j], atol=1e-8)
        for i in range(4):
            assert np.allclose((pj1[i] - pj0[i]) / (2 * h), o["J_jet_lin_full"][i][:, j], atol=1e-8)
            # relative angular velocity in base axes: wRb' vee(dR R')
            W = (Rj1[i] - Rj0[i]) / (2 * h) @ ((Rj1[i] + Rj0[i]) / 2).T
            w = np.array([W[2, 1], W[0, 2], W[1, 0]])
            assert np.allclose(wRb.T @ w, o["J_rel_ang_full"][i][:, j], atol=1e-7)
    # chest jets do not move with any joint; each arm jet only with its own arm
    assert not o["J_jet_lin_full"][2].any() and not o["J_rel_ang_full"][3].any()
    assert not o["J_jet_lin_full"][0][:, 7:].any() and o["J_jet_lin_full"][0][:, 3:7].any()


def test_oracle_momentum_and_mass_matrix_cpu():
    kin = pkg("kinematics")
    model = kin.synthetic_humanoid()
    rs = random_robot_state(model, 1, 4)
    o = oracle_rows(model, rs, 0)
    m = np.asarray(model["mass"])
    n = model["n_links"]
    # move the robot along its velocity for +-dt and differentiate the link CoM positions / orientations
    dt = 1e-6
    w = rs["omega_world"][0]

    def pose(s):
        dR = KO.rot_axis(w / np.linalg.norm(w), s * dt * np.linalg.norm(w))
        wR, wp = KO.forward_kinematics(model, dR @ rs["wRb"][0], rs["base_pos"][0] + s * dt * rs["base_lin_vel"][0],
                                       rs["q"][0] + s * dt * rs["qd"][0])
        return wR, [wp[l] + wR[l] @ model["com"][l] for l in range(n)]

    (R1, c1), (R0, c0) = pose(+1), pose(-1)
    vc = [(c1[l] - c0[l]) / (2 * dt) for l in range(n)]
    h_lin = sum(m[l] * vc[l] for l in range(n))
    h_ang = np.zeros(3)
    for l in range(n):
        W = (R1[l] - R0[l]) / (2 * dt) @ o["link_rot"][l].T
        wl = np.array([W[2, 1], W[0, 2], W[1, 0]])
        h_ang += o["link_rot"][l] @ model["inertia"][l] @ o["link_rot"][l].T @ wl + m[l] * np.cross(o["link_pos"][l] + o["link_rot"][l] @ model["com"][l] - o["p_com"], vc[l])
    assert np.allclose(h_lin, o["h_lin"], rtol=1e-6, atol=1e-7) and np.allclose(h_ang, o["h_ang"], rtol=1e-5, atol=1e-7)
    assert np.allclose(o["momentum_body"], np.concatenate([rs["wRb"][0].T @ o["h_lin"], rs["wRb"][0].T @ o["h_ang"]]))
    # base block of the mass matrix: kinetic energy of a pure base twist (joints locked) = 1/2 nu' M_b nu
    nu = np.concatenate([rs["base_lin_vel"][0], w])
    T = 0.0
    for l in range(n):
        cl = o["link_pos"][l] + o["link_rot"][l] @ model["com"][l]
        v = nu[:3] + np.cross(w, cl - rs["base_pos"][0])
        Iw = o["link_rot"][l] @ model["inertia"][l] @ o["link_rot"][l].T
        T += 0.5 * m[l] * v @ v + 0.5 * w @ Iw @ w
    assert abs(0.5 * nu @ o["M_b"] @ nu - T) <= 1e-12 * T
    assert np.allclose(o["M_b"], o["M_b"].T) and o["mass"] == float(np.float32(m.sum()))


def test_oracle_planar_two_link_chain_closed_form_cpu():
    """Base at the origin, two links of length 1 rotating about z: CoM, jet position Jacobian and A_mom in closed form."""
    z = np.array([0.0, 0.0, 1.0])
    model = dict(n_links=3, n_dof=8, parent=[-1, 0, 1], dof=[-1, 0, 1], R0=[np.eye(3)] * 3,
                 p0=[np.zeros(3), np.zeros(3), np.array([1.0, 0, 0])], axis=[z, z, z], mass=[2.0, 1.0, 1.0],
                 com=[np.zeros(3), np.array([0.5, 0, 0]), np.array([0.5, 0, 0])], inertia=[np.eye(3) * 0.1] * 3,
                 jet_link=[2, 2, 0, 0], jet_pos=[np.array([1.0, 0, 0])] * 4, jet_axis=[np.array([0.0, 1.0, 0.0])] * 4,
                 delta_com=np.zeros(3), gravity=np.array([0, 0, -9.81]), sel=list(range(8)))
    q = np.zeros(8)
    q[0], q[1] = 0.3, 0.5
    o = KO.robot_set_state(model, np.eye(3), np.zeros(3), np.zeros(3), np.zeros(3), q, np.zeros(8))
    c1 = 0.5 * np.array([np.cos(0.3), np.sin(0.3), 0])
    e1 = np.array([np.cos(0.3), np.sin(0.3), 0])
    e2 = np.array([np.cos(0.8), np.sin(0.8), 0])
    assert np.allclose(o["p_com"], (c1 + (e1 + 0.5 * e2)) / 4.0)
    tip = e1 + e2
    assert np.allclose(o["J_jet_lin"][0][:, 0], np.cross(z, tip)) and np.allclose(o["J_jet_lin"][0][:, 1], np.cross(z, e2))
    assert np.allclose(o["jet_axes"][0], [-np.sin(0.8), np.cos(0.8), 0])
    assert np.allclose(o["A_mom_body"][3:6, 0], np.cross(tip - o["p_com"], o["jet_axes"][0]))
    assert np.allclose(o["J_rel_ang"][0][:, 0], z) and not o["J_rel_ang"][2].any()


FIELDS = ("wRb", "omega_world", "rpy", "mass", "gravity", "M_b", "base_pos", "p_com", "momentum_body", "A_mom_body", "jet_axes",
          "jet_arms", "J_rel_ang", "J_jet_lin", "J_com")


@pytest.mark.gpu
def test_kinematics_kernel_matches_oracle():
    kin, bat, P = pkg("kinematics"), pkg("batched"), pkg("pack")
    model = kin.synthetic_humanoid()
    B = 37            # not a multiple of the eight instances of a CTA
    from oracle_driver import oracle_trajectories_to_product
    mpc = bat.BatchedVSMPC(B, None, oracle_trajectories_to_product(load_trajectories()))
    fe = kin.KinematicsFrontEnd(mpc, model)
    for seed, use_update in ((21, False), (22, True)):
        rs = random_robot_state(model, B, seed)
        fe.update(rs) if use_update else fe.configure(rs)
        pk = fe.pack()
        for i in range(B):
            o = oracle_rows(model, rs, i)
            for name in FIELDS:
                off, size = P.PACK_OFFSETS[name]
                ref = np.asarray(o[name], float).reshape(-1)
                got = pk[off:off + size, i]
                assert np.abs(got - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max()), (seed, i, name)
            off = P.PACK_OFFSETS["thrust"][0]
            ks = kin.build_kin_state(model, rs)
            assert np.array_equal(pk[off:, i], ks[kin.KS_Q + 2 * model["n_dof"]:, i])
    mpc.close()


@pytest.mark.gpu
def test_closed_loop_with_live_kinematics_matches_oracle():
    """Eight ticks: the device builds every pack from the robot state (joints at the MPC's last command, moving base); the
    oracle MPC is fed by the NumPy kinematics of the same state.  Outputs per physical quantity, 1e-6."""
    kin, bat, syn = pkg("kinematics"), pkg("batched"), pkg("synthetic")
    from oracle_driver import OracleInstance, oracle_trajectories_to_product
    model = kin.synthetic_humanoid()
    sel = list(model["sel"])
    B = 3
    traj = load_trajectories()
    rs = random_robot_state(model, B, 31, vel=0.2)
    rs["wRb"] = syn.rpy_to_R(np.random.default_rng(1).normal(0, 0.05, (B, 3)))
    rs["q"] = rs["q_cmd"].copy()
    hover = float(np.sum(model["mass"])) * 9.81 / 4.0
    rs["thrust"] = np.full((B, 4), hover) + np.random.default_rng(2).normal(0, 5.0, (B, 4))
    rs["thrust_des"] = rs["thrust"].copy()

    def state_dict(rs):
        rows = [oracle_rows(model, rs, i) for i in range(B)]
        st = {k: np.stack([np.asarray(r[k], float) for r in rows]) for k in
              ("wRb", "omega_world", "rpy", "gravity", "M_b", "base_pos", "p_com", "momentum_body", "A_mom_body", "jet_axes", "jet_arms")}
        st["mass"] = np.array([r["mass"] for r in rows])
        J = np.zeros((B, 4, 6, model["n_dof"]))
        J[:, :, 3:6, :] = np.stack([r["J_rel_ang_full"] for r in rows])
        st["J_rel_body"] = J
        st["J_jet_lin"] = np.stack([r["J_jet_lin_full"] for r in rows])
        st["J_com"] = np.stack([r["J_com_full"] for r in rows])
        for k in ("thrust", "thrust_dot_est", "thrust_des", "thrust_dot_des", "throttle_prev", "q_cmd"):
            st[k] = rs[k].copy()
        st["joint_pos"] = rs["q"].copy()
        return st

    mpc = bat.BatchedVSMPC(B, None, oracle_trajectories_to_product(traj))
    fe = kin.KinematicsFrontEnd(mpc, model)
    fe.configure(rs)
    st = state_dict(rs)
    oracles = [OracleInstance(st, i, trajectories=traj) for i in range(B)]
    g = np.random.default_rng(9)
    for tick in range(8):
        # the base drifts and rotates a little, the joints have reached the last command
        rs["base_pos"] = rs["base_pos"] + 0.005 * rs["base_lin_vel"]
        rs["wRb"] = np.stack([KO.rot_axis(np.array([0.0, 0.0, 1.0]), 0.002 * (tick + 1)) @ rs["wRb"][i] for i in range(B)])
        rs["qd"] = g.normal(0, 0.05, rs["qd"].shape)
        fe.update(rs)
        mpc.solveMPC()
        out, status = mpc.get_output()
        assert (status == 0).all(), status
        st = state_dict(rs)
        for i, o in enumerate(oracles):
            o.update(st)
            o.solve()
            assert_output_rows_close(out[i], o.output_row(), 1e-6, what=(tick, i))
        # feedback of the reference driver (src/variable_sampling_mpc.py:124-131) + position-controlled joints
        rs["throttle_prev"] = out[:, 8:12].copy()
        rs["thrust_des"] = out[:, 12:16].copy()
        rs["thrust_dot_des"] = out[:, 16:20].copy()
        rs["q_cmd"][:, sel] = out[:, 46:54]
        rs["q"][:, sel] = out[:, 46:54]
    mpc.close()
