"""Per-instance model parameters (BASELINE.json configs[4]: jet time constants, mass / inertia, thrust limits with
per-instance constraint sets): every instance gets its own jet coefficients, normalisation and throttle box; mass and
inertia already travel per instance in the pack.  Parity vs the oracle configured instance by instance."""
import numpy as np
import pytest

from helpers import load_trajectories, pkg
from oracle_driver import OracleInstance, oracle_trajectories_to_product

pytestmark = pytest.mark.gpu


def sweep(B, seed=77):
    from oracle.vsmpc_oracle import JetModel
    g = np.random.default_rng(seed)
    base = JetModel()
    coeff = np.tile(np.asarray(base.c, float), (B, 1))
    norm = np.tile(np.asarray(base.n, float), (B, 1))
    coeff[:, 1] *= g.uniform(0.9, 1.1, B)      # c1, c2: time-constant proxies (SURVEY §8d Config 5)
    coeff[:, 2] *= g.uniform(0.9, 1.1, B)
    norm[:, 0] *= g.uniform(0.9, 1.1, B)       # mu_T
    norm[:, 1] *= g.uniform(0.9, 1.1, B)       # sigma_T
    tmin, tmax = g.uniform(0.0, 20.0, B), g.uniform(80.0, 100.0, B)
    ms, isc = g.uniform(0.9, 1.1, B), g.uniform(0.8, 1.2, B)
    return coeff, norm, tmin, tmax, ms, isc


@pytest.mark.parametrize("solver", [0, 1, 2])
def test_parameter_sweep_matches_oracle(solver):
    from oracle.vsmpc_oracle import JetModel
    B = 16
    coeff, norm, tmin, tmax, ms, isc = sweep(B)
    syn, bat = pkg("synthetic"), pkg("batched")
    traj = load_trajectories()
    nom = syn.make_states(B, perturbed=False, mass_scale=ms, inertia_scale=isc)
    per = syn.make_states(B, seed=9, perturbed=True, near_bound_fraction=0.3, mass_scale=ms, inertia_scale=isc)
    mpc = bat.BatchedVSMPC(B, None, oracle_trajectories_to_product(traj), solver=solver, full_solution=True)
    mpc.set_instance_params(coeff, norm, tmin, tmax)
    mpc.configure(nom)
    mpc.update(per)
    mpc.solveMPC()
    z = mpc.getSolution()
    out, status = mpc.get_output()
    assert (status == 0).all()
    n_active = 0
    for i in range(B):
        o = OracleInstance(nom, i, params=dict(throttleMin=tmin[i], throttleMax=tmax[i]), trajectories=traj,
                           jet_model=JetModel(coeff[i], norm[i]))
        o.update(per)
        zo = o.solve()
        n_active += o.mpc.solveInfo["n_active"]
        assert np.abs(z[i] - zo).max() / max(1.0, np.abs(zo).max()) < 1e-6
        row = o.output_row()
        assert np.abs(out[i] - row).max() / max(1.0, np.abs(row).max()) < 1e-6
    assert n_active > 0
    # switching the table off returns to the handle-wide parameters
    mpc.set_instance_params()
    mpc.update(per)
    mpc.solveMPC()
    o = OracleInstance(nom, 0, trajectories=traj)
    o.update(per)         # first tick of a fresh oracle == GPU's second tick only in the bounds; compare the box
    q, l, u = mpc.get_qp_vectors()
    assert abs(l[0, 472] - o.mpc.lowerBound[472]) < 1e-12 and abs(u[0, 472] - o.mpc.upperBound[472]) < 1e-12
    mpc.close()


def test_rollout_with_parameter_sweep():
    """configs[2] x configs[4]: closed loops with per-instance jet models, plant and controller sharing them."""
    from oracle.vsmpc_oracle import JetModel
    from oracle.plant_surrogate import SurrogateLoop, SurrogatePlant
    from test_rollout import geometry, make_case
    B = 4
    coeff, norm, tmin, tmax, _, _ = sweep(B, seed=5)
    rb, st, ms, isc, dT = make_case(B, seed=21)
    bat, ro = pkg("batched"), pkg("rollout")
    traj = load_trajectories()
    mpc = bat.BatchedVSMPC(B, None, oracle_trajectories_to_product(traj))
    mpc.set_instance_params(coeff, norm, tmin, tmax)
    loop = ro.BatchedRollout(mpc, rb)
    loop.init(st, mass_scale=ms, inertia_scale=isc, thrust_disturbance=dT)
    rec = loop.run(25, record_every=1)
    for i in range(B):
        R = st["wRb"][i]
        state = dict(p_com=st["p_com"][i], lin_mom_world=R @ st["momentum_body"][i, :3], rpy=st["rpy"][i],
                     ang_mom_body=st["momentum_body"][i, 3:], thrust=st["thrust"][i], thrust_dot=st["thrust_dot_est"][i],
                     throttle=st["throttle_prev"][i], thrust_des=st["thrust_des"][i],
                     thrust_dot_des=st["thrust_dot_des"][i], q_cmd=st["q_cmd"][i])
        plant = SurrogatePlant(geometry(rb), rb.mass * ms[i], rb.I_body * isc[i], dT[i], state,
                               jet_model=JetModel(coeff[i], norm[i]))
        o = SurrogateLoop(plant, trajectories=traj, params=dict(throttleMin=tmin[i], throttleMax=tmax[i]))
        for t in range(25):
            r = o.tick()
            assert np.abs(rec[t, i, 0:6] - r[0:6]).max() < 1e-7
            assert np.abs(rec[t, i, 6:14] - r[6:14]).max() / 100.0 < 1e-6
    mpc.close()


def test_cuda_graph_of_the_tick_follows_later_setters():
    """The captured tick bakes kernel arguments in (per-instance table pointer, full-solution flag): changing them after a
    graph run must not replay the stale arguments.  Two handles run the same loop, one switching its per-instance table
    on between two graph runs, the other from the start without graphs: identical plant states."""
    syn, bat, ro = pkg("synthetic"), pkg("batched"), pkg("rollout")
    B = 6
    rb = syn.SyntheticRobot()
    g = np.random.default_rng(3)
    st = syn.make_states(B, seed=3, perturbed=True, near_bound_fraction=0.0)
    st["thrust"] = np.full((B, 4), rb.mass * 9.81 / 4.0)
    st["thrust_des"] = st["thrust"].copy()
    st["q_cmd"] = np.tile(rb.joint_pos0, (B, 1))
    tmax = g.uniform(80, 95, B)
    trj = oracle_trajectories_to_product(load_trajectories())

    def loop(use_graph):
        mpc = bat.BatchedVSMPC(B, None, trj)
        lp = ro.BatchedRollout(mpc, rb)
        lp.init(st)
        lp.run(3, use_graph=use_graph)
        mpc.set_instance_params(throttle_max=tmax)          # after the graph was captured
        lp.run(3, use_graph=use_graph)
        ps = lp.plant_state()
        mpc.close()
        return ps

    assert np.array_equal(loop(True), loop(False))
