"""Fallback QP kernel (csrc/vsmpc_qp_fallback.cu: pivoted LU of the KKT system in a bordered-band ordering).

The Riccati recursion of the condensed kernels breaks down on finite, well-posed QPs when the open-loop transition of the
linearised model expands strongly (a vehicle spinning at |omega_B| >~ 30 rad/s: 0.2 % of the Monte Carlo loops of BASELINE
configs[2], profiles/r02_nonsolved_adjudication.md); the reference's solver returns a minimiser there, so must the product.
  * every instance forced through the fallback (mode 2) reproduces the oracle per physical quantity — reference horizon,
    no held block, a single throttle block, 2x knots (the wide kernel's hand-over), pinned and released ticks;
  * spinning states: the default path (mode 1) solves them through the fallback and matches the oracle's exact solve; with
    the fallback off (mode 0) the status gate holds the outputs, as before;
  * the NumPy specification (tools/kkt_lu_model.py) against the oracle on the CPU (not a GPU test)."""
import numpy as np
import pytest

from helpers import assert_output_rows_close, assert_solution_close, load_trajectories, pkg
from oracle_driver import OracleInstance, oracle_trajectories_to_product

VARIANTS = [None, dict(nIter=17, nIterSmall=7, controlHorizon=17), dict(nIter=9, nIterSmall=7, controlHorizon=7),
            dict(nIter=34, nIterSmall=14, controlHorizon=24), dict(nIter=20, nIterSmall=2, controlHorizon=12)]


def horizon(o):
    p = o.params
    return dict(N=p["nIter"], Nc=p["controlHorizon"], nblk=p["controlHorizon"] - p["nIterSmall"] + 1)


def spinning_states(B, seed, rate=100.0):
    """Perturbed states of a vehicle spinning about its body z axis at `rate` rad/s (alternating sign, growing)."""
    syn = pkg("synthetic")
    per = syn.make_states(B, seed=seed, perturbed=True, near_bound_fraction=0.3)
    for i in range(B):
        w_body = np.array([0.5, -0.3, rate * (1 + i) * (-1) ** i])
        per["omega_world"][i] = per["wRb"][i] @ w_body
    return per


@pytest.mark.gpu
@pytest.mark.parametrize("variant", range(len(VARIANTS)))
def test_every_instance_through_the_fallback_matches_oracle(variant):
    params = VARIANTS[variant]
    B = 6
    syn, bat = pkg("synthetic"), pkg("batched")
    traj = load_trajectories()
    nom = syn.make_states(B, perturbed=False)
    per = syn.make_states(B, seed=41 + variant, perturbed=True, near_bound_fraction=0.5)
    mpc = bat.BatchedVSMPC(B, params, oracle_trajectories_to_product(traj), full_solution=True)
    mpc.set_fallback(2)
    mpc.configure(nom)
    oracles = [OracleInstance(nom, i, params=params, trajectories=traj) for i in range(B)]
    ratio = oracles[0].mpc.vectorConstraints[2].ratio
    n_active = 0
    for tick in range(3):
        if tick == 2:       # a released tick
            mpc.debug_set_counters(-1, ratio - 1)
        mpc.update(per)
        mpc.solveMPC()
        z = mpc.getSolution()
        out, status = mpc.get_output()
        nf, ns = mpc.get_counts()
        assert (status == 0).all(), status
        assert (nf == 2).all()         # Riccati attempt + LU factorisation: the fallback kernel produced these results
        for i, o in enumerate(oracles):
            if tick == 2:
                o.mpc.vectorConstraints[2].counter = ratio - 1
            o.update(per)
            zo = o.solve()
            n_active += o.mpc.solveInfo["n_active"]
            assert_solution_close(z[i], zo, 1e-6, what=(variant, tick, i), **horizon(o))
            assert_output_rows_close(out[i], o.output_row(), 1e-6, what=(variant, tick, i))
    assert n_active > 0
    mpc.close()


@pytest.mark.gpu
@pytest.mark.parametrize("params", [None, dict(nIter=34, nIterSmall=14, controlHorizon=24)])
def test_spinning_vehicle_is_solved_through_the_fallback(params):
    B = 8
    syn, bat = pkg("synthetic"), pkg("batched")
    traj = load_trajectories()
    nom = syn.make_states(B, perturbed=False)
    per = spinning_states(B, seed=5)
    # fallback off: the recursion breaks down, status 2, outputs held at what configure left
    mpc = bat.BatchedVSMPC(B, params, oracle_trajectories_to_product(traj), full_solution=True)
    mpc.set_fallback(0)
    mpc.configure(nom)
    mpc.update(per)
    mpc.solveMPC()
    out0, status0 = mpc.get_output()
    assert (status0 == 2).sum() >= B // 2, status0
    held = status0 != 0
    assert not out0[held, :46].any()
    mpc.close()
    # default: solved, and equal to the oracle's exact solve of the same QP
    mpc = bat.BatchedVSMPC(B, params, oracle_trajectories_to_product(traj), full_solution=True)
    mpc.configure(nom)
    mpc.update(per)
    mpc.solveMPC()
    z = mpc.getSolution()
    out, status = mpc.get_output()
    nf, ns = mpc.get_counts()
    assert (status == 0).all(), status
    assert ((nf == 2) == held).all()       # exactly the instances the recursion gave up on went through the fallback
    for i in range(B):
        o = OracleInstance(nom, i, params=params, trajectories=traj)
        o.update(per)
        zo = o.solve()
        assert_solution_close(z[i], zo, 1e-6, what=("spin", i), **horizon(o))
        assert_output_rows_close(out[i], o.output_row(), 1e-6, what=("spin", i))
    mpc.close()


@pytest.mark.parametrize("params,spin", [(None, False), (None, True), (dict(nIter=9, nIterSmall=7, controlHorizon=7), False),
                                         (dict(nIter=12, nIterSmall=5, controlHorizon=12), True)])
def test_kkt_lu_specification_matches_oracle(params, spin):
    import kkt_lu_model as km
    syn = pkg("synthetic")
    traj = load_trajectories()
    nom = syn.make_states(2, perturbed=False)
    per = spinning_states(2, seed=8) if spin else syn.make_states(2, seed=8, perturbed=True, near_bound_fraction=0.5)
    for i in range(2):
        for phase0 in (0, 19):
            o = OracleInstance(nom, i, params=params, trajectories=traj, phase0=phase0)
            o.update(per)
            z = o.solve()
            p, cs, rt, tc = o.params, o.mpc.vectorConstraints[0], o.mpc.vectorCosts[0], o.mpc.vectorConstraints[2]
            N, Ns, Nc = p["nIter"], p["nIterSmall"], p["controlHorizon"]
            jm = o.qp.getJetModel()
            vbar = np.array([jm.compute_v(jm.standardizeThrottle_u2T(u)) for u in o.qp.getThrottleMPC()])
            t0 = 26 * N + 26
            pinned = bool(o.mpc.lowerBound[t0] == o.mpc.upperBound[t0])
            x, dq, v, status, info = km.solve_kkt_lu(
                cs.A, cs.BJ, cs.BT, cs.c, cs.dt, np.diag(rt.Q).copy(), rt.stateReference.T.copy(),
                np.array(p["weightDeltaJoint"]) + p["weightRegularizationJointPos"],
                o.mpc.vectorCosts[3].gradient[26 * (N + 1):26 * (N + 1) + 8].copy(), p["weightThrottle"],
                p["weightInitialThrottle"], vbar, pinned, tc.vMin, tc.vMax, o.mpc.vectorConstraints[1].initialState, N, Ns, Nc)
            assert status == 0 and info["bw"] <= 64
            zf = np.concatenate([x.reshape(-1), dq.reshape(-1), v.reshape(-1)])
            assert_solution_close(zf, z, 1e-8, what=(params, spin, i, phase0), **horizon(o))
