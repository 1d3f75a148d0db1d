"""GPU parity: CUDA path (through the C-ABI) vs the CPU oracle on the same seeded inputs.

Tolerances (BASELINE.json north_star): optimal thrusts / joint commands within 1e-6 relative (FP64).
The assembly quantities (A, B, c, q, l, u) are required to agree to 1e-12 relative.
"""
import numpy as np
import pytest

from helpers import assert_output_rows_close, assert_solution_close, load_trajectories, pkg
from oracle_driver import OracleInstance, oracle_trajectories_to_product

pytestmark = pytest.mark.gpu

REL_SOL = 1e-6      # north_star tolerance on the solution
REL_ASM = 1e-12     # assembly (linearisation, gradient, bounds)


def rel_err(a, b):
    return np.abs(a - b).max() / max(1.0, np.abs(b).max())


def make(B, solver, seed=20251002, near=0.10, params=None, full=True):
    syn = pkg("synthetic")
    bat = pkg("batched")
    traj = load_trajectories()
    nom = syn.make_states(B, perturbed=False)
    per = syn.make_states(B, seed=seed, perturbed=True, near_bound_fraction=near)
    mpc = bat.BatchedVSMPC(B, params, oracle_trajectories_to_product(traj), solver=solver, full_solution=full)
    return mpc, nom, per, traj


@pytest.mark.parametrize("solver", [0, 1, 2])
def test_linearise_matches_oracle(solver):
    B = 24
    mpc, nom, per, traj = make(B, solver)
    mpc.configure(nom)
    mpc.update(per)
    A, BJ, BT, c, dt = mpc.get_dynamics()
    q, l, u = mpc.get_qp_vectors()
    for i in range(B):
        o = OracleInstance(nom, i, trajectories=traj)
        o.update(per)
        oA, oBJ, oBT, oc, odt = o.dynamics()
        assert rel_err(A[i], oA) < REL_ASM
        assert rel_err(BJ[i], oBJ) < REL_ASM
        assert rel_err(BT[i], oBT) < REL_ASM
        assert rel_err(c[i], oc) < REL_ASM
        assert rel_err(dt, odt) < 1e-15
        assert rel_err(q[i], o.mpc.gradient) < REL_ASM
        assert rel_err(l[i], o.mpc.lowerBound) < REL_ASM
        assert rel_err(u[i], o.mpc.upperBound) < REL_ASM
    mpc.close()


@pytest.mark.parametrize("solver", [0, 1, 2])
@pytest.mark.parametrize("free_tick", [False, True])
def test_solve_matches_exact_oracle(solver, free_tick):
    B = 32
    mpc, nom, per, traj = make(B, solver, near=0.3)
    mpc.configure(nom)
    if free_tick:
        mpc.debug_set_counters(-1, 19)   # next tick releases throttle block 0
    mpc.update(per)
    mpc.solveMPC()
    z = mpc.getSolution()
    out, status = mpc.get_output()
    nf, ns = mpc.get_counts()
    assert (status == 0).all()
    n_active = 0
    for i in range(B):
        o = OracleInstance(nom, i, trajectories=traj)
        if free_tick:
            o.mpc.vectorConstraints[2].counter = 19
        o.update(per)
        zo = o.solve()
        n_active += o.mpc.solveInfo["n_active"]
        # every physical quantity against its own magnitude (joint increments ~ 1e-3 rad, thrusts ~ 1e2 N)
        assert_solution_close(z[i], zo, REL_SOL, what=("z", i))
        assert_output_rows_close(out[i], o.output_row(), REL_SOL, what=("row", i))
    assert n_active > 0, "test workload should exercise the active-set path"
    assert (nf == 1).all() and (ns >= 1).all()
    mpc.close()


def test_inner_seams_called_alone():
    """K1 (vsmpc_linearise) and K2 (vsmpc_solve_qp) are separately callable; K1's product is checked
    against the oracle's assembly before K2 runs on it, K2's minimiser against the oracle's exact solve."""
    B = 16
    mpc, nom, per, traj = make(B, 0, near=0.3)
    mpc.configure(nom)
    mpc.linearise(per)
    A, BJ, BT, c, dt = mpc.get_dynamics()
    q, l, u = mpc.get_qp_vectors()
    oracles = []
    for i in range(B):
        o = OracleInstance(nom, i, trajectories=traj)
        o.update(per)
        oA, oBJ, oBT, oc, odt = o.dynamics()
        assert max(rel_err(A[i], oA), rel_err(BJ[i], oBJ), rel_err(BT[i], oBT), rel_err(c[i], oc)) < REL_ASM
        assert max(rel_err(q[i], o.mpc.gradient), rel_err(l[i], o.mpc.lowerBound), rel_err(u[i], o.mpc.upperBound)) < REL_ASM
        oracles.append(o)
    mpc.solve_qp()
    z = mpc.getSolution()
    _, status = mpc.get_output()
    assert (status == 0).all()
    for i, o in enumerate(oracles):
        assert_solution_close(z[i], o.solve(), REL_SOL, what=("z", i))
    # one solve per linearise, like the reference's tick
    with pytest.raises(pkg("batched").VsmpcError):
        mpc.solve_qp()
    mpc.close()


@pytest.mark.parametrize("solver", [0, 2])
@pytest.mark.parametrize("free_tick", [False, True])
def test_outputs_without_full_solution(solver, free_tick):
    """Default mode of the structured kernel: outputs by superposition, no 588-vector written."""
    B = 192
    mpc, nom, per, traj = make(B, solver, near=0.3, full=False)
    mpc.configure(nom)
    if free_tick:
        mpc.debug_set_counters(-1, 19)
    mpc.update(per)
    mpc.solveMPC()
    out, status = mpc.get_output()
    assert (status == 0).all()
    with pytest.raises(Exception):
        mpc.getSolution()
    for i in range(B):
        o = OracleInstance(nom, i, trajectories=traj)
        if free_tick:
            o.mpc.vectorConstraints[2].counter = 19
        o.update(per)
        o.solve()
        assert_output_rows_close(out[i], o.output_row(), REL_SOL, what=("row", i))
    mpc.close()
