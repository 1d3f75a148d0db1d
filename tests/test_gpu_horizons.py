"""Horizon variants (BASELINE.json configs[3]: long-horizon variable-sampling variants, 2-4x the reference knot count
with a coarse tail dt) and model options, every solver against the oracle.  The default solver routes itself:
condensed kernel where the horizon has <= 6 throttle blocks, the condensed kernel with several column warps
(vsmpc_qp_condensed_wide.cu) for longer horizons, the generic dense kernel beyond that."""
import numpy as np
import pytest

from helpers import assert_output_rows_close, assert_solution_close, load_trajectories, pkg
from oracle_driver import OracleInstance, oracle_trajectories_to_product

pytestmark = pytest.mark.gpu

HORIZONS = [
    dict(nIter=12, nIterSmall=5, controlHorizon=8),
    dict(nIter=17, nIterSmall=7, controlHorizon=17),                       # no held joint block
    dict(nIter=20, nIterSmall=7, controlHorizon=12),
    dict(nIter=9, nIterSmall=7, controlHorizon=7),                         # a single throttle block
    dict(nIter=34, nIterSmall=14, controlHorizon=24),                      # 2x knots (SURVEY §8d Config 4)
    dict(nIter=51, nIterSmall=14, controlHorizon=36),                      # 3x knots
    dict(nIter=34, nIterSmall=14, controlHorizon=24, periodMPCLargeSteps=0.2),   # coarse tail dt
    dict(useJetDynamic=False),
    dict(useEstimatedThrust=False),
]


@pytest.mark.parametrize("solver", [0, 1, 2])
@pytest.mark.parametrize("variant", range(len(HORIZONS)))
def test_horizon_variant_matches_oracle(solver, variant):
    params = HORIZONS[variant]
    B = 6
    syn, bat = pkg("synthetic"), pkg("batched")
    traj = load_trajectories()
    nom = syn.make_states(B, perturbed=False)
    per = syn.make_states(B, seed=31 + variant, perturbed=True, near_bound_fraction=0.3)
    try:
        mpc = bat.BatchedVSMPC(B, params, oracle_trajectories_to_product(traj), solver=solver, full_solution=True)
    except bat.VsmpcError as e:
        assert solver == 2 and "structured" in str(e)      # the one-warp structured kernel is reference-horizon only
        return
    mpc.configure(nom)
    for free in (False, True):
        if free:
            mpc.debug_set_counters(-1, int(round(mpc.params["periodMPCLargeSteps"] / mpc.params["periodMPCSmallSteps"])) - 1)
        mpc.update(per)
        mpc.solveMPC()
        z = mpc.getSolution()
        out, status = mpc.get_output()
        assert (status == 0).all(), status
        if not free:
            oracles = [OracleInstance(nom, i, params=params, trajectories=traj) for i in range(B)]
        for i, o in enumerate(oracles):
            if free:
                o.mpc.vectorConstraints[2].counter = o.mpc.vectorConstraints[2].ratio - 1 \
                    if hasattr(o.mpc.vectorConstraints[2], "ratio") else 19
            o.update(per)
            zo = o.solve()
            assert z.shape[1] == zo.size
            hz = dict(N=o.params["nIter"], Nc=o.params["controlHorizon"], nblk=o.params["controlHorizon"] - o.params["nIterSmall"] + 1)
            assert_solution_close(z[i], zo, 1e-6, what=(variant, solver, free, i), **hz)     # per physical quantity
            assert_output_rows_close(out[i], o.output_row(), 1e-6, what=(variant, solver, free, i))
    mpc.close()


# edge cases of the condensed kernel's knot schedule: single / many throttle blocks, control horizon of one knot,
# held block starting at knot 0 or at the last knot, shortest fine grid, longest supported horizon
EDGE = [
    dict(nIter=10, nIterSmall=2, controlHorizon=7),     # 6 throttle blocks, 2 fine knots
    dict(nIter=8, nIterSmall=2, controlHorizon=2),      # one throttle block, joint block 1 held over 7 knots
    dict(nIter=6, nIterSmall=2, controlHorizon=6),      # nothing held
    dict(nIter=32, nIterSmall=7, controlHorizon=12),    # longest horizon the condensed kernel covers
    dict(nIter=12, nIterSmall=3, controlHorizon=4, periodMPCLargeSteps=0.05, periodMPCSmallSteps=0.005),  # ratio 10
    dict(nIter=9, nIterSmall=4, controlHorizon=8),      # held block over the last two knots only
]


@pytest.mark.parametrize("variant", range(len(EDGE)))
def test_condensed_schedule_edge_cases(variant):
    params = EDGE[variant]
    B = 4
    syn, bat = pkg("synthetic"), pkg("batched")
    traj = load_trajectories()
    nom = syn.make_states(B, perturbed=False)
    per = syn.make_states(B, seed=57 + variant, perturbed=True, near_bound_fraction=0.5)
    mpc = bat.BatchedVSMPC(B, params, oracle_trajectories_to_product(traj), solver=0, full_solution=True)
    mpc.configure(nom)
    oracles = [OracleInstance(nom, i, params=params, trajectories=traj) for i in range(B)]
    ratio = oracles[0].mpc.vectorConstraints[2].ratio
    for tick in range(3):
        if tick == 2:       # force a released tick
            mpc.debug_set_counters(-1, ratio - 1)
        mpc.update(per)
        mpc.solveMPC()
        z = mpc.getSolution()
        out, status = mpc.get_output()
        assert (status == 0).all(), status
        for i, o in enumerate(oracles):
            if tick == 2:
                o.mpc.vectorConstraints[2].counter = ratio - 1
            o.update(per)
            zo = o.solve()
            hz = dict(N=o.params["nIter"], Nc=o.params["controlHorizon"], nblk=o.params["controlHorizon"] - o.params["nIterSmall"] + 1)
            assert_solution_close(z[i], zo, 1e-6, what=(variant, tick, i), **hz)     # per physical quantity
            assert_output_rows_close(out[i], o.output_row(), 1e-6, what=(variant, tick, i))
    mpc.close()


# long horizons on the default solver (several column warps): 4x knots with 35 and 21 throttle blocks, a fully
# controlled long horizon (no held block), pinned and released ticks
LONG = [
    dict(nIter=68, nIterSmall=14, controlHorizon=48),
    dict(nIter=68, nIterSmall=28, controlHorizon=48, periodMPCLargeSteps=0.2),
    dict(nIter=40, nIterSmall=10, controlHorizon=40),
    dict(nIter=36, nIterSmall=30, controlHorizon=36),     # 7 throttle blocks: a single column warp, nothing held
    dict(nIter=20, nIterSmall=2, controlHorizon=12),      # 11 throttle blocks behind the shortest fine grid
    dict(nIter=40, nIterSmall=12, controlHorizon=13),     # 2 throttle blocks, joint block 12 held over 27 knots
]


@pytest.mark.parametrize("variant", range(len(LONG)))
def test_long_horizon_default_solver(variant):
    params = LONG[variant]
    B = 3
    syn, bat = pkg("synthetic"), pkg("batched")
    traj = load_trajectories()
    nom = syn.make_states(B, perturbed=False)
    per = syn.make_states(B, seed=71 + variant, perturbed=True, near_bound_fraction=0.4)
    mpc = bat.BatchedVSMPC(B, params, oracle_trajectories_to_product(traj), solver=0, full_solution=True)
    mpc.configure(nom)
    oracles = [OracleInstance(nom, i, params=params, trajectories=traj) for i in range(B)]
    ratio = oracles[0].mpc.vectorConstraints[2].ratio
    for tick in range(2):
        if tick == 1:       # force a released tick
            mpc.debug_set_counters(-1, ratio - 1)
        mpc.update(per)
        mpc.solveMPC()
        z = mpc.getSolution()
        out, status = mpc.get_output()
        assert (status == 0).all(), status
        for i, o in enumerate(oracles):
            if tick == 1:
                o.mpc.vectorConstraints[2].counter = ratio - 1
            o.update(per)
            zo = o.solve()
            hz = dict(N=o.params["nIter"], Nc=o.params["controlHorizon"], nblk=o.params["controlHorizon"] - o.params["nIterSmall"] + 1)
            assert_solution_close(z[i], zo, 1e-6, what=(variant, tick, i), **hz)     # per physical quantity
            assert_output_rows_close(out[i], o.output_row(), 1e-6, what=(variant, tick, i))
    mpc.close()


@pytest.mark.gpu
@pytest.mark.parametrize("variant", [0, 2, 4])
def test_warm_started_working_set_any_guess_same_minimiser(variant):
    """The long-horizon kernel starts its active set from the working set of the instance's previous solve
    (tools/condensed_model.box_qp_pivot_warm).  A tick sequence that jumps between unrelated states — every guess is the
    working set of ANOTHER state — with a throttle release in the middle, then adversarial guesses written into the device
    state (every bound upper, every bound lower, all free, random): every solve lands on the oracle's minimiser, and the
    cold start (vsmpc_set_warm_start 0) gives the same rows with more pivots."""
    params = LONG[variant]
    B = 4
    syn, bat = pkg("synthetic"), pkg("batched")
    traj = load_trajectories()
    nom = syn.make_states(B, perturbed=False)
    states = [syn.make_states(B, seed=300 + 17 * k + variant, perturbed=True, near_bound_fraction=0.1 + 0.2 * (k % 4)) for k in range(4)]
    nblk = params["controlHorizon"] - params["nIterSmall"] + 1
    hz = dict(N=params["nIter"], Nc=params["controlHorizon"], nblk=nblk)
    rng = np.random.default_rng(5)
    guesses = {4: np.ones((B, 4 * nblk)), 5: -np.ones((B, 4 * nblk)), 6: np.zeros((B, 4 * nblk)),
               7: rng.integers(-1, 2, (B, 4 * nblk))}
    pivots = {}
    for warm in (True, False):
        mpc = bat.BatchedVSMPC(B, params, oracle_trajectories_to_product(traj), solver=0, full_solution=True)
        mpc.set_warm_start(warm)
        mpc.configure(nom)
        oracles = [OracleInstance(nom, i, params=params, trajectories=traj) for i in range(B)]
        ratio = oracles[0].mpc.vectorConstraints[2].ratio
        pivots[warm] = 0
        for tick in range(8):
            per = states[tick % len(states)]
            if tick == 3:       # a released tick: block 0 joins the variables, the stored set has no entry for it
                mpc.debug_set_counters(-1, ratio - 1)
            if warm and tick in guesses:
                mpc.debug_set_working_set(guesses[tick])
            mpc.update(per)
            mpc.solveMPC()
            z = mpc.getSolution()
            out, status = mpc.get_output()
            assert (status == 0).all(), (warm, tick, status)
            pivots[warm] += int(mpc.get_pivot_counts().sum())
            for i, o in enumerate(oracles):
                if tick == 3:
                    o.mpc.vectorConstraints[2].counter = ratio - 1
                o.update(per)
                zo = o.solve()
                assert_solution_close(z[i], zo, 1e-6, what=(variant, warm, tick, i), **hz)
                assert_output_rows_close(out[i], o.output_row(), 1e-6, what=(variant, warm, tick, i))
        mpc.close()
    print("exchange pivots over the sequence: warm", pivots[True], "cold", pivots[False])
