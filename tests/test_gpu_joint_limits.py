"""Optional joint-limit rows (BASELINE configs[4] "per-instance constraint sets"; north_star: "... and joint limits").

The reference carries a JointPositionConstraint (constraintsVSMPC.cpp:388-468) that the shipped problem never registers
(variableSamplingMPC.cpp:77-84) and whose parameters jointPos_max / jointPos_min are absent from the XML.  Offered here as an
extension: 8 * nIter rows after the throttle rows, block i < controlHorizon bounding dq_i by limits - q_cmd (:450-453) for
every block (the m_firstIteriation slip of :440-449 is fixed, not ported).  The oracle registers the same class when the two
parameters are given.  At the reference horizon the QP kernel carries the joint boxes itself: a primal-dual working set on
the 8 x controlHorizon boxes, clamped increments held as constants inside the 8 x 8 eliminations, one more factorisation per
change of the working set (2-4 in all when bounds are active, tools/condensed_model.py ClampedCondensedQP); the long-horizon
kernel does the same up to twice the reference knot count.  The KKT fallback kernel, whose active set can carry the joint boxes
too, is the net behind a working set that does not settle.
  * CPU: the oracle's rows, bounds and minimiser (KKT certificate of the exact solver, boxes respected, bounds active);
  * GPU: gradient / bounds / constraint matrix against the oracle's assembly, the minimiser and every output field per
    physical quantity (1e-6 relative) on a workload where well over 10 % of the instances have an active joint bound,
    handle-wide and per-instance limits, a 2x horizon (wide kernel hand-over), a closed tick sequence (accumulated q_cmd)."""
import numpy as np
import pytest

from helpers import assert_output_rows_close, assert_solution_close, kkt_certificate, load_trajectories, pkg, split_hessian
from oracle_driver import OracleInstance, oracle_trajectories_to_product

# degrees, around the synthetic commanded posture (shoulder pitch / roll / yaw, elbow; left then right)
JMIN = [-30.0, 5.0, 20.0, 5.0, -30.0, 5.0, 20.0, 5.0]
JMAX = [-8.0, 30.0, 42.0, 28.0, -8.0, 30.0, 42.0, 28.0]
LIMITS = dict(jointPos_min=JMIN, jointPos_max=JMAX)
SEL = list(range(3, 11))


def n_active_joint_bounds(z, lo, hi, N, Nc, tol=1e-9):
    dq = z[26 * (N + 1):26 * (N + 1) + 8 * Nc].reshape(Nc, 8)
    return int(((dq <= lo + tol) | (dq >= hi - tol)).sum())


def test_oracle_joint_limit_rows_cpu():
    syn = pkg("synthetic")
    traj = load_trajectories()
    nom = syn.make_states(2, perturbed=False)
    per = syn.make_states(2, seed=5, perturbed=True, near_bound_fraction=0.3)
    o = OracleInstance(nom, 0, params=LIMITS, trajectories=traj)
    free = OracleInstance(nom, 0, trajectories=traj)
    o.update(per)
    free.update(per)
    m = o.mpc
    assert m.nConstraints == free.mpc.nConstraints + 8 * 17          # nJoints * nIter rows (:391)
    z = o.solve()
    zf = free.solve()
    lo = np.radians(JMIN) - per["q_cmd"][0][SEL]
    hi = np.radians(JMAX) - per["q_cmd"][0][SEL]
    dq = z[26 * 18:26 * 18 + 96].reshape(12, 8)
    assert (dq >= lo - 1e-9).all() and (dq <= hi + 1e-9).all()
    assert n_active_joint_bounds(z, lo, hi, 17, 12) > 0
    assert np.abs(z - zf).max() > 1e-3                               # the rows change the minimiser
    # rows: identity on dq block i for i < controlHorizon, zero rows with zero bounds after (sized nIter blocks)
    A = m.linearMatrix[-8 * 17:]
    for i in range(17):
        blk = A[8 * i:8 * i + 8]
        if i < 12:
            assert np.array_equal(blk[:, 26 * 18 + 8 * i:26 * 18 + 8 * i + 8], np.eye(8)) and np.count_nonzero(blk) == 8
            assert np.allclose(m.lowerBound[-8 * 17:][8 * i:8 * i + 8], lo, atol=0, rtol=1e-15)
        else:
            assert np.count_nonzero(blk) == 0 and not m.lowerBound[-8 * 17:][8 * i:8 * i + 8].any()


def test_clamped_recursion_specification_and_certificate_cpu():
    """tools/condensed_model.ClampedCondensedQP — the NumPy specification of the QP kernel's working set on the joint boxes —
    against the oracle's exact solver, and tests/helpers.kkt_certificate with the joint boxes on the oracle's minimisers."""
    import os, sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    from condensed_model import ClampedCondensedQP
    import oracle.vsmpc_oracle as O
    syn = pkg("synthetic")
    traj = load_trajectories()
    B = 4
    nom = syn.make_states(B, perturbed=False)
    per = syn.make_states(B, seed=5, perturbed=True, near_bound_fraction=0.3)
    p = O.default_params()
    zs, As, BJs, BTs, qs, ls, us, los, his = [], [], [], [], [], [], [], [], []
    n_clamped = 0
    for i in range(B):
        o = OracleInstance(nom, i, params=LIMITS, trajectories=traj, phase0=(0 if i % 2 else 19))
        o.update(per)
        z = o.solve().copy()
        cs, rt, tc, jl = o.mpc.vectorConstraints[0], o.mpc.vectorCosts[0], o.mpc.vectorConstraints[2], o.mpc.vectorConstraints[3]
        jm = o.qp.getJetModel()
        vbar = np.array([jm.compute_v(jm.standardizeThrottle_u2T(u)) for u in o.qp.getThrottleMPC()])
        pinned = not np.array_equal(o.mpc.lowerBound[468:472], np.full(4, tc.vMin))
        m = ClampedCondensedQP(cs.A, cs.BJ, cs.BT, cs.c, cs.dt, np.diag(rt.Q).copy(), rt.stateReference.T.copy(),
                               np.array(p["weightDeltaJoint"]) + p["weightRegularizationJointPos"],
                               o.mpc.vectorCosts[3].gradient[468:476].copy(), p["weightThrottle"], p["weightInitialThrottle"],
                               vbar, pinned, tc.vMin, tc.vMax, o.mpc.vectorConstraints[1].initialState, 17, 7, 12)
        lo, hi = jl.lowerBound[:8].copy(), jl.upperBound[:8].copy()
        x, dq, v, passes, clamp = m.solve_boxes(lo, hi)
        assert 1 <= passes <= 6 and m.status == 0
        assert np.abs(m.pack_z(x, dq, v) - z).max() / max(1.0, np.abs(z).max()) < 1e-10
        n_clamped += int((clamp != 0).sum())
        A, BJ, BT, c, dt = o.dynamics()
        zs.append(z); As.append(A.copy()); BJs.append(BJ.copy()); BTs.append(BT.copy()); los.append(lo); his.append(hi)
        qs.append(o.mpc.gradient.copy()); ls.append(o.mpc.lowerBound.copy()); us.append(o.mpc.upperBound.copy())
        H = o.mpc.hessian
    assert n_clamped > 0
    Pd, w_t = split_hessian(H, 17, 12)
    arr = np.array
    k = kkt_certificate(arr(zs), arr(As), arr(BJs), arr(BTs), dt, arr(qs), arr(ls), arr(us), Pd, w_t, dq_lo=arr(los), dq_hi=arr(his))
    assert k["stationarity_dq"].max() < 1e-10 and k["dual_sign_dq"].max() < 1e-10 and k["box_dq"].max() < 1e-12
    assert k["complementarity"].max() < 1e-10 and k["dual_sign"].max() < 1e-10 and k["n_dq_at_bound"] == n_clamped
    # the same points against boxes that are a little wider: the increments on the old bounds are now strictly inside with a
    # non-zero residual, which the certificate must report
    kb = kkt_certificate(arr(zs), arr(As), arr(BJs), arr(BTs), dt, arr(qs), arr(ls), arr(us), Pd, w_t,
                         dq_lo=arr(los) - 1e-3, dq_hi=arr(his) + 1e-3)
    assert kb["stationarity_dq"].max() > 1e-6


def _compare_tick(mpc, oracles, per, N, Nc, nblk, what):
    mpc.update(per)
    q, l, u = mpc.get_qp_vectors()
    mpc.solveMPC()
    z = mpc.getSolution()
    out, status = mpc.get_output()
    assert (status == 0).all(), (what, status)
    n_act = 0
    for i, o in enumerate(oracles):
        o.update(per)
        zo = o.solve()
        m = o.mpc
        assert np.abs(q[i] - m.gradient).max() <= 1e-12 * max(1.0, np.abs(m.gradient).max())
        assert np.abs(l[i] - m.lowerBound).max() <= 1e-12 * max(1.0, np.abs(m.lowerBound).max()), what
        assert np.abs(u[i] - m.upperBound).max() <= 1e-12 * max(1.0, np.abs(m.upperBound).max()), what
        assert_solution_close(z[i], zo, 1e-6, N=N, Nc=Nc, nblk=nblk, what=(what, i))
        assert_output_rows_close(out[i], o.output_row(), 1e-6, what=(what, i))
        jl = m.vectorConstraints[3]
        lo, hi = jl.lowerBound[:8], jl.upperBound[:8]
        n_act += n_active_joint_bounds(zo, lo, hi, N, Nc) > 0
    return n_act


@pytest.mark.gpu
@pytest.mark.parametrize("horizon", [None, dict(nIter=34, nIterSmall=14, controlHorizon=24)])
def test_joint_limit_rows_match_oracle(horizon):
    B = 10
    syn, bat = pkg("synthetic"), pkg("batched")
    traj = load_trajectories()
    params = dict(LIMITS)
    params.update(horizon or {})
    nom = syn.make_states(B, perturbed=False)
    mpc = bat.BatchedVSMPC(B, params, oracle_trajectories_to_product(traj), full_solution=True)
    mpc.configure(nom)
    mpc.set_fallback(0)           # both kernels carry the joint boxes themselves (working set + re-factorisation): no KKT fallback
    oracles = [OracleInstance(nom, i, params=params, trajectories=traj) for i in range(B)]
    p = oracles[0].params
    N, Nc, nblk = p["nIter"], p["controlHorizon"], p["controlHorizon"] - p["nIterSmall"] + 1
    assert mpc.n_con == oracles[0].mpc.nConstraints
    total = 0
    for tick in range(3):
        per = syn.make_states(B, seed=90 + tick, perturbed=True, near_bound_fraction=0.3)
        if tick == 2:
            mpc.debug_set_counters(-1, oracles[0].mpc.vectorConstraints[2].ratio - 1)      # a released tick
            for o in oracles:
                o.mpc.vectorConstraints[2].counter = o.mpc.vectorConstraints[2].ratio - 1
        total += _compare_tick(mpc, oracles, per, N, Nc, nblk, ("tick", tick))
        if tick == 0:
            A, Ao = mpc.getLinearConstraintMatrix(1), oracles[1].mpc.linearMatrix
            assert np.array_equal(A != 0, Ao != 0) and np.abs(A - Ao).max() <= 1e-12 * np.abs(Ao).max()
    assert total >= 0.1 * 3 * B, total       # well over 10 % of the solves have an active joint bound
    mpc.close()


@pytest.mark.gpu
def test_per_instance_joint_limits_match_oracle():
    """vsmpc_set_joint_limits: every instance its own box (some wide open, some tight around the commanded posture)."""
    B = 12
    syn, bat = pkg("synthetic"), pkg("batched")
    traj = load_trajectories()
    nom = syn.make_states(B, perturbed=False)
    per = syn.make_states(B, seed=17, perturbed=True, near_bound_fraction=0.3)
    rng = np.random.default_rng(3)
    width = np.where(np.arange(B)[:, None] % 3 == 0, 3.0, rng.uniform(0.02, 0.3, (B, 8)))   # rad; every third instance open
    qc = nom["q_cmd"][:, SEL]
    lo, hi = qc - width, qc + width * rng.uniform(0.5, 1.5, (B, 8))
    mpc = bat.BatchedVSMPC(B, LIMITS, oracle_trajectories_to_product(traj), full_solution=True)
    mpc.set_joint_limits(lo, hi)
    mpc.configure(nom)
    oracles = []
    for i in range(B):
        pi = dict(jointPos_min=list(np.degrees(lo[i])), jointPos_max=list(np.degrees(hi[i])))
        oracles.append(OracleInstance(nom, i, params=pi, trajectories=traj))
    mpc.set_fallback(0)            # no KKT fallback: the passes below are the QP kernel's own
    n_act = _compare_tick(mpc, oracles, per, 17, 12, 6, "per-instance")
    nf, _ = mpc.get_counts()
    assert 2 <= n_act < B          # both kinds in one batch: no joint bound active (one factorisation) and a working set
    assert (nf >= 2).sum() >= n_act and (nf == 1).sum() >= 1 and nf.max() <= 6
    mpc.set_fallback(1)
    # back to the handle-wide limits
    mpc.set_joint_limits(None, None)
    oracles = [OracleInstance(nom, i, params=LIMITS, trajectories=traj) for i in range(B)]
    mpc.configure(nom)
    _compare_tick(mpc, oracles, per, 17, 12, 6, "handle-wide again")
    mpc.close()


@pytest.mark.gpu
def test_joint_limits_need_the_flag_and_the_default_solver():
    syn, bat = pkg("synthetic"), pkg("batched")
    traj = oracle_trajectories_to_product(load_trajectories())
    with pytest.raises(bat.VsmpcError):
        bat.BatchedVSMPC(2, LIMITS, traj, solver=2)
    mpc = bat.BatchedVSMPC(2, None, traj)
    with pytest.raises(bat.VsmpcError):
        mpc.set_joint_limits(np.zeros((2, 8)) - 1, np.zeros((2, 8)) + 1)
    mpc.close()
    with pytest.raises(bat.VsmpcError):
        bat.BatchedVSMPC(2, dict(jointPos_min=JMAX, jointPos_max=JMIN), traj)


@pytest.mark.gpu
def test_joint_boxes_inside_the_qp_kernel_kkt_certificate_at_1024_instances():
    """Reference horizon, B = 1024, joint-limit rows on, fallback kernel OFF: every instance is solved by the QP kernel's own
    working set on the joint boxes.  Solver-independent KKT certificate (stationarity, complementarity and multiplier signs of
    the throttle AND the joint boxes) for all instances + the oracle on a seeded sample."""
    B = 1024
    syn, bat, P = pkg("synthetic"), pkg("batched"), pkg("pack")
    traj = load_trajectories()
    nom = syn.make_states(B, perturbed=False)
    per = syn.make_states(B, seed=4242, perturbed=True, near_bound_fraction=0.3)
    mpc = bat.BatchedVSMPC(B, LIMITS, oracle_trajectories_to_product(traj), full_solution=True)
    phase0 = (np.arange(B) % 20).astype(np.int32)
    mpc.configure_pack(P.build_pack(nom), np.ascontiguousarray(nom["joint_pos"][:, P.DEFAULT_JOINT_SELECTOR].T), phase0)
    mpc.set_fallback(0)
    mpc.update(per)
    A, BJ, BT, c, dt = mpc.get_dynamics()
    q, l, u = mpc.get_qp_vectors()
    H = mpc.getHessian(0)
    mpc.solveMPC()
    z = mpc.getSolution()
    out, status = mpc.get_output()
    nf, ns = mpc.get_counts()
    mpc.close()
    assert (status == 0).all(), np.unique(status, return_counts=True)
    Pd, w_t = split_hessian(H, 17, 12)
    lo = np.radians(JMIN)[None, :] - per["q_cmd"][:, SEL]
    hi = np.radians(JMAX)[None, :] - per["q_cmd"][:, SEL]
    k = kkt_certificate(z, A, BJ, BT, dt, q, l, u, Pd, w_t, 17, 7, 12, dq_lo=lo, dq_hi=hi)
    assert k["stationarity_dq"].max() < 1e-9, k["stationarity_dq"].max()
    assert k["complementarity"].max() < 1e-9 and k["dual_sign"].max() < 1e-9
    assert k["dual_sign_dq"].max() < 1e-9, k["dual_sign_dq"].max()
    assert k["box"].max() < 1e-12 and k["box_dq"].max() <= 1e-9
    assert k["instances_dq_at_bound"] > B // 10 and k["n_dq_at_bound"] > B       # the workload exercises the joint boxes
    assert (nf[(ns == 1)] >= 1).all() and nf.max() <= 16 and (nf >= 2).sum() >= k["instances_dq_at_bound"]
    for i in np.random.default_rng(7).choice(B, 6, replace=False):
        o = OracleInstance(nom, int(i), params=LIMITS, trajectories=traj, phase0=int(phase0[i]))
        o.update(per)
        zo = o.solve()
        assert_solution_close(z[i], zo, 1e-6, N=17, Nc=12, nblk=6, what=("z", int(i)))
        assert_output_rows_close(out[i], o.output_row(), 1e-6, what=("row", int(i)))


@pytest.mark.gpu
def test_joint_box_working_set_warm_start_same_minimiser_fewer_factorisations():
    """The working set of the joint boxes a solve ended with is the next tick's first guess (vsmpc_set_warm_start): the same
    minimiser as a cold start, and on a repeated state hardly any re-factorisation."""
    B = 256
    syn, bat = pkg("synthetic"), pkg("batched")
    traj = oracle_trajectories_to_product(load_trajectories())
    nom = syn.make_states(B, perturbed=False)
    pers = [syn.make_states(B, seed=70 + j, perturbed=True, near_bound_fraction=0.3) for j in range(2)]
    runs = {}
    for warm in (0, 1):
        mpc = bat.BatchedVSMPC(B, LIMITS, traj, full_solution=True)
        mpc.configure(nom)
        mpc.set_fallback(0)
        mpc.set_warm_start(warm)
        zs, nfs = [], []
        for per in (pers[0], pers[0], pers[1], pers[1]):
            mpc.update(per)
            mpc.solveMPC()
            _, status = mpc.get_output()
            assert (status == 0).all()
            zs.append(mpc.getSolution().copy())
            nfs.append(mpc.get_counts()[0].copy())
        runs[warm] = (zs, nfs)
        mpc.close()
    for t in range(4):
        zc, zw = runs[0][0][t], runs[1][0][t]
        assert np.abs(zc - zw).max() <= 1e-9 * max(1.0, np.abs(zc).max()), t
    cold, warm = runs[0][1], runs[1][1]
    assert cold[1].mean() > 2.0                       # the workload needs its working set
    assert np.array_equal(cold[0], warm[0])           # first tick after configure: empty guess either way
    # the same state again: the guess is the answer up to the 20-tick phase of the throttle rows
    assert warm[1].mean() < 1.5 and warm[3].mean() < 1.5 and warm[1].mean() < 0.6 * cold[1].mean()


@pytest.mark.gpu
@pytest.mark.parametrize("hz", [(34, 14, 24), (28, 10, 20)])
def test_joint_boxes_inside_the_long_horizon_kernel_kkt_certificate(hz):
    """Twice the reference knot count and a horizon in between (the JL build of the long-horizon kernel, two column warps),
    joint-limit rows on, fallback kernel OFF: KKT certificate including the joint boxes for every instance."""
    N, Ns, Nc = hz
    B = 128
    syn, bat, P = pkg("synthetic"), pkg("batched"), pkg("pack")
    traj = load_trajectories()
    params = dict(LIMITS, nIter=N, nIterSmall=Ns, controlHorizon=Nc)
    nom = syn.make_states(B, perturbed=False)
    per = syn.make_states(B, seed=515 + N, perturbed=True, near_bound_fraction=0.3)
    mpc = bat.BatchedVSMPC(B, params, oracle_trajectories_to_product(traj), full_solution=True)
    phase0 = (np.arange(B) % 20).astype(np.int32)
    mpc.configure_pack(P.build_pack(nom), np.ascontiguousarray(nom["joint_pos"][:, P.DEFAULT_JOINT_SELECTOR].T), phase0)
    mpc.set_fallback(0)
    for tick in range(2):                  # the second tick starts from the stored working sets (throttle and joint boxes)
        mpc.update(per)
        A, BJ, BT, c, dt = mpc.get_dynamics()
        q, l, u = mpc.get_qp_vectors()
        mpc.solveMPC()
        z = mpc.getSolution()
        _, status = mpc.get_output()
        nf, ns = mpc.get_counts()
        assert (status == 0).all(), (tick, np.unique(status, return_counts=True))
        if tick == 0:
            H = mpc.getHessian(0)
            Pd, w_t = split_hessian(H, N, Nc)
            nf0 = nf.copy()
        # the commanded posture the rows are written against moves with the accumulated joint references: bounds from the library
        nrow_thr = l.shape[1] - 26 * N - 26 - 8 * N
        lo, hi = l[:, 26 * N + 26 + nrow_thr:][:, :8], u[:, 26 * N + 26 + nrow_thr:][:, :8]
        k = kkt_certificate(z, A, BJ, BT, dt, q, l, u, Pd, w_t, N, Ns, Nc, dq_lo=lo, dq_hi=hi)
        assert k["stationarity_dq"].max() < 1e-9 and k["dual_sign_dq"].max() < 1e-9 and k["box_dq"].max() <= 1e-9, (tick, k)
        assert k["complementarity"].max() < 1e-9 and k["dual_sign"].max() < 1e-9 and k["box"].max() < 1e-12
        assert k["instances_dq_at_bound"] > B // 10
    assert nf0.max() <= 24 and (nf0 >= 2).sum() > B // 10 and nf.mean() < nf0.mean()
    mpc.close()
