"""CPU tests pinning the oracle (oracle/vsmpc_oracle.py) — no GPU needed.

The reference ships no tests for this path, so the oracle is pinned against: the reference's own
fixtures (trajectory files), the second statement of the jet model in
/root/reference/src/mujoco_lib/jet_kalman_filter.py:6-45, closed-form identities of SURVEY App. A,
a KKT certificate of the exact solver, and frozen golden vectors (tests/golden/make_golden.py).
"""
import math

import numpy as np
import pytest

from helpers import golden, load_trajectories, pkg, state_from_pack
from oracle import vsmpc_oracle as O
from oracle_driver import OracleInstance


# ---- jet model ------------------------------------------------------------------------------------
def ekf_statement_Tddot(T, Tdot, u):
    """Second statement of the jet model: src/mujoco_lib/jet_kalman_filter.py:6-45 (T_dot_dot * std_thrust)."""
    c = [-4.64730485e-01, -8.13171858e+00, -6.19539230e+00, 6.61113140e-01, 1.67673231e+00, -4.83287064e-01,
         8.77996617e+00, -1.01096376e+00, -5.86442286e-01, 5.19093322e-01, -4.23782666e-01, -1.45705257e+00,
         -7.83052261e-03]
    mean_thrust, std_thrust, mean_throttle, std_throttle = 108.309, 65.793, 47.333, 31.483
    ts, tds, us = (T - mean_thrust) / std_thrust, Tdot / std_thrust, (u - mean_throttle) / std_throttle
    f = c[0] + c[1] * ts + c[2] * tds + c[3] * ts * tds + c[4] * ts ** 2 + c[5] * tds ** 2
    g = c[6] + c[7] * ts + c[8] * tds + c[9] * ts * tds + c[10] * ts ** 2 + c[11] * tds ** 2
    v = us + c[12] * us ** 2
    return (f + g * v) * std_thrust


def test_jet_model_matches_second_statement():
    jm = O.JetModel()
    rng = np.random.default_rng(1)
    qp = O.QPInput()
    qp.setEmptyJetModel()
    jd = O.JetDynamicVS(26, 8, 4)
    jd.jetModel = jm
    for _ in range(50):
        T, Td, u = rng.uniform(20, 220), rng.normal(0, 30), rng.uniform(0, 100)
        v = jm.compute_v(jm.standardizeThrottle_u2T(u))
        h = jd.computeF(T, Td) + jd.computeG(T, Td) * v
        assert h == pytest.approx(ekf_statement_Tddot(T, Td, u), rel=1e-13)
        # Jacobians of h used in A (systemDynamicsVSMPC.cpp:410-413) vs central differences of the 2nd statement
        e = 1e-4
        dT = (ekf_statement_Tddot(T + e, Td, u) - ekf_statement_Tddot(T - e, Td, u)) / (2 * e)
        dTd = (ekf_statement_Tddot(T, Td + e, u) - ekf_statement_Tddot(T, Td - e, u)) / (2 * e)
        assert jd.compute_dh_dT(T, Td, u) == pytest.approx(dT, rel=1e-6, abs=1e-8)
        assert jd.compute_dh_dTDot(T, Td, u) == pytest.approx(dTd, rel=1e-6, abs=1e-8)


def _oracle_jet_tables(g):
    j = O.JetModel()
    poly = np.array([[getattr(j, str(n))(a, b) for a, b in zip(g["T_std"], g["Tdot_std"])] for n in g["poly_names"]])
    fn = {"destandardizeThrust_u2T": lambda x: x * j.n[1] + j.n[0], "destandardizeThrustDot_u2T": lambda x: x * j.n[1]}
    scal = np.array([[fn.get(str(n), getattr(j, str(n), None))(float(x)) for x in g["x"]] for n in g["scalar_names"]])
    return j, poly, scal


def test_jet_model_matches_reference_object_code():
    """PINNED: the oracle's JetModel against outputs of the reference's own JetModel.cpp, compiled from
    /root/reference by oracle/build_ref.py and frozen by tests/golden/make_jet_model_golden.py (rows a6 / a17:
    f, g, the four partial derivatives, standardisations, destandardizeThrottle incl. both clips)."""
    g = golden("jet_model_ref.npz")
    j, poly, scal = _oracle_jet_tables(g)
    np.testing.assert_allclose(poly, g["poly"], rtol=1e-14, atol=1e-14)
    np.testing.assert_allclose(scal, g["scalar"], rtol=1e-14, atol=1e-14, equal_nan=True)
    assert j.getThrustStandardDeviation_u2T() == float(g["thrust_std"])
    # destandardizeThrottle clips to [0, 100] % and is NaN beyond the root of its discriminant (v > 31.9), on both sides
    u = scal[list(g["scalar_names"]).index("destandardizeThrottle_u2T")]
    assert np.nanmin(u) == 0.0 and np.nanmax(u) == 100.0 and np.isnan(u).sum() == 2


def test_golden_jet_vectors_are_what_the_compiled_reference_returns():
    """Where oracle/_ref/libjetmodel_ref.so exists (built from the reference source in the build container; it travels
    to the GPU box), the frozen vectors are re-derived from it bit for bit."""
    import importlib.util
    import os
    import sys
    from oracle import build_ref
    lib = build_ref.load()
    if lib is None:
        pytest.skip("oracle/_ref/libjetmodel_ref.so not built (no /root/reference here)")
    spec = importlib.util.spec_from_file_location(
        "make_jet_model_golden", os.path.join(os.path.dirname(__file__), "golden", "make_jet_model_golden.py"))
    m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
    g = golden("jet_model_ref.npz")
    poly, scal, sigma = m.evaluate(lib, g["T_std"], g["Tdot_std"], g["x"])
    assert np.array_equal(poly, g["poly"]) and np.array_equal(scal, g["scalar"], equal_nan=True)
    assert sigma == float(g["thrust_std"])


def test_throttle_transform_roundtrip_and_bounds():
    jm = O.JetModel()
    for u in np.linspace(0, 100, 41):
        assert jm.destandardizeThrottle_u2T(jm.compute_v(jm.standardizeThrottle_u2T(u))) == pytest.approx(u, abs=1e-10)
    # v(std(0)), v(std(100))
    assert jm.compute_v(jm.standardizeThrottle_u2T(0.0)) == pytest.approx(-1.5211460, abs=1e-6)
    assert jm.compute_v(jm.standardizeThrottle_u2T(100.0)) == pytest.approx(1.6509574, abs=1e-6)
    assert jm.destandardizeThrottle_u2T(5.0) == 100.0 and jm.destandardizeThrottle_u2T(-5.0) == 0.0  # clipping


# ---- time grid, dimensions, fixtures ------------------------------------------------------------------
def test_time_grid():
    dt = O.time_grid(O.default_params())
    assert len(dt) == 17
    assert dt[:7].sum() == pytest.approx(0.1, abs=1e-15)     # the fine part is exactly one coarse step
    assert dt.sum() == pytest.approx(1.1, abs=1e-14)
    assert dt[0] == pytest.approx(0.005) and dt[6] == pytest.approx(0.0235714285714, rel=1e-9)
    assert (dt[7:] == 0.1).all()


def test_fixtures_match_survey_appendix_d():
    t = load_trajectories()
    a = t["TRAJECTORY_MANAGER"]["arrays"]["alphaGravity"]
    assert a.shape == (1, 351) and t["TRAJECTORY_MANAGER"]["fps"] == 10
    assert (a[0, :21] == 0.08).all() and a[0, 21] > 0.08 and (a[0, 200:] == 1.0).all()
    p = t["POSITION_TRAJECTORY"]["arrays"]
    for k in ("positionCoM", "velocityCoM", "RPY", "RPYDot"):
        assert p[k].shape == (3, 1481)
    assert (p["RPY"] == 0).all() and (p["RPYDot"] == 0).all()
    assert (p["positionCoM"][:, :201] == 0).all() and p["positionCoM"][2, 201] > 0
    assert p["positionCoM"][:, -1] == pytest.approx([0, 0, 2.5], abs=1e-12)
    # upsampling drops the last sample: (351-1)*20 = 7000 samples at 200 Hz (TrajectoryManager.cpp:27-35)
    tm = O.TrajectoryManager({"alphaGravity": a}, 10, 1 / 0.005)
    assert tm.trajectorySize == 7000
    assert tm.trajectories["alphaGravity"][0, 21 * 20 + 10] == pytest.approx(0.5 * (a[0, 21] + a[0, 22]))
    tm2 = O.TrajectoryManager(p, 10, 1 / 0.1)
    assert tm2.trajectorySize == 1481   # same fps: no resampling


@pytest.fixture(scope="module")
def small_batch():
    syn = pkg("synthetic")
    B = 4
    return syn.make_states(B, perturbed=False), syn.make_states(B, perturbed=True, near_bound_fraction=0.5)


def test_dimensions_and_structure(small_batch):
    nom, per = small_batch
    o = OracleInstance(nom, 0)
    o.update(per)
    m = o.mpc
    assert m.nVar == 588 and m.nConstraints == 512
    assert (np.abs(m.linearMatrix).sum(axis=1) == 0).sum() == 20       # rows 492..511 are identically zero
    assert (np.abs(m.linearMatrix[492:]).sum() == 0) and (m.lowerBound[492:] == 0).all()
    P = m.hessian
    assert np.allclose(P, P.T) and np.count_nonzero(P) == 466
    assert np.count_nonzero(P[:26, :26]) == 0                            # x_0 has no cost
    A, BJ, BT, c, dt = o.dynamics()
    assert np.count_nonzero(BT) == 4 and np.count_nonzero(BJ) <= 48
    assert (A[12:16, 16:20] == np.eye(4)).all() and (A[20:23, 0:3] == np.eye(3)).all()


def test_tick_phase_semantics(small_batch):
    """configure() is tick 0 of every counter: block 0 pinned on ticks 1-19, free on 20."""
    nom, per = small_batch
    o = OracleInstance(nom, 1)
    jm = o.qp.getJetModel()
    ref_idx0 = o.mpc.vectorCosts[0].trajManager.trajectoryIndex
    assert ref_idx0 == 1                                  # the cursor advanced once during configure
    alpha_idx0 = o.mpc.vectorConstraints[0].systemDynamicVS.vectorDynamic[1].trajectoryManager.trajectoryIndex
    assert alpha_idx0 == 1
    for tick in range(1, 42):
        o.update(per)
        l, u = o.mpc.lowerBound[468:472], o.mpc.upperBound[468:472]
        vbar = [jm.compute_v(jm.standardizeThrottle_u2T(x)) for x in per["throttle_prev"][1]]
        if tick % 20 == 0:
            assert (l < u).all()                          # released
            assert o.mpc.vectorCosts[0].trajManager.trajectoryIndex == ref_idx0 + tick // 20
        else:
            assert np.allclose(l, vbar) and np.allclose(u, vbar)
        assert (o.mpc.lowerBound[472:492] == o.mpc.vectorConstraints[2].vMin).all()
    assert o.mpc.vectorConstraints[0].systemDynamicVS.vectorDynamic[1].trajectoryManager.trajectoryIndex == alpha_idx0 + 41


def test_exact_solver_certificate_and_riccati_model(small_batch):
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    from riccati_model import RiccatiQP
    nom, per = small_batch
    p = O.default_params()
    for i in range(4):
        for free in (False, True):
            o = OracleInstance(nom, i)
            if free:
                o.mpc.vectorConstraints[2].counter = 19
            o.update(per)
            z = o.solve()
            info = o.mpc.solveInfo
            assert info["stationarity"] < 1e-8 and info["primal"] < 1e-10 and info["dual_sign"] == 0.0
            cs, rt = o.mpc.vectorConstraints[0], o.mpc.vectorCosts[0]
            jm = o.qp.getJetModel()
            vbar = np.array([jm.compute_v(jm.standardizeThrottle_u2T(u)) for u in o.qp.getThrottleMPC()])
            tc = o.mpc.vectorConstraints[2]
            m = RiccatiQP(cs.A, cs.BJ, cs.BT, cs.c, cs.dt, np.diag(rt.Q).copy(), rt.stateReference.T.copy(),
                          np.array(p["weightDeltaJoint"]) + p["weightRegularizationJointPos"],
                          o.mpc.vectorCosts[3].gradient[468:476].copy(), p["weightThrottle"],
                          p["weightInitialThrottle"], vbar, not free, tc.vMin, tc.vMax,
                          o.mpc.vectorConstraints[1].initialState, 17, 7, 12)
            zm = m.pack_z(*m.solve_box())
            assert m.status == 0 and len(m.active) == info["n_active"]
            assert np.abs(zm - z).max() / max(1.0, np.abs(z).max()) < 1e-10
            # the specification of the default kernel (condensed-throttle Riccati, tools/condensed_model.py)
            from condensed_model import CondensedQP
            mc = CondensedQP(cs.A, cs.BJ, cs.BT, cs.c, cs.dt, np.diag(rt.Q).copy(), rt.stateReference.T.copy(),
                             np.array(p["weightDeltaJoint"]) + p["weightRegularizationJointPos"],
                             o.mpc.vectorCosts[3].gradient[468:476].copy(), p["weightThrottle"],
                             p["weightInitialThrottle"], vbar, not free, tc.vMin, tc.vMax,
                             o.mpc.vectorConstraints[1].initialState, 17, 7, 12)
            zc = mc.pack_z(*mc.solve_box())
            assert mc.status == 0 and len(mc.active) == info["n_active"]
            assert np.abs(zc - z).max() / max(1.0, np.abs(z).max()) < 1e-10


# ---- frozen golden vectors ---------------------------------------------------------------------------
def test_oracle_reproduces_golden_single_tick():
    g = golden("golden_qp.npz")
    nom = state_from_pack(g["nom_pack"], g["joint_pos_sel"])
    per = state_from_pack(g["per_pack"])
    for tag, free in (("pin", False), ("free", True)):
        for i in range(nom["wRb"].shape[0]):
            o = OracleInstance(nom, i)
            if free:
                o.mpc.vectorConstraints[2].counter = 19
            o.update(per)
            z = o.solve()
            A, BJ, BT, c, _ = o.dynamics()
            assert np.abs(A - g[f"{tag}_A"][i]).max() < 1e-12
            assert np.abs(BJ - g[f"{tag}_BJ"][i]).max() < 1e-10
            assert np.abs(o.mpc.gradient - g[f"{tag}_q"][i]).max() < 1e-7
            assert np.abs(z - g[f"{tag}_z"][i]).max() / np.abs(z).max() < 1e-9
            assert o.mpc.solveInfo["n_active"] == g[f"{tag}_n_active"][i]


def test_oracle_reproduces_golden_tick_sequence():
    g = golden("golden_ticks.npz")
    nom = state_from_pack(g["nom_pack"], g["joint_pos_sel"])
    B = nom["wRb"].shape[0]
    inst = [OracleInstance(nom, i) for i in range(B)]
    for t in range(g["packs"].shape[0]):
        st = state_from_pack(g["packs"][t])
        for i, o in enumerate(inst):
            o.update(st)
            o.solve()
            row = o.output_row()
            assert np.abs(row - g["rows"][t, i]).max() / max(1.0, np.abs(row).max()) < 1e-9, (t, i)


def test_pivot_active_set_spec_on_random_box_qps():
    """tools/condensed_model.box_qp_pivot (the active set of the condensed kernels: exchange pivots on the principal pivot
    transform) against the KKT conditions of random strictly convex box QPs, from unconstrained to almost fully
    saturated."""
    from condensed_model import box_qp_pivot, exchange_pivot
    rng = np.random.default_rng(7)
    # the pivot is an involution and pivoting every index inverts
    A = rng.normal(size=(9, 9)); H = A @ A.T + 9 * np.eye(9)
    T = H.copy()
    exchange_pivot(T, 3); exchange_pivot(T, 3)
    assert np.abs(T - H).max() < 1e-12
    for q in range(9):
        exchange_pivot(T, q)
    assert np.abs(T @ H - np.eye(9)).max() < 1e-12
    for n, scale in ((8, 0.2), (24, 1.0), (24, 30.0), (60, 5.0), (140, 50.0)):
        for rep in range(4):
            A = rng.normal(size=(n, n))
            H = A @ A.T / n + np.diag(rng.uniform(0.5, 2.0, n))
            g = scale * rng.normal(size=n)
            lo, up = -1.52115, 1.65096
            v, active, status = box_qp_pivot(H, g, lo, up, max_iter=6 * n)
            assert status == 0
            y = H @ v + g                      # stationarity: y_F = 0, y_a = -s_a lam_a with lam_a >= 0
            assert (v <= up + 1e-12).all() and (v >= lo - 1e-12).all()
            free = np.ones(n, dtype=bool)
            for i, sgn, lam in active:
                free[i] = False
                assert v[i] == (up if sgn > 0 else lo) and lam >= -1e-9
                assert abs(y[i] + sgn * lam) < 1e-8 * max(1.0, abs(y[i]))
            assert np.abs(y[free]).max(initial=0.0) < 1e-8 * max(1.0, np.abs(g).max())


def test_warm_started_pivot_active_set_spec():
    """tools/condensed_model.box_qp_pivot_warm (specification of the warm start planned for the long-horizon kernel: start
    from a guessed working set, T = H and only the guessed-free indices pivoted, drop phase, then the dual iterations):
    whatever the guess — exact, none, a vertex, noisy, random — it returns the minimiser of the cold method, and an exact
    guess of an almost saturated problem costs a small fraction of the cold method's pivots."""
    from condensed_model import box_qp_pivot, box_qp_pivot_warm
    rng = np.random.default_rng(11)
    lo, up = -1.52115, 1.65096
    saved = []
    for n, scale, shift in ((6, 0.3, 0.0), (24, 1.0, 0.0), (44, 3.0, 2.0), (92, 10.0, -3.0), (140, 30.0, 20.0)):
        for rep in range(3):
            A = rng.normal(size=(n, n + 3))
            H = A @ A.T / (n + 3) + rng.uniform(1e-3, 0.2) * np.eye(n)
            g = scale * rng.normal(size=n) + shift
            v, active, status = box_qp_pivot(H, g, lo, up, max_iter=20 * n)
            assert status == 0
            exact = np.zeros(n, dtype=int)
            for i, sgn, _ in active:
                exact[i] = int(sgn)
            guesses = dict(exact=exact, none=np.zeros(n, dtype=int), all_lo=-np.ones(n, dtype=int), all_up=np.ones(n, dtype=int),
                           noisy=np.where(rng.random(n) < 0.15, rng.integers(-1, 2, n), exact), random=rng.integers(-1, 2, n))
            for name, a0 in guesses.items():
                w, aw, sw, pivots = box_qp_pivot_warm(H, g, lo, up, a0, max_iter=40 * n)
                assert sw == 0, (n, name)
                assert np.abs(w - v).max() < 1e-8, (n, name)
                assert {(i, int(s)) for i, s, _ in aw} == {(i, int(s)) for i, s, _ in active}, (n, name)
                assert all(lam >= -1e-9 for _, _, lam in aw)
                if name == "exact":
                    assert pivots == n - len(active)          # only the free indices are pivoted, nothing else happens
                    saved.append((n + len(active), pivots))
                if name == "none":
                    assert pivots >= n                        # = the cold method (inverse, then the dual iterations)
    cold, warm = (sum(x) for x in zip(*saved))
    assert warm < 0.5 * cold


def test_kkt_certificate_helper_on_oracle_solutions():
    """tests/helpers.kkt_certificate (used at the bench size on the GPU, where the oracle is too slow for all 1024
    instances) accepts the oracle's exact minimisers and rejects perturbed ones."""
    from helpers import kkt_certificate, split_hessian
    syn = pkg("synthetic")
    B = 6
    traj = load_trajectories()
    nom = syn.make_states(B, perturbed=False)
    per = syn.make_states(B, seed=5, perturbed=True, near_bound_fraction=0.5)
    zs, As, BJs, BTs, qs, ls, us = [], [], [], [], [], [], []
    for i in range(B):
        o = OracleInstance(nom, i, trajectories=traj, phase0=(0 if i % 2 else 19))    # pinned and released ticks
        o.update(per)
        zs.append(o.solve().copy())
        A, BJ, BT, c, dt = o.dynamics()
        As.append(A.copy()); BJs.append(BJ.copy()); BTs.append(BT.copy())
        qs.append(o.mpc.gradient.copy()); ls.append(o.mpc.lowerBound.copy()); us.append(o.mpc.upperBound.copy())
        H = o.mpc.hessian
    Pd, w_t = split_hessian(H, 17, 12)
    assert w_t == 80000.0
    arr = lambda x: np.array(x)
    k = kkt_certificate(arr(zs), arr(As), arr(BJs), arr(BTs), dt, arr(qs), arr(ls), arr(us), Pd, w_t)
    assert k["stationarity_dq"].max() < 1e-10 and k["complementarity"].max() < 1e-10 and k["dual_sign"].max() < 1e-10
    assert k["box"].max() < 1e-12 and k["n_at_bound"] > 0 and k["n_inside"] > 0
    # a feasible but non-optimal point: shift one joint-increment block (the states follow the dynamics)
    z_bad = arr(zs).copy()
    z_bad[:, 26 * 18 + 3] += 1e-4
    kb = kkt_certificate(z_bad, arr(As), arr(BJs), arr(BTs), dt, arr(qs), arr(ls), arr(us), Pd, w_t)
    assert kb["stationarity_dq"].min() > 1e-6
    # an interior throttle variable moved: complementarity is violated
    z_bad = arr(zs).copy()
    z_bad[:, 26 * 18 + 96 + 20] += 1e-3
    kb = kkt_certificate(z_bad, arr(As), arr(BJs), arr(BTs), dt, arr(qs), arr(ls), arr(us), Pd, w_t)
    assert max(kb["complementarity"].max(), kb["dual_sign"].max()) > 1e-6


def test_block_exchange_pivot_equals_two_single_pivots():
    """tools/condensed_model.block_exchange_pivot (the two-indices-per-step inverse of the reference-horizon kernel)."""
    from condensed_model import block_exchange_pivot, exchange_pivot
    g = np.random.default_rng(12)
    for n in (8, 20, 24):
        A = g.normal(size=(n, n))
        H = A @ A.T + n * np.eye(n)
        T1, T2 = H.copy(), H.copy()
        for p in range(0, n, 2):
            exchange_pivot(T1, p)
            exchange_pivot(T1, p + 1)
            assert block_exchange_pivot(T2, p, p + 1) > 0
            assert np.abs(T1 - T2).max() <= 1e-12 * np.abs(T1).max()
        assert np.abs(T2 - np.linalg.inv(H)).max() <= 1e-12 * np.abs(T2).max()
