"""The oracle PINNED against the reference's own code.

oracle/build_ref.py compiles the reference's VariableSamplingMPC — its own 13 translation units of the hot path, read where
they lie under /root/reference — against stand-in headers for the absent third-party libraries (oracle/ref_stubs/: a small
eager Eigen, iDynTree value types, YARP logging, the BLF parameter table, an OsqpEigen::Solver that records the QP and
forwards it to an installable solve function, matio over in-memory arrays).  tests/golden/make_reference_golden.py froze
what that library returns in tests/golden/reference_{qp,ticks}.npz; these tests hold the oracle to those vectors (everywhere)
and to the live library (where oracle/_ref/libvsmpc_reference.so exists: the build container and, since the .so travels,
the GPU box).  tests/test_gpu_golden.py holds the CUDA path to the same files."""
import numpy as np
import pytest

from helpers import golden, load_trajectories, pkg, state_from_pack
from oracle_driver import OracleInstance


def rel(a, b):
    return float(np.abs(np.asarray(a) - np.asarray(b)).max() / max(1.0, np.abs(b).max()))


def test_oracle_assembles_the_reference_qp_and_solution():
    """First tick after configure, 8 perturbed instances: the dense QP as IMPCProblem holds it (Hessian, gradient, the
    constraint matrix entry by entry incl. its sparsity pattern, both bounds) at 1e-12, minimiser and outputs at 1e-9."""
    g = golden("reference_qp.npz")
    traj = load_trajectories()
    nom = state_from_pack(g["nom_pack"], g["joint_pos_sel"])
    per = state_from_pack(g["per_pack"])
    B = g["nom_pack"].shape[1]
    for i in range(B):
        o = OracleInstance(nom, i, trajectories=traj)
        o.update(per)
        m = o.mpc
        assert m.nVar == 588 and m.nConstraints == 512
        assert np.array_equal(np.flatnonzero(m.hessian), g["P_index"])
        assert rel(m.hessian.reshape(-1)[g["P_index"]], g["P_value"]) < 1e-12
        assert np.array_equal(np.flatnonzero(m.linearMatrix), g["pin_A_index"])
        assert rel(m.linearMatrix.reshape(-1)[g["pin_A_index"]], g["pin_A_value"][i]) < 1e-12
        assert rel(m.gradient, g["pin_q"][i]) < 1e-12
        assert rel(m.lowerBound, g["pin_l"][i]) < 1e-12 and rel(m.upperBound, g["pin_u"][i]) < 1e-12
        z = o.solve()
        assert rel(z, g["pin_z"][i]) < 1e-9
        assert rel(o.output_row(), g["pin_row"][i]) < 1e-9


def test_oracle_follows_the_reference_over_a_tick_sequence():
    """24 consecutive ticks with the driver's feedback of the outputs: 20-tick reference-window shift, throttle release,
    alpha_g cursor, RPY unwrapping through +-pi, joint accumulator — every row the reference returned."""
    g = golden("reference_ticks.npz")
    traj = load_trajectories()
    nom = state_from_pack(g["nom_pack"], g["joint_pos_sel"])
    B = g["nom_pack"].shape[1]
    inst = [OracleInstance(nom, i, trajectories=traj) for i in range(B)]
    for t in range(g["packs"].shape[0]):
        st = state_from_pack(g["packs"][t])
        for i, o in enumerate(inst):
            o.update(st)
            z = o.solve()
            assert rel(z, g["z"][t, i]) < 1e-9, (t, i)
            assert rel(o.output_row(), g["rows"][t, i]) < 1e-9, (t, i)


def test_oracle_made_golden_files_equal_the_reference_made_ones():
    """golden_qp / golden_ticks (frozen from the oracle earlier in the round, used by every GPU golden test) carry the same
    numbers as the files frozen from the reference."""
    a, b = golden("golden_qp.npz"), golden("reference_qp.npz")
    for k in ("nom_pack", "per_pack", "pin_z", "pin_row", "pin_q", "pin_l", "pin_u"):
        assert rel(a[k], b[k]) < 1e-12, k
    a, b = golden("golden_ticks.npz"), golden("reference_ticks.npz")
    for k in ("packs", "rows", "z"):
        assert rel(a[k], b[k]) < 1e-12, k


def _reference():
    import reference_driver
    if reference_driver.lib() is None:
        pytest.skip("oracle/_ref/libvsmpc_reference.so not built (no /root/reference here)")
    return reference_driver


def test_live_reference_against_oracle_on_fresh_states():
    """The compiled reference and the oracle side by side on states the golden files do not hold: 42 ticks (two throttle
    releases) with a new perturbed state every tick, the QP compared entry by entry on every tick."""
    rd = _reference()
    syn = pkg("synthetic")
    B = 2
    nom = syn.make_states(B, seed=5, perturbed=False)
    for i in range(B):
        o, r = OracleInstance(nom, i), rd.ReferenceInstance(nom, i)
        assert (r.n_var, r.n_con) == (588, 512)
        for t in range(42):
            st = syn.make_states(B, seed=9000 + t, perturbed=True, near_bound_fraction=0.5)
            o.update(st); r.update(st)
            m = o.mpc
            for name, x, y in zip("PqAlu", r.qp(), (m.hessian, m.gradient, m.linearMatrix, m.lowerBound, m.upperBound)):
                assert rel(y, x) < 1e-12, (name, t)
            zo, zr = o.solve(), r.solve()
            assert rel(zo, zr) < 1e-9, t
            out = r.output()
            assert out["status"] == 1
            assert rel(m.getThrottleReference(), out["throttle"]) < 1e-9
            assert rel(m.getThrustReference(), out["thrust"]) < 1e-9 and rel(m.getThrustDotReference(), out["thrust_dot"]) < 1e-9
            assert rel(m.getJointsReferencePosition(), out["joints"]) < 1e-9
            assert abs(o.qp.alphaGravity - out["qp_input"][0]) < 1e-15
            assert rel(o.qp.posCoMReference, out["qp_input"][1:4]) < 1e-12 and rel(o.qp.RPYReference, out["qp_input"][4:7]) < 1e-12
            assert rel(o.qp.momentumReference, out["qp_input"][7:13]) < 1e-12
        r.close()


def test_live_reference_with_other_horizons_and_weights():
    """Configuration is read by the reference's own readConfigParameters: a longer horizon, another grid, other weights and
    throttle limits, no jet dynamics, measured thrust instead of the estimate."""
    rd = _reference()
    syn = pkg("synthetic")
    nom = syn.make_states(1, seed=6, perturbed=False)
    cases = [dict(nIter=34, nIterSmall=14, controlHorizon=24), dict(nIter=12, nIterSmall=4, controlHorizon=12),
             dict(periodMPCLargeSteps=0.2, weightThrottle=1000.0, weightCoMPos=[50.0, 60.0, 70.0], throttleMin=10.0, throttleMax=90.0),
             dict(useJetDynamic=False), dict(useEstimatedThrust=False)]
    for c in cases:
        o, r = OracleInstance(nom, 0, params=c), rd.ReferenceInstance(nom, 0, params=c)
        assert (r.n_var, r.n_con) == (o.mpc.nVar, o.mpc.nConstraints), c
        for t in range(22):
            st = syn.make_states(1, seed=300 + t, perturbed=True, near_bound_fraction=0.5)
            o.update(st); r.update(st)
            m = o.mpc
            for name, x, y in zip("PqAlu", r.qp(), (m.hessian, m.gradient, m.linearMatrix, m.lowerBound, m.upperBound)):
                assert rel(y, x) < 1e-12, (c, name, t)
            assert rel(o.solve(), r.solve()) < 1e-9, (c, t)
        r.close()


def test_reference_golden_files_are_what_the_compiled_reference_returns():
    rd = _reference()
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location(
        "make_reference_golden", os.path.join(os.path.dirname(__file__), "golden", "make_reference_golden.py"))
    m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
    a, g = m.single_tick(), golden("reference_qp.npz")
    for k in g.files:
        assert np.array_equal(a[k], g[k]) or rel(a[k], g[k]) < 1e-13, k
    b, g = m.tick_sequence(), golden("reference_ticks.npz")
    for k in g.files:
        assert np.array_equal(b[k], g[k]) or rel(b[k], g[k]) < 1e-13, k
