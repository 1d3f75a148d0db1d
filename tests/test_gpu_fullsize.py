"""Full-size checks at BASELINE.json's bench configuration (B = 1024 instances, reference horizon) through
size-independent properties — the oracle is too slow to solve 1024 instances in seconds:
  * primal feasibility of the returned 588-vector: dynamics rows, x0 rows, throttle box / pin,
    with the dense (A, B_J, B_T, c, dt) the library itself exposes;
  * two independent algorithms (condensed kernel vs one-warp structured kernel) agree to 1e-9;
  * results do not depend on the batch composition: ragged batch sizes reproduce the big batch bit for bit."""
import numpy as np
import pytest

from helpers import load_trajectories, pkg
from oracle_driver import oracle_trajectories_to_product

pytestmark = pytest.mark.gpu
N, NS, NC_ = 17, 7, 12


def run(B, solver, states=None, seed=20251002):
    syn, bat, P = pkg("synthetic"), pkg("batched"), pkg("pack")
    nom = syn.make_states(B, perturbed=False) if states is None else states[0]
    per = syn.make_states(B, seed=seed, perturbed=True) if states is None else states[1]
    mpc = bat.BatchedVSMPC(B, None, oracle_trajectories_to_product(load_trajectories()), solver=solver, full_solution=True)
    phase0 = (np.arange(B) % 20).astype(np.int32)
    mpc.configure_pack(P.build_pack(nom), np.ascontiguousarray(nom["joint_pos"][:, P.DEFAULT_JOINT_SELECTOR].T), phase0)
    mpc.update(per)
    mpc.solveMPC()
    z = mpc.getSolution()
    out, status = mpc.get_output()
    extra = (mpc.get_dynamics(), mpc.get_qp_vectors())
    mpc.close()
    return z, out, status, extra, (nom, per)


def test_full_batch_feasible_and_solvers_agree():
    B = 1024
    z, out, status, ((A, BJ, BT, c, dt), (q, l, u)), st = run(B, 0)
    assert (status == 0).all()
    x = z[:, :26 * (N + 1)].reshape(B, N + 1, 26)
    dq = z[:, 26 * (N + 1):26 * (N + 1) + 8 * NC_].reshape(B, NC_, 8)
    v = z[:, 26 * (N + 1) + 8 * NC_:].reshape(B, NC_ - NS + 1, 4)
    scale = np.abs(x).max()
    for k in range(N):
        jb, tb = min(k, NC_ - 1), (0 if k < NS else (k - (NS - 1) if k < NC_ else NC_ - NS))
        xn = x[:, k] + dt[k] * (np.einsum("bij,bj->bi", A, x[:, k]) + np.einsum("bij,bj->bi", BJ, dq[:, jb])
                                + np.einsum("bij,bj->bi", BT, v[:, tb]) + c)
        assert np.abs(xn - x[:, k + 1]).max() < 1e-9 * scale, k
    # x0 rows and the throttle box / pin, straight from the bounds the library reports
    assert np.abs(x[:, 0] - l[:, 442:468]).max() < 1e-12 * scale
    vb = z[:, 564:588]
    assert (vb >= l[:, 468:492] - 1e-12).all() and (vb <= u[:, 468:492] + 1e-12).all()
    pinned = l[:, 468] == u[:, 468]
    assert pinned.sum() > 900 and (~pinned).sum() > 20        # staggered 20-tick phases: 95 % / 5 %
    # second, independent algorithm on the same inputs
    z2, out2, status2, _, _ = run(B, 2, states=st)
    assert (status2 == 0).all()
    assert np.abs(z - z2).max() / np.abs(z).max() < 1e-9
    assert np.abs(out - out2).max() / np.abs(out).max() < 1e-9


@pytest.mark.parametrize("solver", [0, 2])
def test_ragged_batches_reproduce_the_big_batch_bit_for_bit(solver):
    syn = pkg("synthetic")
    B = 1025
    zb, outb, statusb, _, (nom, per) = run(B, solver)
    for n, off in ((1, 0), (7, 3), (33, 100), (1023, 2)):
        # phases are assigned by position in the batch: keep them aligned by choosing offsets that are multiples of 20
        off20 = (off // 20) * 20
        sub = lambda d, o=off20: {k: val[o:o + n] for k, val in d.items()}
        z, out, status, _, _ = run(n, solver, states=(sub(nom), sub(per)))
        assert (status == 0).all()
        assert np.array_equal(z, zb[off20:off20 + n]) and np.array_equal(out, outb[off20:off20 + n])


def test_kkt_certificate_and_sampled_oracle_at_the_bench_configuration():
    """BASELINE configs[1] exactly as bench.py runs it (B = 1024, staggered 20-tick phases): a KKT certificate for ALL
    instances + the oracle's exact solve on a seeded sample of 64 of them, per physical quantity at 1e-6."""
    from helpers import assert_output_rows_close, assert_solution_close, kkt_certificate, split_hessian
    from oracle_driver import OracleInstance
    syn, bat, P = pkg("synthetic"), pkg("batched"), pkg("pack")
    B = 1024
    traj = load_trajectories()
    nom = syn.make_states(B, perturbed=False)
    per = syn.make_states(B, seed=20251002, perturbed=True)            # bench.py: make_workload(B, 20251002 + rank, ...)
    mpc = bat.BatchedVSMPC(B, None, oracle_trajectories_to_product(traj), solver=0, full_solution=True)
    phase0 = (np.arange(B) % 20).astype(np.int32)
    mpc.configure_pack(P.build_pack(nom), np.ascontiguousarray(nom["joint_pos"][:, P.DEFAULT_JOINT_SELECTOR].T), phase0)
    mpc.update(per)
    A, BJ, BT, c, dt = mpc.get_dynamics()
    q, l, u = mpc.get_qp_vectors()
    H = mpc.getHessian(0)
    mpc.solveMPC()
    z = mpc.getSolution()
    out, status = mpc.get_output()
    piv = mpc.get_pivot_counts()
    mpc.close()
    assert (status == 0).all()
    Pd, w_t = split_hessian(H, N, NC_)
    k = kkt_certificate(z, A, BJ, BT, dt, q, l, u, Pd, w_t, N, NS, NC_)
    assert k["stationarity_dq"].max() < 1e-9, k["stationarity_dq"].max()
    assert k["complementarity"].max() < 1e-9, k["complementarity"].max()
    assert k["dual_sign"].max() < 1e-9, k["dual_sign"].max()
    assert k["box"].max() < 1e-12
    assert k["n_at_bound"] > 50 and k["n_inside"] > 1000          # the workload exercises the active set
    assert piv.min() >= 20 and piv.max() > 24                      # pinned ticks invert 20 variables; active bounds add pivots
    # the oracle on a seeded sample of the same batch (its phase counters started ahead like the product's)
    sample = np.random.default_rng(64).choice(B, 64, replace=False)
    sample = np.unique(np.concatenate([sample, np.flatnonzero(l[:, 468] != u[:, 468])[:4]]))   # + released ticks
    for i in sample:
        o = OracleInstance(nom, int(i), trajectories=traj, phase0=int(phase0[i]))
        o.update(per)
        zo = o.solve()
        assert np.abs(q[i] - o.mpc.gradient).max() <= 1e-12 * np.abs(o.mpc.gradient).max()
        assert np.abs(l[i] - o.mpc.lowerBound).max() <= 1e-12 * max(1.0, np.abs(o.mpc.lowerBound).max())
        assert_solution_close(z[i], zo, 1e-6, what=("z", int(i)))
        assert_output_rows_close(out[i], o.output_row(), 1e-6, what=("row", int(i)))
