"""Test helpers: bridge the product's batched synthetic states to the per-instance oracle."""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PKG_NAME = "paper_gorbani_2025_humanoids_multi-rate-mpc-ironcub_b200"


def pkg(sub: str = ""):
    return importlib.import_module(PKG_NAME + (("." + sub) if sub else ""))


def load_trajectories():
    """Trajectory fixtures converted from the reference's .mat files (tests/golden/make_fixtures.py)."""
    d = np.load(os.path.join(ROOT, "tests", "golden", "trajectories.npz"))
    return {
        "TRAJECTORY_MANAGER": dict(fps=int(d["alpha_fps"]), arrays={"alphaGravity": d["alphaGravity"]}),
        "POSITION_TRAJECTORY": dict(fps=int(d["traj_fps"]), arrays={
            "positionCoM": d["positionCoM"], "velocityCoM": d["velocityCoM"],
            "RPY": d["RPY"], "RPYDot": d["RPYDot"]}),
    }


def robot_data(state: dict, i: int):
    """Instance ``i`` of a getter-level batch -> oracle RobotData."""
    from oracle.vsmpc_oracle import RobotData
    return RobotData(
        wRb=state["wRb"][i].copy(), base_pos=state["base_pos"][i].copy(),
        omega_world=state["omega_world"][i].copy(), rpy=state["rpy"][i].copy(),
        mass_matrix_base=state["M_b"][i].copy(), p_com=state["p_com"][i].copy(),
        momentum_body=state["momentum_body"][i].copy(), A_mom_body=state["A_mom_body"][i].copy(),
        jet_axes=state["jet_axes"][i].copy(), jet_arms=state["jet_arms"][i].copy(),
        J_rel_body=state["J_rel_body"][i].copy(), J_jet_lin=state["J_jet_lin"][i].copy(),
        J_com=state["J_com"][i].copy(), jet_thrusts=state["thrust"][i].copy(),
        joint_pos=state["joint_pos"][i].copy(), gravity=state["gravity"][i].copy())


def set_robot_state(robot, state: dict, i: int):
    """In-place update of a RobotData (the reference mutates one shared Robot via setState)."""
    new = robot_data(state, i)
    for k, v in new.__dict__.items():
        setattr(robot, k, v)


def fill_qp_input(qp, state: dict, i: int):
    qp.setThrottleMPC(state["throttle_prev"][i])
    qp.setThrustDesMPC(state["thrust_des"][i])
    qp.setThrustDotDesMPC(state["thrust_dot_des"][i])
    qp.setEstimatedThrustDot(state["thrust_dot_est"][i])
    qp.setOutputQPJointsPosition(state["q_cmd"][i])


def state_from_pack(pack, joint_pos_sel=None, nJ=23, sel=tuple(range(3, 11))):
    """Rebuild a getter-level batch dict from a stored SoA pack (golden inputs).  Jacobian columns of
    the non-controlled joints are zero; they never enter the MPC (only the controlled columns are read)."""
    P = pkg("pack")
    B = pack.shape[1]
    f = lambda name, shape: np.ascontiguousarray(pack[P.PACK_OFFSETS[name][0]:sum(P.PACK_OFFSETS[name])].T).reshape((B,) + shape)
    J_rel_body = np.zeros((B, 4, 6, nJ))
    J_rel_body[:, :, 3:6, :][..., list(sel)] = f("J_rel_ang", (4, 3, 8))
    J_jet_lin = np.zeros((B, 4, 3, nJ))
    J_jet_lin[..., list(sel)] = f("J_jet_lin", (4, 3, 8))
    J_com = np.zeros((B, 3, nJ))
    J_com[..., list(sel)] = f("J_com", (3, 8))
    q_cmd = np.zeros((B, nJ))
    q_cmd[:, list(sel)] = f("q_cmd", (8,))
    joint_pos = np.zeros((B, nJ))
    if joint_pos_sel is not None:
        joint_pos[:, list(sel)] = np.asarray(joint_pos_sel).T
    return dict(
        wRb=f("wRb", (3, 3)), omega_world=f("omega_world", (3,)), rpy=f("rpy", (3,)), mass=f("mass", (1,))[:, 0],
        gravity=f("gravity", (3,)), M_b=f("M_b", (6, 6)), base_pos=f("base_pos", (3,)), p_com=f("p_com", (3,)),
        momentum_body=f("momentum_body", (6,)), A_mom_body=f("A_mom_body", (6, 4)), jet_axes=f("jet_axes", (4, 3)),
        jet_arms=f("jet_arms", (4, 3)), J_rel_body=J_rel_body, J_jet_lin=J_jet_lin, J_com=J_com,
        thrust=f("thrust", (4,)), thrust_dot_est=f("thrust_dot_est", (4,)), thrust_des=f("thrust_des", (4,)),
        thrust_dot_des=f("thrust_dot_des", (4,)), throttle_prev=f("throttle_prev", (4,)), q_cmd=q_cmd,
        joint_pos=joint_pos)


def golden(name):
    return np.load(os.path.join(ROOT, "tests", "golden", name))


# ---- per-field relative checks (north_star: "optimal thrusts and joint commands must agree within 1e-6 relative") ----------
# Every physical quantity is compared against ITS OWN magnitude: a mixed-unit vector divided by its largest entry
# (thrusts ~ 1e2 N) would let a 1e-4 error on a 1e-3 rad joint command through.  `floor` keeps the quotient defined for a
# field that is identically zero in the reference (e.g. error integrals at the first knot).
OUT_FIELDS = (("delta_q", 0, 8), ("throttle", 8, 12), ("thrust", 12, 16), ("thrust_dot", 16, 20),
              ("final_com", 20, 23), ("final_lin_mom", 23, 26), ("final_rpy", 26, 29), ("final_ang_mom", 29, 32),
              ("final_thrust", 32, 36), ("final_thrust_dot", 36, 40), ("final_pos_err", 40, 43), ("final_rpy_err", 43, 46),
              ("joints_ref", 46, 54))
STATE_FIELDS = (("com", 0, 3), ("lin_mom", 3, 6), ("rpy", 6, 9), ("ang_mom", 9, 12), ("thrust", 12, 16),
                ("thrust_dot", 16, 20), ("pos_err", 20, 23), ("rpy_err", 23, 26))


def field_rel(a, b, floor=1e-9):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), floor))


def output_row_errors(row, ref, floor=1e-9):
    """{field: relative error of that field of a 54-double output row}."""
    return {n: field_rel(row[lo:hi], ref[lo:hi], floor) for n, lo, hi in OUT_FIELDS}


def assert_output_rows_close(rows, refs, tol, what=""):
    rows, refs = np.atleast_2d(rows), np.atleast_2d(refs)
    for i in range(rows.shape[0]):
        errs = output_row_errors(rows[i], refs[i])
        bad = {k: v for k, v in errs.items() if not v < tol}
        assert not bad, (what, i, bad)


def solution_errors(z, zo, N, Nc, nblk, floor=1e-9):
    """Per-quantity relative errors of a full primal [x_0..x_N | dq_0..dq_{Nc-1} | v_0..v_{nblk-1}]: every state field over
    all knots, the joint increments, the throttle variables."""
    z, zo = np.asarray(z, float), np.asarray(zo, float)
    nx = 26 * (N + 1)
    x, xo = z[:nx].reshape(N + 1, 26), zo[:nx].reshape(N + 1, 26)
    e = {"x_" + n: field_rel(x[:, lo:hi], xo[:, lo:hi], floor) for n, lo, hi in STATE_FIELDS}
    e["delta_q"] = field_rel(z[nx:nx + 8 * Nc], zo[nx:nx + 8 * Nc], floor)
    e["v"] = field_rel(z[nx + 8 * Nc:nx + 8 * Nc + 4 * nblk], zo[nx + 8 * Nc:nx + 8 * Nc + 4 * nblk], floor)
    return e


def assert_solution_close(z, zo, tol, N=17, Nc=12, nblk=6, what=""):
    errs = solution_errors(z, zo, N, Nc, nblk)
    bad = {k: v for k, v in errs.items() if not v < tol}
    assert not bad, (what, bad)


def split_hessian(H, n_iter, n_ctrl):
    """IMPCProblem::getHessian = diagonal + the throttle path-graph Laplacian (costsVSMPC.cpp:383-409): returns
    (diag(H), w_t) after checking that nothing else is off the diagonal."""
    Pd = np.diag(H).copy()
    off = H - np.diag(Pd)
    nxj = 26 * (n_iter + 1) + 8 * n_ctrl
    assert np.count_nonzero(off[:nxj]) == 0 and np.count_nonzero(off[:, :nxj]) == 0
    w = np.unique(off[off != 0])
    assert w.size <= 1
    return Pd, (-float(w[0]) if w.size else 0.0)


def kkt_certificate(z, A, BJ, BT, dt, q, l, u, Pd, w_t, n_iter=17, n_small=7, n_ctrl=12, dq_lo=None, dq_hi=None):
    """Solver-independent KKT certificate of a batch of returned primals, in NumPy, from what the library itself exposes
    (dense continuous-time blocks, dt grid, gradient, bounds, Hessian).  Lagrangian g + A'y = 0 with g = P z + q:
      * the multipliers of the dynamics rows follow from stationarity in x_N .. x_1 by a backward costate recursion
        (y_{N-1} = g_{x_N}, y_{k-1} = g_{x_k} + T_k' y_k), the x_0 rows absorb whatever is left at knot 0;
      * stationarity in every joint-increment block is then a genuine residual: g_dq_j + sum_{k: block j} dt_k B_J' y_k;
      * in every throttle block the same sum IS minus the multiplier of its box row: it must vanish strictly inside the box,
        be <= 0 on the lower bound and >= 0 on the upper bound (l <= v <= u; pinned rows are equalities: free sign).
    With the optional joint-limit rows (dq_lo / dq_hi: (B, 8) bounds of every increment block) the joint-increment residual
    is read the same way: zero strictly inside the box, <= 0 on the upper bound and >= 0 on the lower one (keys *_dq, box_dq,
    n_dq_at_bound); stationarity_dq then covers the increments strictly inside.
    Returns relative residuals (each against the magnitude of the terms it is the difference of)."""
    B = z.shape[0]
    nblk = n_ctrl - n_small + 1
    nx = 26 * (n_iter + 1)
    g = z * Pd[None, :] + q          # P = diag(Pd) - w_t * (off-diagonals of the throttle path-graph Laplacian), added below
    v = z[:, nx + 8 * n_ctrl:].reshape(B, nblk, 4)
    gv = g[:, nx + 8 * n_ctrl:].reshape(B, nblk, 4).copy()
    gv[:, 1:] -= w_t * v[:, :-1]
    gv[:, :-1] -= w_t * v[:, 1:]
    gx = g[:, :nx].reshape(B, n_iter + 1, 26)
    gdq = g[:, nx:nx + 8 * n_ctrl].reshape(B, n_ctrl, 8)
    I = np.eye(26)[None]
    y = np.zeros((B, n_iter, 26))
    y[:, n_iter - 1] = gx[:, n_iter]
    for k in range(n_iter - 1, 0, -1):
        Tk = I + dt[k] * A
        y[:, k - 1] = gx[:, k] + np.einsum("bji,bj->bi", Tk, y[:, k])
    sdq, sdq_mag = gdq.copy(), np.abs(gdq)
    sv, sv_mag = gv.copy(), np.abs(gv)
    for k in range(n_iter):
        jb = min(k, n_ctrl - 1)
        tb = 0 if k < n_small else (k - (n_small - 1) if k < n_ctrl else n_ctrl - n_small)
        tj = dt[k] * np.einsum("bji,bj->bi", BJ, y[:, k])
        tt = dt[k] * np.einsum("bji,bj->bi", BT, y[:, k])
        sdq[:, jb] += tj
        sdq_mag[:, jb] += np.abs(tj)
        sv[:, tb] += tt
        sv_mag[:, tb] += np.abs(tt)
    extra = {}
    if dq_lo is not None:
        dq = z[:, nx:nx + 8 * n_ctrl].reshape(B, n_ctrl, 8)
        jl_, jh_ = dq_lo[:, None, :], dq_hi[:, None, :]
        j_lo, j_up = dq <= jl_ + 1e-12, dq >= jh_ - 1e-12
        j_in = ~(j_lo | j_up)
        jscale = np.maximum(sdq_mag.max(axis=(1, 2), keepdims=True), 1e-300)
        extra = dict(dual_sign_dq=((np.where(j_lo, np.maximum(-sdq, 0.0), 0.0) + np.where(j_up, np.maximum(sdq, 0.0), 0.0)) / jscale).max(axis=(1, 2)),
                     box_dq=np.maximum(np.maximum(jl_ - dq, dq - jh_), 0.0).max(axis=(1, 2)),
                     n_dq_at_bound=int((j_lo | j_up).sum()), instances_dq_at_bound=int((j_lo | j_up).any(axis=(1, 2)).sum()))
        sdq = np.where(j_in, sdq, 0.0)
    stat_dq = np.abs(sdq).max(axis=(1, 2)) / np.maximum(sdq_mag.max(axis=(1, 2)), 1e-300)
    t0 = 26 * n_iter + 26
    lo, up = l[:, t0:t0 + 4 * nblk].reshape(B, nblk, 4), u[:, t0:t0 + 4 * nblk].reshape(B, nblk, 4)
    nu = -sv                                   # multiplier of the box rows
    scale = np.maximum(sv_mag.max(axis=(1, 2), keepdims=True), 1e-300)
    eq = lo == up
    at_lo = (~eq) & (v <= lo + 1e-12)
    at_up = (~eq) & (v >= up - 1e-12)
    inside = ~(eq | at_lo | at_up)
    comp = np.where(inside, np.abs(nu), 0.0) / scale
    sign = (np.where(at_lo, np.maximum(nu, 0.0), 0.0) + np.where(at_up, np.maximum(-nu, 0.0), 0.0)) / scale
    box = np.maximum(np.maximum(lo - v, v - up), 0.0)
    return dict(stationarity_dq=stat_dq, complementarity=comp.max(axis=(1, 2)), dual_sign=sign.max(axis=(1, 2)),
                box=box.max(axis=(1, 2)), n_at_bound=int((at_lo | at_up).sum()), n_inside=int(inside.sum()), **extra)


