"""NumPy specification of the FALLBACK QP kernel (csrc/vsmpc_qp_fallback.cu): the same QP solved by
forward condensing + an equilibrated Cholesky factorisation in FORWARD variable order.

Why a second algorithm: the Riccati recursion of the default kernels is a block elimination in BACKWARD order.  When the
open-loop transition T_k = I + dt_k A_c has a spectral radius well above one — a vehicle spinning at |omega_B| >~ 30 rad/s
makes the explicit-Euler momentum block I - dt S(omega_B) expand by |1 + i dt omega| ~ 10 per coarse knot — the cost-to-go P
grows by that factor squared per knot (1e34 over the horizon), the condensed Hessian is GRADED over 30 decades, and
eliminating its small-scale variables (the late knots) first destroys the large-scale ones by cancellation: non-positive
pivots, status 2.  The QP itself stays well posed (the oracle's pivoted sparse-KKT solve and the reference's OSQP both
return its minimiser; profiles/r02_nonsolved_adjudication.md).  For a graded SPD matrix the stable elimination order is
largest scale first = the EARLY knots first = forward order, after scaling to a unit diagonal.

    x_k = a_k + S_k w,  w = (dq_0 .. dq_{Nc-1}, v_0 .. v_{nblk-1}):   S_{k+1} = T_k S_k + E_k,  a_{k+1} = T_k a_k + dt_k c
    H = R_w + sum_k S_k' Q S_k,   g = g_w + sum_k S_k' Q (a_k - xref_k)
    unit-diagonal scaling, Cholesky of the dq block (natural = forward order), Schur complement onto the throttle
    variables, the SAME principal-pivot dual active set as the default kernels on the <= 4 nblk boxes, back-substitution,
    forward rollout.
"""
import numpy as np

from condensed_model import box_qp_pivot

NX, NJ, NT = 26, 8, 4


def solve_forward_condensed(Ac, BJ, BT, c, dt, Qd, xref, Rqd, gq, w_t, w_i, vbar, pinned, vmin, vmax, x0, N, Ns, Nc,
                            max_iter=400, tol=1e-10):
    """Returns (x (N+1, 26), dq (Nc, 8), v (nblk, 4), status, info)."""
    nblk = Nc - Ns + 1
    nq, nv = NJ * Nc, NT * nblk
    nw = nq + nv
    jb = [min(k, Nc - 1) for k in range(N)]
    tb = [0 if k < Ns else (k - (Ns - 1) if k < Nc else Nc - Ns) for k in range(N)]
    H = np.zeros((nw, nw))
    g = np.zeros(nw)
    for j in range(Nc):
        H[NJ * j:NJ * j + NJ, NJ * j:NJ * j + NJ] = np.diag(Rqd)
        g[NJ * j:NJ * j + NJ] = gq
    L = np.zeros((nblk, nblk))
    for b in range(nblk - 1):
        L[b, b] += 1; L[b + 1, b + 1] += 1; L[b, b + 1] -= 1; L[b + 1, b] -= 1
    H[nq:, nq:] = w_t * np.kron(L, np.eye(NT))
    H[nq:nq + NT, nq:nq + NT] += w_i * np.eye(NT)
    g[nq:nq + NT] -= w_i * vbar
    S = np.zeros((NX, nw))
    a = np.array(x0, float)
    I = np.eye(NX)
    for k in range(N):
        T = I + dt[k] * Ac
        S = T @ S
        S[:, NJ * jb[k]:NJ * jb[k] + NJ] += dt[k] * BJ
        S[:, nq + NT * tb[k]:nq + NT * tb[k] + NT] += dt[k] * BT
        a = T @ a + dt[k] * c
        QS = Qd[:, None] * S
        H += S.T @ QS
        g += QS.T @ (a - xref[k])
    first = NT if pinned else 0
    if pinned:      # block 0 is a parameter
        g = g + H[:, nq:nq + NT] @ vbar
    free_v = np.arange(nq + first, nw)
    keep = np.concatenate([np.arange(nq), free_v])
    Hk, gk = H[np.ix_(keep, keep)], g[keep]
    d = 1.0 / np.sqrt(np.diag(Hk))                       # unit-diagonal scaling
    Hs, gs = Hk * d[:, None] * d[None, :], gk * d
    Lq = np.linalg.cholesky(Hs[:nq, :nq])                # forward order: dq_0 first
    Y = np.linalg.solve(Lq, Hs[:nq, nq:])                # Lq^-1 H_qv
    yq = np.linalg.solve(Lq, gs[:nq])
    Hr = Hs[nq:, nq:] - Y.T @ Y
    gr = gs[nq:] - Y.T @ yq
    dv = d[nq:]
    vs, active, status = box_qp_pivot(Hr, gr, vmin / dv, vmax / dv, max_iter=max_iter, tol=tol)
    ws_q = -np.linalg.solve(Lq.T, yq + Y @ vs)
    w = np.zeros(nw)
    w[:nq] = ws_q * d[:nq]
    w[free_v] = vs * dv
    for i, s, _ in active:
        w[free_v[i]] = vmax if s > 0 else vmin           # exactly on the bound
    if pinned:
        w[nq:nq + NT] = vbar
    dq, v = w[:nq].reshape(Nc, NJ), w[nq:].reshape(nblk, NT)
    x = np.zeros((N + 1, NX))
    x[0] = x0
    for k in range(N):
        x[k + 1] = (I + dt[k] * Ac) @ x[k] + dt[k] * (BJ @ dq[jb[k]] + BT @ v[tb[k]] + c)
    return x, dq, v, status, dict(n_active=len(active), cond_scaled_q=float(np.linalg.cond(Hs[:nq, :nq])),
                                  cond_reduced=float(np.linalg.cond(Hr)), diag_range=float(np.diag(Hk).max() / np.diag(Hk).min()))
