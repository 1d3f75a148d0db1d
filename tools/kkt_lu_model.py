"""NumPy specification of the FALLBACK QP kernel (csrc/vsmpc_qp_fallback.cu): the same QP solved through a
partially pivoted LU factorisation of its equality-constrained KKT system in a bordered-band ordering.

Why a second algorithm.  The Riccati recursion of the default kernels is a block elimination WITHOUT pivoting in backward
time order.  When the open-loop transition T_k = I + dt_k A_c expands — a vehicle spinning at |omega_B| >~ 30 rad/s makes the
explicit-Euler momentum block I - dt S(omega_B) grow by |1 + i dt omega| ~ 10 per coarse knot — the cost-to-go P grows by
that factor squared per knot (1e30 over the horizon) and the rank-8 down-dates cancel catastrophically: non-positive pivots,
status 2 (0.2 % of the Monte Carlo loops of configs[2], all of them lost vehicles; profiles/r02_nonsolved_adjudication.md).
Forward condensing fails the same way (the early controls all excite the same expanding directions).  The QP itself stays
well posed: the oracle's pivoted sparse-KKT solve and the reference's OSQP return its minimiser, and so does a banded LU with
ROW pivoting, which picks per unknown whether a dynamics row is solved forward (for x_{k+1}) or backward (for x_k).

Ordering (all sizes from (N, Ns, Nc); `kkt_layout`):
    [ mu (x0 rows) | stage 0 | stage 1 | ... | stage N | border ],   stage k = [ x_k | dq_k | v_b(k) | nu_k ]
    * dq_k is in its stage when joint block k acts on knot k only; v_b is in the stage of the knot it acts on when it acts
      on one knot only; nu_k = multipliers of the dynamics rows of knot k (absent in stage N);
    * border = the input blocks that act over several knots (first / last throttle block, held joint block) + the four pin
      rows of throttle block 0 (an identity row/column when the tick is released): <= 20 unknowns.
    Half bandwidth of the band part: 90 at the reference horizon.
Elimination: column by column, pivot row searched within the band window only (border rows are eliminated, never chosen,
until the band part is done), right-hand sides carried as extra columns: (-q, b) and one unit vector per throttle variable.
The vv block of K^-1 is the inverse of the reduced throttle Hessian, i.e. the fully exchanged principal pivot transform the
dual active set of the default kernels starts from: the same Goldfarb-Idnani iterations run on it, and the minimiser is
z = z_unc - sum_a s_a lam_a K^-1 e_a.
"""
import numpy as np

from condensed_model import _dual_pivot_loop

NX, NJ, NT = 26, 8, 4


def kkt_layout(N, Ns, Nc):
    """Positions of the unknowns in the bordered-band ordering.
    Returns dict(n, nb, bw, pos_x[(N+1)*26], pos_dq[Nc*8], pos_v[nblk*4], pos_nu[N*26], pos_mu[26], pos_pin[4])."""
    nblk = Nc - Ns + 1
    jb = [min(k, Nc - 1) for k in range(N)]
    tb = [0 if k < Ns else (k - (Ns - 1) if k < Nc else Nc - Ns) for k in range(N)]
    span_j = [sum(1 for k in range(N) if jb[k] == j) for j in range(Nc)]
    span_t = [sum(1 for k in range(N) if tb[k] == b) for b in range(nblk)]
    pos_x = np.zeros((N + 1) * NX, int); pos_dq = -np.ones(Nc * NJ, int); pos_v = -np.ones(nblk * NT, int)
    pos_nu = np.zeros(N * NX, int); pos_mu = np.zeros(NX, int); pos_pin = np.zeros(NT, int)
    o = 0
    pos_mu[:] = np.arange(o, o + NX); o += NX
    for k in range(N + 1):
        pos_x[k * NX:(k + 1) * NX] = np.arange(o, o + NX); o += NX
        if k < N:
            j, b = jb[k], tb[k]
            if span_j[j] == 1:
                pos_dq[j * NJ:(j + 1) * NJ] = np.arange(o, o + NJ); o += NJ
            if span_t[b] == 1 and b != 0:      # block 0 always sits in the border next to its pin rows
                pos_v[b * NT:(b + 1) * NT] = np.arange(o, o + NT); o += NT
            pos_nu[k * NX:(k + 1) * NX] = np.arange(o, o + NX); o += NX
    nb = o
    for j in range(Nc):
        if pos_dq[j * NJ] < 0:
            pos_dq[j * NJ:(j + 1) * NJ] = np.arange(o, o + NJ); o += NJ
    for b in range(nblk):
        if pos_v[b * NT] < 0:
            pos_v[b * NT:(b + 1) * NT] = np.arange(o, o + NT); o += NT
    pos_pin[:] = np.arange(o, o + NT); o += NT
    return dict(n=o, nb=nb, pos_x=pos_x, pos_dq=pos_dq, pos_v=pos_v, pos_nu=pos_nu, pos_mu=pos_mu, pos_pin=pos_pin,
                jb=jb, tb=tb, nblk=nblk)


def assemble_kkt(L, Ac, BJ, BT, c, dt, Qd, xref, Rqd, gq, w_t, w_i, vbar, pinned, x0, N, Ns, Nc):
    """Dense K (n x n) and the right-hand side of the equality-constrained problem in the layout L."""
    n, nblk = L["n"], L["nblk"]
    K = np.zeros((n, n)); r = np.zeros(n)
    px, pq, pv, pnu, pmu, ppin = L["pos_x"], L["pos_dq"], L["pos_v"], L["pos_nu"], L["pos_mu"], L["pos_pin"]
    for k in range(1, N + 1):                                   # tracking cost on x_1 .. x_N
        for i in range(NX):
            K[px[k * NX + i], px[k * NX + i]] = Qd[i]
            r[px[k * NX + i]] = Qd[i] * xref[k - 1][i]          # -q
    for j in range(Nc):
        for a in range(NJ):
            K[pq[j * NJ + a], pq[j * NJ + a]] = Rqd[a]
            r[pq[j * NJ + a]] = -gq[a]
    for b in range(nblk):
        for a in range(NT):
            p = pv[b * NT + a]
            K[p, p] = w_t * ((1.0 if b > 0 else 0.0) + (1.0 if b < nblk - 1 else 0.0)) + (w_i if b == 0 else 0.0)
            if b > 0:
                K[p, pv[(b - 1) * NT + a]] = -w_t
            if b < nblk - 1:
                K[p, pv[(b + 1) * NT + a]] = -w_t
    for a in range(NT):
        r[pv[a]] = w_i * vbar[a]
    I = np.eye(NX)
    for k in range(N):                                          # dynamics rows: T x_k - x_{k+1} + dt B_J dq + dt B_T v = -dt c
        T = I + dt[k] * Ac
        for i in range(NX):
            row = pnu[k * NX + i]
            for j in range(NX):
                if T[i, j] != 0.0:
                    K[row, px[k * NX + j]] = K[px[k * NX + j], row] = T[i, j]
            K[row, px[(k + 1) * NX + i]] = K[px[(k + 1) * NX + i], row] = -1.0
            for a in range(NJ):
                if BJ[i, a] != 0.0:
                    q_ = pq[L["jb"][k] * NJ + a]
                    K[row, q_] = K[q_, row] = dt[k] * BJ[i, a]
            for a in range(NT):
                if BT[i, a] != 0.0:
                    v_ = pv[L["tb"][k] * NT + a]
                    K[row, v_] = K[v_, row] = dt[k] * BT[i, a]
            r[row] = -dt[k] * c[i]
    for i in range(NX):                                         # x_0 = x0
        K[pmu[i], px[i]] = K[px[i], pmu[i]] = 1.0
        r[pmu[i]] = x0[i]
    for a in range(NT):                                         # pin rows v_0 = vbar (identity when released)
        if pinned:
            K[ppin[a], pv[a]] = K[pv[a], ppin[a]] = 1.0
            r[ppin[a]] = vbar[a]
        else:
            K[ppin[a], ppin[a]] = 1.0
    return K, r


def bandwidth(K, nb):
    i, j = np.nonzero(K[:nb, :nb])
    return int(np.abs(i - j).max())


def lu_solve_bordered_band(K, R, nb, bw):
    """Gaussian elimination of [K | R] with row pivoting restricted to the band window for the first nb columns (rows
    k .. k + bw, never a border row), unrestricted in the border block; back-substitution.  Returns (X, ok)."""
    n = K.shape[0]
    M = np.concatenate([K, R], axis=1).astype(float)
    ok = True
    for k in range(n):
        hi = min(k + bw, nb - 1) if k < nb else n - 1
        p = k + int(np.argmax(np.abs(M[k:hi + 1, k])))
        if not np.abs(M[p, k]) > 0.0 or not np.isfinite(M[p, k]):
            return None, False
        if p != k:
            M[[k, p]] = M[[p, k]]
        rows = np.concatenate([np.arange(k + 1, hi + 1), np.arange(max(nb, k + 1), n)]) if k < nb else np.arange(k + 1, n)
        rows = rows[M[rows, k] != 0.0]
        if rows.size:
            cmax = min(k + 2 * bw, nb - 1) if k < nb else n - 1
            cols = np.concatenate([np.arange(k + 1, cmax + 1), np.arange(max(nb, k + 1), M.shape[1])]) if k < nb \
                else np.arange(k + 1, M.shape[1])
            l = M[rows, k] / M[k, k]
            M[np.ix_(rows, cols)] -= np.outer(l, M[k, cols])
            M[rows, k] = 0.0
    X = np.zeros((n, R.shape[1]))
    for k in range(n - 1, -1, -1):
        X[k] = (M[k, n:] - M[k, k + 1:n] @ X[k + 1:]) / M[k, k]
    return X, ok


def solve_kkt_lu(Ac, BJ, BT, c, dt, Qd, xref, Rqd, gq, w_t, w_i, vbar, pinned, vmin, vmax, x0, N, Ns, Nc,
                 max_iter=400, tol=1e-10):
    """Returns (x (N+1, 26), dq (Nc, 8), v (nblk, 4), status, info)."""
    L = kkt_layout(N, Ns, Nc)
    n, nblk = L["n"], L["nblk"]
    K, r = assemble_kkt(L, Ac, BJ, BT, c, dt, Qd, xref, Rqd, gq, w_t, w_i, vbar, pinned, x0, N, Ns, Nc)
    bw = bandwidth(K, L["nb"])
    first = NT if pinned else 0
    vidx = L["pos_v"][first:]
    nvf = vidx.size
    R = np.zeros((n, 1 + nvf))
    R[:, 0] = r
    R[vidx, 1 + np.arange(nvf)] = 1.0
    X, ok = lu_solve_bordered_band(K, R, L["nb"], bw)
    if not ok or not np.isfinite(X).all():
        return None, None, None, 2, dict(bw=bw, n=n)
    G = X[vidx, 1:]
    G = 0.5 * (G + G.T)
    vv = X[vidx, 0].copy()
    act, lam = np.zeros(nvf, int), np.zeros(nvf)
    status, it = _dual_pivot_loop(G, vv, act, lam, vmin, vmax, max_iter, tol) if nvf else (0, 0)
    z = X[:, 0].copy()
    for a in np.flatnonzero(act):
        z -= act[a] * lam[a] * X[:, 1 + a]
    x = z[L["pos_x"]].reshape(N + 1, NX)
    dq = z[L["pos_dq"]].reshape(Nc, NJ)
    v = z[L["pos_v"]].reshape(nblk, NT)
    for a in np.flatnonzero(act):
        v.reshape(-1)[first + a] = vmax if act[a] > 0 else vmin          # exactly on the bound
    return x, dq, v, status, dict(bw=bw, n=n, nb=L["nb"], n_active=int((act != 0).sum()), iters=it)
