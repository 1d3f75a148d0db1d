"""Development model (NumPy) of K2 v3: Riccati recursion on the 26 states with the throttle blocks carried
as *parameter columns*, followed by a dense box-QP in the <= 24 throttle variables.

NOT the oracle and NOT on the product path: the executable specification of csrc/vsmpc_qp_condensed.cu
(tests/test_oracle.py compares it with the oracle's exact dense solve).

Value function at knot k, theta = (v_0 .. v_{nblk-1}, 1, dq_held):
    V_k(x; theta) = 1/2 x'P x + x'Psi theta + 1/2 theta'Om theta
  * P   (26 x 26): the ordinary Riccati matrix of the joint-increment LQR (8 x 8 eliminations) — independent of
    the throttle columns ("warp A" of the kernel);
  * Psi (26 x 32), Om (32 x 32): linear propagation / rank-8 downdates ("warp B"); column l lives on lane l:
    lanes 4b..4b+3 = throttle block b, lane 24 = affine column (references, c, gradients), and during the tail
    (knots >= Nc-1, where joint block Nc-1 is held) lanes D0..D0+7 = the held joint increments.
After knot 0:  reduced Hessian H_r = Om_vv + Laplacian + w_i E_0, gradient from Psi' x0 and the affine column;
dual active set on the boxes by exchange pivots on the principal pivot transform of H_r (box_qp_pivot below; the same
routine is the specification of the long-horizon kernel's active set); one forward pass with the stored gains
K_k (8 x 26) and F_k (8 x 32).
"""
from __future__ import annotations

import numpy as np

NX, NJ, NT4 = 26, 8, 4
NL = 32          # parameter columns = lanes of warp B
AFF = 24         # affine column
LA = [3, 4, 5, 9, 10, 11]   # rows of B_J that are non-zero (linear / angular momentum)


def block_maps(N, Ns, Nc):
    jb = [min(k, Nc - 1) for k in range(N)]
    tb = [0 if k < Ns else (k - (Ns - 1) if k < Nc else Nc - Ns) for k in range(N)]
    return jb, tb


def exchange_pivot(T, q):
    """In (outputs) = T (inputs): swap the roles of input q and output q.  Pivoting every index of H gives H^-1; pivoting
    q twice is the identity.  This is the pivot of csrc/vsmpc_qp_condensed.cu (rp_pivot) and of the long-horizon kernel
    (there as the rank-1 change T - (T[:, q] + e_q)(T[q, :] - e_q)' / T[q, q], deferred eight at a time)."""
    d = T[q, q]
    u, v = T[:, q].copy(), T[q, :].copy()
    T -= np.outer(u, v) / d
    T[q, :] = -v / d
    T[:, q] = u / d
    T[q, q] = 1.0 / d
    return d


def block_exchange_pivot(T, q1, q2):
    """Exchange pivot on two indices at once (both in the same class): T' = [[D^-1, -D^-1 V], [U D^-1, T - U D^-1 V]] with
    D = T[Q, Q].  Equal to exchange_pivot(T, q1) followed by exchange_pivot(T, q2); the inverse phase of
    csrc/vsmpc_qp_condensed.cu pivots two indices per step this way (half the publish / read / reciprocal round trips)."""
    Q = [q1, q2]
    D = T[np.ix_(Q, Q)].copy()
    Di = np.linalg.inv(D)
    U, V = T[:, Q].copy(), T[Q, :].copy()
    T -= U @ Di @ V
    T[Q, :] = -Di @ V
    T[:, Q] = U @ Di
    T[np.ix_(Q, Q)] = Di
    return np.linalg.det(D)


def _dual_pivot_loop(T, vv, act, lam, lo, up, max_iter, tol):
    """Goldfarb-Idnani dual iterations on the principal pivot transform, from any S-pair: T = transform of H over the free
    set {act == 0}, vv = minimiser of the sub-problem with the working set held at its bounds, lam >= 0 on the working set.
    Returns (status, iterations); T, vv, act, lam are updated in place."""
    status, it = 0, 0
    while True:
        viol = np.where(act == 0, np.maximum(np.maximum(vv - up, lo - vv), 0.0), 0.0)
        p = int(np.argmax(viol))
        if not viol[p] > tol:
            break
        s = 1.0 if vv[p] - up > lo - vv[p] else -1.0
        bound = up if s > 0 else lo
        lam_p = 0.0
        while True:
            it += 1
            if it > max_iter:
                status = 1
                break
            c = T[:, p].copy()
            zp = T[p, p]
            r = np.where(act != 0, -act * s * c, 0.0)
            cand = np.where((act != 0) & (r > 0), np.maximum(lam, 0.0) / np.where(r > 0, r, 1.0), np.inf)
            drop = int(np.argmin(cand))
            t1 = cand[drop]
            t2 = (s * vv[p] - s * bound) / zp if zp > 1e-300 else np.inf
            t = min(t1, t2)
            if not np.isfinite(t):
                status = 2
                break
            vv[:] = np.where(act == 0, vv - t * s * c, vv)
            lam[:] = np.where(act != 0, lam - t * r, lam)
            lam_p += t
            if t2 <= t1:
                exchange_pivot(T, p)
                act[p], lam[p], vv[p] = (1 if s > 0 else -1), lam_p, bound
                break
            exchange_pivot(T, drop)
            act[drop], lam[drop] = 0, 0.0
        if status:
            break
    return status, it


def box_qp_pivot(H, g, lo, up, max_iter=200, tol=1e-10):
    """min 1/2 v'Hv + g'v, lo <= v <= up: Goldfarb-Idnani dual active set on the principal pivot transform.
    T starts as H^-1 (every index exchanged) and stays the transform of H over the free set F: (v_F, y_W) = T (y_F, v_W)
    with y = Hv = -g - sum_a s_a lam_a e_a.  For a violated free p with sign s, raising its multiplier by t moves v_F by
    -t s T[F, p] and lam_a by -t r_a, r_a = -s_a s T[a, p]; T[p, p] is the step denominator.  A full step pivots p into
    the working set, a blocked step pivots the blocking index back out.  Returns (v, [(index, sign, multiplier)], status)."""
    n = H.shape[0]
    T = np.array(H, dtype=float)
    for q in range(n):
        if not exchange_pivot(T, q) > 0:
            return np.zeros(n), [], 2
    vv = -T @ g
    act = np.zeros(n, dtype=int)
    lam = np.zeros(n)
    status, _ = _dual_pivot_loop(T, vv, act, lam, lo, up, max_iter, tol)
    active = [(int(i), float(act[i]), float(lam[i])) for i in np.nonzero(act)[0]]
    return vv, active, status


def box_qp_pivot_warm(H, g, lo, up, act0, max_iter=200, tol=1e-10):
    """The same problem started from a guessed working set act0 (+1: at `up`, -1: at `lo`, 0: free; e.g. the working set of the
    previous controller tick) — the specification of the warm start of csrc/vsmpc_qp_condensed_wide.cu.
    T starts as H itself (every index in the working set: no inverse) and only the guessed-free indices are pivoted:
    |F0| pivots instead of n + |W|.
      phase 1 (drop only, finite): solve the sub-problem on the guessed set, (v_F, y_W) = T (-g_F, b_W),
        lam_a = -s_a (y_a + g_a); while some lam_a < 0, pivot the most negative one out of the working set and re-solve.
        It ends on an S-pair (sub-problem optimum, lam >= 0) whatever the guess was;
      phase 2: the dual iterations of box_qp_pivot from that S-pair (adds the bounds still violated).
    Returns (v, [(index, sign, multiplier)], status, pivots) with pivots = exchange pivots executed in total."""
    n = H.shape[0]
    lo_v, up_v = np.broadcast_to(np.asarray(lo, float), (n,)), np.broadcast_to(np.asarray(up, float), (n,))
    act = np.array(act0, dtype=int).copy()
    T = np.array(H, dtype=float)
    pivots = 0
    for q in np.flatnonzero(act == 0):
        pivots += 1
        if not exchange_pivot(T, q) > 0:
            return np.zeros(n), [], 2, pivots
    b = np.where(act > 0, up_v, lo_v)
    out = T @ np.where(act == 0, -g, b)         # the one matrix-vector product: v on the free set, y = Hv on the working set
    while True:
        lam = np.where(act != 0, -act * (out + g), 0.0)
        a = int(np.argmin(lam))
        if not lam[a] < -tol:
            break
        pivots += 1
        if not exchange_pivot(T, a) > 0:
            return np.zeros(n), [], 2, pivots
        # a leaves the working set.  The pivot only relabels input a (was v_a = b_a) and output a (was y_a = out_a): the
        # same point satisfies the new relation; then the new input y_a moves from out_a to its free-variable value -g_a and
        # every output follows column a of the new transform (what the kernel does: no second matrix-vector product)
        delta = -g[a] - out[a]
        col = T[:, a].copy()
        out = out + col * delta
        out[a] = b[a] + col[a] * delta
        act[a] = 0
    vv = np.where(act == 0, out, b)
    lam = np.maximum(lam, 0.0)
    status, it = _dual_pivot_loop(T, vv, act, lam, lo, up, max_iter, tol)
    active = [(int(i), float(act[i]), float(lam[i])) for i in np.nonzero(act)[0]]
    return vv, active, status, pivots + it


class CondensedQP:
    def __init__(self, Ac, BJ, BT, c, dt, Qd, xref, Rqd, gq, w_t, w_i, vbar, pinned, vmin, vmax, x0,
                 N, Ns, Nc):
        self.__dict__.update(locals())
        self.jb, self.tb = block_maps(N, Ns, Nc)
        self.nblk = Nc - Ns + 1
        assert 4 * self.nblk <= AFF
        self.D0 = 0 if self.nblk >= 3 else 16     # lanes of the held joint block during the tail
        self.held = Nc - 1 < N - 1                # joint block Nc-1 spans more than one knot

    def _D(self, k):
        """x+ = T x + B_u u + D theta: columns of D at knot k."""
        D = np.zeros((NX, NL))
        b = self.tb[k]
        D[:, 4 * b:4 * b + 4] = self.dt[k] * self.BT
        D[:, AFF] = self.dt[k] * self.c
        if self.held and k >= self.Nc - 1:
            D[:, self.D0:self.D0 + NJ] = self.dt[k] * self.BJ
        return D

    def factor(self):
        N, Nc = self.N, self.Nc
        P = np.zeros((NX, NX))
        Psi = np.zeros((NX, NL))
        Om = np.zeros((NL, NL))
        self.K = [None] * N
        self.F = [None] * N
        I = np.eye(NX)
        for k in range(N - 1, -1, -1):
            T = I + self.dt[k] * self.Ac
            Bu = self.dt[k] * self.BJ
            D = self._D(k)
            Pp = P + np.diag(self.Qd)                       # P'
            Psp = Psi.copy()                                # Psi' incl. the tracking gradient of x_{k+1}
            Psp[:, AFF] -= self.Qd * self.xref[k]
            Ps2 = Psp + Pp @ D                              # Psi''
            PT = T.T @ Pp @ T
            PsT = T.T @ Ps2
            OmT = Om + D.T @ Ps2 + Psp.T @ D
            tail = self.held and k >= Nc - 1
            if not tail:
                # fresh joint block: eliminate inside the stage
                Huu = np.diag(self.Rqd) + Bu.T @ Pp @ Bu
                Hux = Bu.T @ Pp @ T
                Hut = Bu.T @ Ps2
                Hut[:, AFF] += self.gq
            elif k == Nc - 1:
                # held block: it was a parameter through the tail; Schur complement on its columns
                d = slice(self.D0, self.D0 + NJ)
                Huu = OmT[d, d] + np.diag(self.Rqd)
                Hux = PsT[:, d].T.copy()
                Hut = OmT[d, :].copy()
                Hut[:, AFF] += self.gq
                Hut[:, d] = 0.0
                PsT[:, d] = 0.0
                OmT[d, :] = 0.0
                OmT[:, d] = 0.0
            else:
                P, Psi, Om = PT, PsT, OmT
                continue
            Hi = np.linalg.inv(Huu)
            Kk = Hi @ Hux
            Fk = Hi @ Hut
            P = PT - Hux.T @ Kk
            Psi = PsT - Hux.T @ Fk
            Om = OmT - Hut.T @ Fk
            self.K[k], self.F[k] = Kk, Fk
        self.P0, self.Psi0, self.Om0 = P, Psi, Om

    def reduced_qp(self):
        """min 1/2 v'H v + g'v over the free throttle variables (block 0 removed when pinned)."""
        nv = 4 * self.nblk
        H = self.Om0[:nv, :nv].copy()
        g = self.Psi0[:, :nv].T @ self.x0 + self.Om0[:nv, AFF]
        L = np.zeros((self.nblk, self.nblk))
        for b in range(self.nblk - 1):
            L[b, b] += 1; L[b + 1, b + 1] += 1; L[b, b + 1] -= 1; L[b + 1, b] -= 1
        H += self.w_t * np.kron(L, np.eye(4))
        H[:4, :4] += self.w_i * np.eye(4)
        g[:4] -= self.w_i * self.vbar
        first = 4 if self.pinned else 0
        if self.pinned:
            g = g[4:] + H[4:, :4] @ self.vbar
            H = H[4:, 4:]
        return H, g, first

    def solve_box(self, max_iter=200, tol=1e-10):
        self.factor()
        H, g, first = self.reduced_qp()
        vv, active, status = box_qp_pivot(H, g, self.vmin, self.vmax, max_iter=max_iter, tol=tol)
        self.active = active
        self.status = status
        v = np.concatenate([self.vbar, vv]) if self.pinned else vv
        return self.forward(v.reshape(self.nblk, 4))

    def forward(self, v):
        N = self.N
        theta = np.zeros(NL)
        theta[:4 * self.nblk] = v.reshape(-1)
        theta[AFF] = 1.0
        x = np.zeros((N + 1, NX))
        dq = np.zeros((self.Nc, NJ))
        x[0] = self.x0
        I = np.eye(NX)
        for k in range(N):
            if self.K[k] is not None:
                u = -self.K[k] @ x[k] - self.F[k] @ theta
                dq[self.jb[k]] = u
            u = dq[self.jb[k]]
            x[k + 1] = (I + self.dt[k] * self.Ac) @ x[k] + self.dt[k] * (self.BJ @ u + self.BT @ v[self.tb[k]] + self.c)
        return x, dq, v

    def pack_z(self, x, dq, v):
        return np.concatenate([x.reshape(-1), dq.reshape(-1), v.reshape(-1)])


class ClampedCondensedQP(CondensedQP):
    """The optional joint-limit rows inside the condensed recursion (JL build of csrc/vsmpc_qp_condensed.cu).

    Every joint-increment block k < Nc has the box lo <= dq_k <= hi (JointPositionConstraint semantics, constraintsVSMPC.cpp:450-453).
    A primal-dual active set runs AROUND the solve: with a working set clamp[k][c] in {0, +1 (upper), -1 (lower)} a clamped
    component is the constant b_c inside the elimination of its block,
        H_uu restricted to the free components (identity on the clamped ones),
        H_utheta[m, affine] += sum_c H_uu[m][c] b_c                      for the free rows m,
        Psi[:, affine]      += sum_c b_c H_ux[c, :]',      Om[:, affine] += b_c H_utheta[c, :]'  (+ its transpose),
    and P, Psi, Om skip the clamped rows.  The forward pass sets dq_c = b_c and reads the gradient of the cost-to-go in dq_c,
        grad_c = H_ux[c, :] x_k + H_uu[c, :] dq_k + H_utheta[c, :] theta,
    off the raw rows (upper bound: multiplier -grad_c, lower bound: +grad_c).  Next working set: free components outside their
    box join at the bound they crossed, clamped ones with a negative multiplier leave; a fixed point is the minimiser (all KKT
    conditions hold).  The iteration is not guaranteed to settle in general — the kernel hands an instance to the KKT fallback
    after 14 passes — on the test workloads it takes 2-4 (solve_boxes has the exact rules)."""

    def factor_clamped(self, clamp, lo, hi):
        N, Nc = self.N, self.Nc
        P = np.zeros((NX, NX))
        Psi = np.zeros((NX, NL))
        Om = np.zeros((NL, NL))
        self.K, self.F, self.raw = [None] * N, [None] * N, [None] * N
        I = np.eye(NX)
        for k in range(N - 1, -1, -1):
            T = I + self.dt[k] * self.Ac
            Bu = self.dt[k] * self.BJ
            D = self._D(k)
            Pp = P + np.diag(self.Qd)
            Psp = Psi.copy()
            Psp[:, AFF] -= self.Qd * self.xref[k]
            Ps2 = Psp + Pp @ D
            PT = T.T @ Pp @ T
            PsT = T.T @ Ps2
            OmT = Om + D.T @ Ps2 + Psp.T @ D
            tail = self.held and k >= Nc - 1
            if not tail:
                Huu = np.diag(self.Rqd) + Bu.T @ Pp @ Bu
                Hux = Bu.T @ Pp @ T
                Hut = Bu.T @ Ps2
            elif k == Nc - 1:
                d = slice(self.D0, self.D0 + NJ)
                Huu = OmT[d, d] + np.diag(self.Rqd)
                Hux = PsT[:, d].T.copy()
                Hut = OmT[d, :].copy()
                Hut[:, d] = 0.0
                PsT[:, d] = 0.0
                OmT[d, :] = 0.0
                OmT[:, d] = 0.0
            else:
                P, Psi, Om = PT, PsT, OmT
                continue
            Hut[:, AFF] += self.gq
            cl = clamp[k]
            C = cl != 0
            b = np.where(cl > 0, hi, np.where(cl < 0, lo, 0.0))
            self.raw[k] = (Hux.copy(), Huu.copy(), Hut.copy())
            Hm = Huu.copy()
            Hm[C, :] = 0.0
            Hm[:, C] = 0.0
            Hm[C, C] = 1.0
            Hi = np.linalg.inv(Hm)
            hut = Hut.copy()
            hut[~C, AFF] += ((Huu - np.diag(self.Rqd)) @ b)[~C]       # off-diagonal coupling (the diagonal is never a free-clamped pair)
            Kk = Hi @ Hux                                            # clamped rows: the raw H_ux rows
            Fk = Hi @ hut                                            # clamped rows: the raw H_utheta rows
            Kf, Ff, Huxf, hutf = Kk.copy(), Fk.copy(), Hux.copy(), hut.copy()
            Kf[C], Ff[C], Huxf[C], hutf[C] = 0.0, 0.0, 0.0, 0.0
            P = PT - Huxf.T @ Kf
            Psi = PsT - Huxf.T @ Ff
            Psi[:, AFF] += Hux[C].T @ b[C]
            Om = OmT - hutf.T @ Ff
            for c in np.nonzero(C)[0]:
                Om[:, AFF] += b[c] * Hut[c]
                Om[AFF, :] += b[c] * Hut[c]
            self.K[k], self.F[k] = Kk, Fk
        self.P0, self.Psi0, self.Om0 = P, Psi, Om

    def forward_clamped(self, v, clamp, lo, hi):
        N = self.N
        theta = np.zeros(NL)
        theta[:4 * self.nblk] = v.reshape(-1)
        theta[AFF] = 1.0
        x = np.zeros((N + 1, NX))
        dq = np.zeros((self.Nc, NJ))
        grad = np.zeros((self.Nc, NJ))
        gmag = np.ones((self.Nc, NJ))
        x[0] = self.x0
        I = np.eye(NX)
        for k in range(N):
            if self.K[k] is not None:
                raw = self.K[k] @ x[k] + self.F[k] @ theta          # free rows: -dq ; clamped rows: H_ux x + H_utheta theta
                cl = clamp[k]
                u = np.where(cl > 0, hi, np.where(cl < 0, lo, -raw))
                dq[self.jb[k]] = u
                grad[k] = raw + self.raw[k][1] @ u
                gmag[k] = np.abs(raw) + np.abs(self.raw[k][1]) @ np.abs(u)
            u = dq[self.jb[k]]
            x[k + 1] = (I + self.dt[k] * self.Ac) @ x[k] + self.dt[k] * (self.BJ @ u + self.BT @ v[self.tb[k]] + self.c)
        self.gmag = gmag
        return x, dq, v, grad

    def solve_boxes(self, lo, hi, clamp0=None, max_pass=16, plain_passes=5, tol=1e-9):
        """-> x, dq, v, passes (-1: the working set did not settle), clamp.  clamp0: first guess (warm start).
        Joins: free increments more than tol outside their box.  Leaves: clamped increments whose multiplier is negative beyond
        the rounding of its own terms (1e-9 relative: a degenerate increment — on its bound, zero multiplier — must not be
        released on noise, it would come back 1e-9 outside and the iteration would never settle).  Passes 0 .. plain_passes - 1
        (from the guess) and plain_passes .. 2 plain_passes - 1 (from the EMPTY set, if the first group did not settle: an
        unrelated guess can cycle where the cold start settles) apply joins and leaves together; after that leaves wait until
        no free increment is outside its box."""
        clamp = np.zeros((self.Nc, NJ), dtype=int) if clamp0 is None else np.array(clamp0, dtype=int)
        for p in range(max_pass):
            self.factor_clamped(clamp, lo, hi)
            H, g, first = self.reduced_qp()
            vv, self.active, self.status = box_qp_pivot(H, g, self.vmin, self.vmax)
            v = (np.concatenate([self.vbar, vv]) if self.pinned else vv).reshape(self.nblk, 4)
            x, dq, v, grad = self.forward_clamped(v, clamp, lo, hi)
            add_u, add_l = (clamp == 0) & (dq > hi + tol), (clamp == 0) & (dq < lo - tol)
            rel = ((clamp > 0) & (grad > 1e-9 * self.gmag)) | ((clamp < 0) & (grad < -1e-9 * self.gmag))
            new = clamp.copy()
            new[add_u] = 1
            new[add_l] = -1
            if not (p >= 2 * plain_passes and (add_u.any() or add_l.any())):
                new[rel] = 0
            if (new == clamp).all():
                return x, dq, v, p + 1, clamp
            clamp = np.zeros_like(new) if p == plain_passes - 1 else new
        return x, dq, v, -1, clamp
