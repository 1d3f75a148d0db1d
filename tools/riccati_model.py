"""Development model (NumPy) of the K2 structured-QP algorithm: stage-wise Riccati recursion with
move-blocked inputs + Goldfarb-Idnani dual active set on the throttle boxes.

This is NOT the oracle and NOT on the product path; it is the executable specification the CUDA
kernel (csrc/vsmpc_qp_structured.cu) was written from, kept so the algorithm can be studied and
unit-tested on CPU (tests/test_oracle.py compares it with the oracle's exact dense solve).

Problem:  z = [x_0..x_N | dq_0..dq_{Nc-1} | v_0..v_{NT-1}],
  x_{k+1} = (I+dt_k A) x_k + dt_k B_J dq_{jb(k)} + dt_k B_T v_{tb(k)} + dt_k c,   x_0 given,
  cost = sum_{i=1..N} 1/2 x_i'Q x_i - (Q xref_{i-1})'x_i + sum_j 1/2 dq_j'Rq dq_j + g_q'dq_j
         + w_t/2 sum_b |v_{b+1}-v_b|^2 + w_i/2 |v_0|^2 - w_i vbar'v_0,
  vmin <= v_b <= vmax, and v_0 = vbar when pinned.
"""
from __future__ import annotations

import numpy as np

NX, NJ, NT4 = 26, 8, 4
NZ = NX + NT4 + NJ  # (x, v, d)


def block_maps(N, Ns, Nc):
    jb = [min(k, Nc - 1) for k in range(N)]
    tb = [0 if k < Ns else (k - (Ns - 1) if k < Nc else Nc - Ns) for k in range(N)]
    return jb, tb


class RiccatiQP:
    def __init__(self, Ac, BJ, BT, c, dt, Qd, xref, Rqd, gq, w_t, w_i, vbar, pinned, vmin, vmax, x0,
                 N, Ns, Nc):
        self.__dict__.update(locals())
        self.jb, self.tb = block_maps(N, Ns, Nc)
        self.nblk = Nc - Ns + 1
        self.n_factor = 0
        self.n_solve = 0

    # ---- per-knot affine map z+ = T z + t over (x, v, d) -----------------------------------------
    def _T(self, k):
        T = np.eye(NZ)
        T[:NX, :NX] += self.dt[k] * self.Ac
        T[:NX, NX:NX + NT4] = self.dt[k] * self.BT
        T[:NX, NX + NT4:] = self.dt[k] * self.BJ
        t = np.zeros(NZ)
        t[:NX] = self.dt[k] * self.c
        return T, t

    def kind(self, k):
        if k == 0:
            return "0"
        new_j = self.jb[k] != self.jb[k - 1]
        new_t = self.tb[k] != self.tb[k - 1]
        if new_j and new_t:
            return "M"
        if new_j:
            return "H"
        assert not new_t
        return "T"

    # ---- factorisation: matrix part of the backward recursion ------------------------------------
    def factor(self):
        self.n_factor += 1
        N = self.N
        P = np.zeros((NZ, NZ))          # over (x+, v, d)
        self.K = [None] * N             # feedback gains
        self.Hinv = [None] * N
        self.Pst = [None] * (N + 1)     # value matrices before adding Q (for the vector pass)
        self.Pst[N] = P.copy()
        for k in range(N - 1, -1, -1):
            T, _ = self._T(k)
            Pt = P.copy()
            Pt[:NX, :NX] += np.diag(self.Qd)
            Phi = T.T @ Pt @ T           # over (x, v, d)
            kind = self.kind(k)
            ix = np.arange(NX)
            iv = np.arange(NX, NX + NT4)
            idd = np.arange(NX + NT4, NZ)
            Pn = np.zeros((NZ, NZ))
            if kind == "T":
                Pn = Phi
            elif kind in ("H", "0"):
                y = np.concatenate([ix, iv])
                Hdd = Phi[np.ix_(idd, idd)] + np.diag(self.Rqd)
                Hi = np.linalg.inv(Hdd)
                Kk = Hi @ Phi[np.ix_(idd, y)]                       # (8, 30)
                Pn[np.ix_(y, y)] = Phi[np.ix_(y, y)] - Phi[np.ix_(y, idd)] @ Kk
                self.K[k], self.Hinv[k] = Kk, Hi
            else:  # "M": eliminate u=(v,d) with Laplacian coupling to vprev
                u = np.concatenate([iv, idd])
                Huu = Phi[np.ix_(u, u)].copy()
                Huu[:NT4, :NT4] += self.w_t * np.eye(NT4)
                Huu[NT4:, NT4:] += np.diag(self.Rqd)
                Hi = np.linalg.inv(Huu)
                Huy = np.zeros((NT4 + NJ, NX + NT4))                # y = (x, vprev)
                Huy[:, :NX] = Phi[np.ix_(u, ix)]
                Huy[:NT4, NX:] = -self.w_t * np.eye(NT4)
                Kk = Hi @ Huy                                        # (12, 30)
                Pyy = np.zeros((NX + NT4, NX + NT4))
                Pyy[:NX, :NX] = Phi[np.ix_(ix, ix)]
                Pyy[NX:, NX:] = self.w_t * np.eye(NT4)
                Pyy -= Huy.T @ Kk
                y = np.concatenate([ix, iv])
                Pn[np.ix_(y, y)] = Pyy
                self.K[k], self.Hinv[k] = Kk, Hi
            P = Pn
            self.Pst[k] = P.copy()
        # V_0(x0, v0): P over (x, v); free-v0 system matrix
        self.M0 = P[NX:NX + NT4, NX:NX + NT4] + self.w_i * np.eye(NT4)
        self.M0inv = np.linalg.inv(self.M0)

    # ---- vector pass + forward rollout -------------------------------------------------------------
    def solve(self, gamma=None, homogeneous=False):
        """Solve with extra linear cost gamma[b]'v_b.  ``homogeneous``: drop every affine term
        (x0=0, c=0, xref=0, gq=0, vbar=0) so the result is the linear response to gamma alone."""
        self.n_solve += 1
        N = self.N
        hom = homogeneous
        if gamma is None:
            gamma = np.zeros((self.nblk, NT4))
        p = np.zeros(NZ)
        kff = [None] * N
        for k in range(N - 1, -1, -1):
            T, t = self._T(k)
            if hom:
                t = np.zeros(NZ)
            Pt = self.Pst[k + 1].copy()
            Pt[:NX, :NX] += np.diag(self.Qd)
            pt = p.copy()
            if not hom:
                pt[:NX] -= self.Qd * self.xref[k]
            phi = T.T @ (pt + Pt @ t)
            kind = self.kind(k)
            pn = np.zeros(NZ)
            if kind == "T":
                pn = phi
            elif kind in ("H", "0"):
                hd = phi[NX + NT4:] + (0 if hom else self.gq)
                kff[k] = self.Hinv[k] @ hd
                pn[:NX + NT4] = phi[:NX + NT4] - self.K[k].T @ hd
                if kind == "0":
                    pn[NX:NX + NT4] += gamma[0]
            else:
                hu = phi[NX:].copy()
                hu[:NT4] += gamma[self.tb[k]]
                if not hom:
                    hu[NT4:] += self.gq
                kff[k] = self.Hinv[k] @ hu
                pn[:NX] = phi[:NX]
                pn[:NX + NT4] -= self.K[k].T @ hu
            p = pn
        x = np.zeros((N + 1, NX))
        dq = np.zeros((self.Nc, NJ))
        v = np.zeros((self.nblk, NT4))
        x[0] = 0 if hom else self.x0
        vb = np.zeros(NT4) if hom else self.vbar
        P0 = self.Pst[0]
        if self.pinned:
            v[0] = vb
        else:
            rhs = P0[NX:NX + NT4, :NX] @ x[0] + p[NX:NX + NT4] - self.w_i * vb
            v[0] = -self.M0inv @ rhs
        for k in range(N):
            kind = self.kind(k)
            if kind in ("H", "0"):
                y = np.concatenate([x[k], v[0]])
                dq[self.jb[k]] = -self.K[k] @ y - kff[k]
            elif kind == "M":
                y = np.concatenate([x[k], v[self.tb[k] - 1]])
                u = -self.K[k] @ y - kff[k]
                v[self.tb[k]] = u[:NT4]
                dq[self.jb[k]] = u[NT4:]
            T, t = self._T(k)
            if hom:
                t = np.zeros(NZ)
            zk = np.concatenate([x[k], v[self.tb[k]], dq[self.jb[k]]])
            x[k + 1] = (T @ zk + t)[:NX]
        return x, dq, v

    # ---- Goldfarb-Idnani dual active set on the throttle boxes ------------------------------------
    def solve_box(self, max_iter=200, tol=1e-10):
        self.factor()
        x, dq, v = self.solve()
        first = 1 if self.pinned else 0            # pinned block 0 is a parameter, not a variable
        nv = (self.nblk - first) * NT4
        vv = v[first:].reshape(-1).copy()
        lo, up = self.vmin, self.vmax
        Gcol = {}                                   # lazily computed columns of G = H_r^{-1}

        def col(i):
            if i not in Gcol:
                gam = np.zeros((self.nblk, NT4))
                gam[first + i // NT4, i % NT4] = 1.0
                _, _, vh = self.solve(gam, homogeneous=True)
                Gcol[i] = -vh[first:].reshape(-1)   # minimiser moves by -G e_i per unit gradient
            return Gcol[i]

        W, sgn, lam = [], [], []                    # active set, constraint sign (+1 upper, -1 lower), multipliers
        status = 0
        it = 0
        while True:
            viol_up = vv - up
            viol_lo = lo - vv
            viol = np.maximum(viol_up, viol_lo)
            for i in W:
                viol[i] = -np.inf
            p_idx = int(np.argmax(viol))
            if viol[p_idx] <= tol:
                break
            s = 1.0 if viol_up[p_idx] > viol_lo[p_idx] else -1.0   # constraint s*v_p <= s*bound
            lam_p = 0.0
            while True:
                it += 1
                if it > max_iter:
                    status = 1
                    break
                gp = col(p_idx)
                if W:
                    GWW = np.array([[col(j)[i] * sgn[a] * sgn[b] for b, j in enumerate(W)]
                                    for a, i in enumerate(W)])
                    GWp = np.array([gp[i] * sgn[a] * s for a, i in enumerate(W)])
                    r = np.linalg.solve(GWW, GWp)
                    zdir = s * gp - sum(r[a] * sgn[a] * col(j) for a, j in enumerate(W))
                else:
                    r = np.zeros(0)
                    zdir = s * gp
                # moving along -zdir*t decreases s*v_p at rate zp = s*zdir[p] > 0
                zp = s * zdir[p_idx]
                t2 = (s * vv[p_idx] - s * (up if s > 0 else lo)) / zp if zp > 1e-300 else np.inf
                t1, drop = np.inf, -1
                for a in range(len(W)):
                    if r[a] > 0 and lam[a] / r[a] < t1:
                        t1, drop = lam[a] / r[a], a
                t = min(t1, t2)
                if not np.isfinite(t):
                    status = 2
                    break
                vv = vv - t * zdir
                for a in range(len(W)):
                    lam[a] -= t * r[a]
                lam_p += t
                if t == t2:
                    W.append(p_idx)
                    sgn.append(s)
                    lam.append(lam_p)
                    break
                W.pop(drop)
                sgn.pop(drop)
                lam.pop(drop)
            if status:
                break
        self.active = list(zip(W, sgn, lam))
        self.status = status
        if W and status == 0:
            gam = np.zeros((self.nblk, NT4))
            for i, s, l in zip(W, sgn, lam):
                gam[first + i // NT4, i % NT4] += s * l
            x, dq, v = self.solve(gam)
            for i, s in zip(W, sgn):                # land exactly on the bound
                v[first + i // NT4, i % NT4] = up if s > 0 else lo
        return x, dq, v

    def pack_z(self, x, dq, v):
        return np.concatenate([x.reshape(-1), dq.reshape(-1), v.reshape(-1)])
