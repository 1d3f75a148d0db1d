"""TEST INFRASTRUCTURE — a stand-in for the handful of CasADi calls the reference's per-jet EKF makes
(src/mujoco_lib/jet_kalman_filter.py: ``MX.sym``, arithmetic, ``vertcat``, ``Function``, ``jacobian``, ``DM.eye``, ``inv``,
``@`` / ``.T``), so that the reference file can be EXECUTED UNMODIFIED in a container without CasADi
(tests/golden/make_jet_ekf_golden.py registers this module as ``casadi`` before importing it).

Symbolic expressions are closures evaluated on demand; ``jacobian`` is exact forward-mode differentiation with dual numbers
(the same derivative values CasADi's symbolic differentiation produces, up to rounding), and ``DM`` is a float64 matrix.
Nothing here knows about jets or filters: every model equation and the filter update come from the reference file."""
from __future__ import annotations

import itertools

import numpy as np


class _Dual:
    """a + b·ε, ε² = 0."""
    __slots__ = ("a", "b")

    def __init__(self, a, b=0.0):
        self.a, self.b = float(a), float(b)

    @staticmethod
    def _of(o):
        return o if isinstance(o, _Dual) else _Dual(o)

    def __add__(self, o):
        o = _Dual._of(o); return _Dual(self.a + o.a, self.b + o.b)
    __radd__ = __add__

    def __sub__(self, o):
        o = _Dual._of(o); return _Dual(self.a - o.a, self.b - o.b)

    def __rsub__(self, o):
        return _Dual._of(o) - self

    def __mul__(self, o):
        o = _Dual._of(o); return _Dual(self.a * o.a, self.a * o.b + self.b * o.a)
    __rmul__ = __mul__

    def __truediv__(self, o):
        o = _Dual._of(o); return _Dual(self.a / o.a, (self.b * o.a - self.a * o.b) / (o.a * o.a))

    def __rtruediv__(self, o):
        return _Dual._of(o) / self

    def __neg__(self):
        return _Dual(-self.a, -self.b)

    def __pow__(self, n):
        assert isinstance(n, int) and n >= 1
        r = self
        for _ in range(n - 1):
            r = r * self
        return r


def _obj(v):
    """anything numeric -> 2-D object array (column vector for 1-D input)."""
    if isinstance(v, DM):
        v = v.m
    a = np.array(v, dtype=object)
    if a.ndim == 0:
        a = a.reshape(1, 1)
    elif a.ndim == 1:
        a = a.reshape(-1, 1)
    return a


class MX:
    """A symbolic matrix: ``ev(env)`` returns a 2-D object array of floats / duals given the symbol values."""
    _ids = itertools.count()

    def __init__(self, ev, key=None):
        self.ev, self.key = ev, key

    @staticmethod
    def sym(name, n=1):
        key = (name, next(MX._ids))
        return MX(lambda env: env[key], key)

    @staticmethod
    def _lift(o):
        return o if isinstance(o, MX) else MX(lambda env, c=_obj(o): c)

    def _bin(self, o, op):
        o = MX._lift(o)
        return MX(lambda env: op(self.ev(env), o.ev(env)))

    def __add__(self, o): return self._bin(o, lambda a, b: a + b)
    def __radd__(self, o): return MX._lift(o)._bin(self, lambda a, b: a + b)
    def __sub__(self, o): return self._bin(o, lambda a, b: a - b)
    def __rsub__(self, o): return MX._lift(o)._bin(self, lambda a, b: a - b)
    def __mul__(self, o): return self._bin(o, lambda a, b: a * b)
    def __rmul__(self, o): return MX._lift(o)._bin(self, lambda a, b: a * b)
    def __truediv__(self, o): return self._bin(o, lambda a, b: a / b)
    def __neg__(self): return MX(lambda env: -self.ev(env))

    def __pow__(self, n):
        return MX(lambda env: self.ev(env) ** n)

    def __getitem__(self, i):
        assert isinstance(i, int)
        return MX(lambda env: self.ev(env)[i:i + 1, :])


class DM:
    """float64 matrix with the operators the filter update uses."""
    __array_ufunc__ = None          # numpy arrays defer to the reflected operators below

    def __init__(self, v):
        m = np.array(v.m if isinstance(v, DM) else v, dtype=np.float64)
        self.m = m.reshape(1, 1) if m.ndim == 0 else (m.reshape(-1, 1) if m.ndim == 1 else m)

    @staticmethod
    def eye(n):
        return DM(np.eye(n))

    @staticmethod
    def _m(o):
        return DM(o).m

    @property
    def T(self): return DM(self.m.T)
    def __matmul__(self, o): return DM(self.m @ DM._m(o))
    def __rmatmul__(self, o): return DM(DM._m(o) @ self.m)
    def __add__(self, o): return DM(self.m + DM._m(o))
    __radd__ = __add__
    def __sub__(self, o): return DM(self.m - DM._m(o))
    def __rsub__(self, o): return DM(DM._m(o) - self.m)
    def __mul__(self, o): return DM(self.m * DM._m(o))
    __rmul__ = __mul__
    def __neg__(self): return DM(-self.m)

    def __getitem__(self, i):
        return DM(self.m.reshape(-1)[i]) if isinstance(i, int) else DM(self.m[i])

    def __float__(self):
        assert self.m.size == 1
        return float(self.m.reshape(-1)[0])

    def __array__(self, dtype=None, copy=None):
        return self.m.astype(dtype) if dtype is not None else self.m

    def full(self):
        return self.m.copy()


def vertcat(*parts):
    if any(isinstance(p, MX) for p in parts):
        ps = [MX._lift(p) for p in parts]
        return MX(lambda env: np.vstack([p.ev(env) for p in ps]))
    return DM(np.vstack([DM._m(p) for p in parts]))


def inv(a):
    return DM(np.linalg.inv(DM._m(a)))


class Function:
    def __init__(self, name, ins, outs):
        assert len(outs) == 1 and all(isinstance(s, MX) and s.key is not None for s in ins)
        self.name, self.ins, self.out = name, ins, outs[0]

    def __call__(self, *args):
        assert len(args) == len(self.ins)
        if any(isinstance(a, MX) for a in args):
            la = [MX._lift(a) for a in args]
            return MX(lambda env: self.out.ev({s.key: a.ev(env) for s, a in zip(self.ins, la)}))
        r = self.out.ev({s.key: _obj(a) for s, a in zip(self.ins, args)})
        return DM(r.astype(np.float64))


def jacobian(expr, x):
    assert isinstance(expr, MX) and isinstance(x, MX) and x.key is not None

    def ev(env):
        xv = env[x.key]
        n = xv.shape[0]
        cols = []
        for j in range(n):
            seeded = np.array([[_Dual(float(xv[i, 0]), 1.0 if i == j else 0.0)] for i in range(n)], dtype=object)
            e2 = dict(env); e2[x.key] = seeded
            y = expr.ev(e2)
            cols.append([_Dual._of(v).b for v in y[:, 0]])
        return np.array(cols, dtype=object).T
    return MX(ev)
