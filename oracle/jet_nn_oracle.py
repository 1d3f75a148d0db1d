"""TEST INFRASTRUCTURE — CPU restatement of the reference's jet plant + estimator pair:

* neural jet plant: ``NeuralJetModel.get_state`` / ``JetModelTotal.get_state`` (src/mujoco_lib/nn_jet_model.py:21-30,
  86-109).  The reference feeds the LSTM one time step with a fresh zero state on every call, so the cell reduces to
  c = sigma(i) * tanh(g), h = sigma(o) * tanh(c) with gates = W_ih x + b_ih + b_hh; thrust rate = fc(h); float32.
  PINNED against outputs of the reference module itself (tests/golden/jet_nn.npz, tests/golden/make_jet_nn_golden.py).
* per-jet EKF: ``SecondOrderJetModel.update`` (src/mujoco_lib/jet_kalman_filter.py:30-65): predict with the second-order
  jet model, covariance with the Jacobian evaluated at the PREDICTED state (:58), measurement = (T_nn, Tdot_nn), H = I.
  PINNED against outputs of the reference file itself, executed unmodified over a CasADi stand-in
  (tests/golden/jet_ekf.npz, tests/golden/make_jet_ekf_golden.py; the Jacobian here is written out analytically).
Only tests/ and bench.py's cpu_baseline may import this module."""
from __future__ import annotations

import numpy as np

from .vsmpc_oracle import JetModel


def _sig(x):
    return (1.0 / (1.0 + np.exp(-x.astype(np.float32)))).astype(np.float32)


def nn_jet_step(T, u, w, dt):
    """T, u: float32 arrays (n,). w: dict with w_ih (320,2), b_ih, b_hh (320,), fc_w (80,), fc_b (1,), norm (4,)."""
    mean_T, std_T, mean_u, std_u = [float(x) for x in w["norm"]]
    T = np.asarray(T, np.float32); u = np.asarray(u, np.float32)
    Tn = ((T.astype(np.float64) - mean_T) / std_T).astype(np.float32)          # python-float arithmetic, then float32
    un = ((u.astype(np.float64) - mean_u) / std_u).astype(np.float32)
    x = np.stack([Tn, un], axis=1)                                              # (n, 2)
    gates = (x @ w["w_ih"].T.astype(np.float32) + w["b_ih"] + w["b_hh"]).astype(np.float32)
    i, g, o = _sig(gates[:, 0:80]), np.tanh(gates[:, 160:240]).astype(np.float32), _sig(gates[:, 240:320])
    c = (i * g).astype(np.float32)
    h = (o * np.tanh(c)).astype(np.float32)
    td = (h @ w["fc_w"].astype(np.float32) + w["fc_b"][0]).astype(np.float32)
    T_next_norm = (Tn + td * np.float32(dt)).astype(np.float32)
    return (T_next_norm * np.float32(std_T) + np.float32(mean_T)).astype(np.float32), (td * np.float32(std_T)).astype(np.float32)


class JetEKF:
    """One jet: jet_kalman_filter.py:4-65 with the constants of :6-22 (the same 13 coefficients as JetModel.cpp)."""

    def __init__(self, R, Q, P, dt, jet_model: JetModel = None):
        self.jet = jet_model or JetModel()
        self.R, self.Q, self.P, self.dt = np.array(R, float), np.array(Q, float), np.array(P, float), float(dt)

    def f(self, x, u):
        j, dt = self.jet, self.dt
        sig = j.getThrustStandardDeviation_u2T()
        Ts, Tds = j.standardizeThrust_u2T(x[0]), j.standardizeThrustDot_u2T(x[1])
        v = j.compute_v(j.standardizeThrottle_u2T(u))
        tdd = j.compute_f(Ts, Tds) + j.compute_g(Ts, Tds) * v
        Td = x[1] + tdd * sig * dt
        return np.array([x[0] + Td * dt, Td])

    def A(self, x, u):
        j, dt = self.jet, self.dt
        Ts, Tds = j.standardizeThrust_u2T(x[0]), j.standardizeThrustDot_u2T(x[1])
        v = j.compute_v(j.standardizeThrottle_u2T(u))
        hT = j.compute_df_dT(Ts, Tds) + j.compute_dg_dT(Ts, Tds) * v       # d(Tdd_norm)/d(T_norm)
        hTd = j.compute_df_dTdot(Ts, Tds) + j.compute_dg_dTdot(Ts, Tds) * v
        a10, a11 = dt * hT, 1.0 + dt * hTd                                  # the sigma_T factors cancel
        return np.array([[1.0 + dt * a10, dt * a11], [a10, a11]])

    def update(self, x, u, z):
        x = self.f(np.asarray(x, float), u)
        A = self.A(x, u)                       # evaluated at the predicted state, as the reference does (:58)
        self.P = A @ self.P @ A.T + self.Q
        S = self.P + self.R
        K = self.P @ np.linalg.inv(S)
        x = x + K @ (np.asarray(z, float) - x)
        self.P = (np.eye(2) - K) @ self.P
        return x
