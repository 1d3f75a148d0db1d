#!/usr/bin/env python
"""Golden vectors of the jet model, produced by the REFERENCE'S OWN OBJECT CODE:
oracle/build_ref.py compiles src/flight-controller/utils/src/JetModel.cpp from /root/reference into
oracle/_ref/libjetmodel_ref.so; this script evaluates it on seeded inputs and freezes inputs and outputs in
tests/golden/jet_model_ref.npz (the library cannot be rebuilt on the GPU box, the vectors travel).

    python tests/golden/make_jet_model_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import build_ref  # noqa: E402

POLY = ("compute_f", "compute_g", "compute_df_dT", "compute_df_dTdot", "compute_dg_dT", "compute_dg_dTdot")
SCALAR = ("compute_v", "standardizeThrust_u2T", "standardizeThrustDot_u2T", "standardizeThrottle_u2T",
          "destandardizeThrust_u2T", "destandardizeThrustDot_u2T", "destandardizeThrottle_u2T")


def evaluate(lib, Ts, Tds, xs):
    poly = np.array([[lib.ref_jet_poly(w, float(a), float(b)) for a, b in zip(Ts, Tds)] for w in range(len(POLY))])
    scal = np.array([[lib.ref_jet_scalar(w, float(x)) for x in xs] for w in range(len(SCALAR))])
    return poly, scal, lib.ref_jet_scalar(7, 0.0)


def inputs(n=256, seed=20251002):
    rng = np.random.default_rng(seed)
    Ts, Tds = rng.uniform(-2.0, 2.0, n), rng.normal(0.0, 0.8, n)          # standardised thrust / thrust rate
    # scalar arguments: throttles, thrusts and transformed throttles incl. both clip regions of destandardizeThrottle
    xs = np.concatenate([rng.uniform(-3.0, 3.0, n - 8), [-1.52115, 1.65096, 0.0, 100.0, 47.333, -2.5, 2.5, 1e-3]])
    return Ts, Tds, xs


if __name__ == "__main__":
    lib = build_ref.load()
    assert lib is not None, "needs /root/reference (run in the build container)"
    Ts, Tds, xs = inputs()
    poly, scal, sigma = evaluate(lib, Ts, Tds, xs)
    out = os.path.join(HERE, "jet_model_ref.npz")
    np.savez_compressed(out, T_std=Ts, Tdot_std=Tds, x=xs, poly=poly, scalar=scal, thrust_std=sigma,
                        poly_names=np.array(POLY), scalar_names=np.array(SCALAR))
    print("wrote", out, os.path.getsize(out), "bytes; sigma_T", sigma)
