#!/usr/bin/env python
"""Development: replay the joint-box working-set iteration of dumped instances (tools/dump_unsettled.py) with the NumPy
specification and print the working-set changes per pass."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from condensed_model import ClampedCondensedQP, box_qp_pivot, NJ

d = np.load(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/unsettled.npz")
Hd, Htt = d["Hdiag"], d["H_tt"]
Qd, Rqd = Hd[26:52].copy(), Hd[468:476].copy()
w_t = -Htt[0, 4]
w_i = Htt[0, 0] - w_t
for j, inst in enumerate(d["idx"]):
    q, l, u = d["q"][j], d["l"][j], d["u"][j]
    xref = np.zeros((17, 26))
    for k in range(17):
        xref[k] = np.where(Qd > 0, -q[26 * (k + 1):26 * (k + 2)] / np.where(Qd > 0, Qd, 1.0), 0.0)
    vbar = -q[564:568] / w_i
    pinned = l[468] == u[468]
    m = ClampedCondensedQP(d["A"][j], d["BJ"][j], d["BT"][j], d["c"][j], d["dt"], Qd, xref, Rqd, q[468:476].copy(), w_t, w_i, vbar,
                           pinned, l[472], u[472], l[442:468].copy(), 17, 7, 12)
    lo, hi = l[512:520].copy(), u[512:520].copy()
    print(f"== instance {inst} (kernel: {d['nf'][j]} factorisations, status {d['status'][j]}) box width {np.round(hi - lo, 3).tolist()} pinned {pinned}")
    clamp = np.zeros((12, NJ), dtype=int)
    seen = {}
    for p in range(16):
        m.factor_clamped(clamp, lo, hi)
        H, g, first = m.reduced_qp()
        vv, act, st = box_qp_pivot(H, g, m.vmin, m.vmax)
        v = (np.concatenate([m.vbar, vv]) if m.pinned else vv).reshape(m.nblk, 4)
        x, dq, v, grad = m.forward_clamped(v, clamp, lo, hi)
        add_u, add_l = (clamp == 0) & (dq > hi + 1e-9), (clamp == 0) & (dq < lo - 1e-9)
        rel = ((clamp > 0) & (grad > 1e-9 * m.gmag)) | ((clamp < 0) & (grad < -1e-9 * m.gmag))
        key = clamp.tobytes()
        rep = seen.get(key)
        seen[key] = p
        adds = [(int(k), int(c), float(np.round(max(dq[k, c] - hi[c], lo[c] - dq[k, c]), 6))) for k, c in np.argwhere(add_u | add_l)]
        rels = [(int(k), int(c), float(np.round(grad[k, c] / m.gmag[k, c], 4))) for k, c in np.argwhere(rel)]
        print(f" pass {p}: clamped {int((clamp != 0).sum())} status {st} joins {adds[:8]} leaves {rels[:8]}" + (f"  <- same set as pass {rep}" if rep is not None else ""))
        new = clamp.copy()
        new[add_u] = 1; new[add_l] = -1
        if not (p >= 5 and (add_u.any() or add_l.any())):
            new[rel] = 0
        if (new == clamp).all():
            print("  settled after", p + 1)
            break
        clamp = new
