#!/usr/bin/env python
"""One-line digest of a bench.py JSON line (for the tails of gpurun calls)."""
import json
import sys

for path in sys.argv[1:]:
    try:
        d = json.loads(open(path).read().strip().splitlines()[-1])
        r = d["roofline"]
        print("%s: value=%.3f M/s e2e=%.3f M/s k2=%.1f us k1=%.1f us frac=%.3f solved=%.4f" % (
            path, d["value"] / 1e6, d["e2e"]["value"] / 1e6, r["kernel_ms_per_launch"] * 1e3,
            r["linearise_ms_per_launch"] * 1e3, r["frac"], d["solved_fraction"]))
    except Exception as e:       # noqa: BLE001
        print(path, "unreadable:", e)
