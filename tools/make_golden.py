"""Regenerates tests/golden/oracle_trajectories.json from the CPU oracle (see tests/golden_cases.py for what the
vectors are and are not).  Floats are written with repr(), which round-trips binary64 exactly.

    python tools/make_golden.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from golden_cases import CASES  # noqa: E402
from oracle import oracle as orc  # noqa: E402

orc.build()
out = {}
for name, script in CASES.items():
    r = script(orc)
    out[name] = dict(status=r["status"], k=int(r["k"]), reason=r["reason"], x=[float(v) for v in r["x"]],
                     active_set=None if r["active_set"] is None else [int(v) for v in r["active_set"]])
    print("%-16s %-15s k=%-5d reason=%s" % (name, r["status"], r["k"], r["reason"]))
path = os.path.join(ROOT, "tests", "golden", "oracle_trajectories.json")
with open(path, "w") as f:
    json.dump(out, f, indent=0, sort_keys=True)
print("wrote", path, os.path.getsize(path), "bytes")
