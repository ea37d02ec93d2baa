"""Extract the Osborne-2 data table, bounds and saved starting point from the reference test
fixture (test/problems/osborne2.jl:10-102) into tests/golden/osborne2.json.

Run in the build container only (needs /root/reference); the JSON is committed.
"""
import json
import re
import sys

src = open("/root/reference/test/problems/osborne2.jl").read()
blk = src[src.index("dataset = [") + len("dataset = ["):]
blk = blk[: blk.index("]")]
rows = []
for line in blk.split(";"):
    parts = line.split()
    if len(parts) == 3:
        rows.append([float(v) for v in parts])
assert len(rows) == 65, len(rows)


def vec(name):
    m = re.search(name + r"\s*=\s*\[(.*?)\]", src, re.S)
    return [float(v) for v in re.split(r"[,\s]+", m.group(1).strip()) if v]


out = {"t": [r[1] for r in rows], "y": [r[2] for r in rows], "x_low": vec("low_bounds"),
       "x_upp": vec("upp_bounds"), "x0": vec("x0")}
assert len(out["x0"]) == 11 and len(out["x_low"]) == 11 and len(out["x_upp"]) == 11
json.dump(out, open(sys.argv[1] if len(sys.argv) > 1 else "tests/golden/osborne2.json", "w"), indent=0)
print("ok", out["x0"][:2], out["t"][-1], out["y"][-1])
