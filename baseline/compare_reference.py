"""Pin the oracle against iteration logs of the real Enlsip.jl (made by baseline/run_reference.jl on a machine that
has Julia; this repository's image has none, so this comparison has not been run -- see DESIGN.md section 4).

    python baseline/compare_reference.py baseline/reference_traces.json
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import enlsip_oracle as O, problems as P         # noqa: E402

STATUS = {1: "found_first_order_stationary_point", -1: "failed", -2: "maximum_iterations_exceeded", -11: "time_limit_exceeded"}


def main(path):
    ok = True
    for ref in json.load(open(path)):
        if ref["name"] == "hs65":
            prob = P.hs65()
        else:
            print("skip", ref["name"], "(no oracle problem of that name wired here)")
            continue
        r = O.solve(prob, wallclock=False)
        its = np.array(ref["iterations"])
        mine = np.array([[t.f_new, t.active_cx_sum, t.p_norm, t.alpha, t.progress] for t in r.trace][: len(its)])
        same_status = STATUS[r.status] == ref["status"].lstrip(":")
        same_len = len(its) == r.iterations
        err = np.abs(mine - its).max() / max(1.0, np.abs(its).max()) if same_len else float("inf")
        xerr = np.linalg.norm(np.array(ref["solution"]) - r.x) / np.linalg.norm(r.x)
        print("%-20s status %s  iterations %s (%d vs %d)  max |log - log_ref| %.2e  x rel err %.2e" %
              (ref["name"], same_status, same_len, r.iterations, len(its), err, xerr))
        ok = ok and same_status and same_len and err <= 1e-10 and xerr <= 1e-10
    print("ORACLE PINNED" if ok else "MISMATCH")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main(sys.argv[1]))
