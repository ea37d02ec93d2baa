"""Golden vector of BASELINE.json config 5 at the NAMED size (n = 4096, m = 16384, 1024 inequalities + 8192 bounds):
one oracle solve (numpy + SciPy LAPACK restatement of Enlsip.jl, oracle/enlsip_oracle.py).  About 10 minutes on 8 cores.

    python tests/golden/make_c5_fixture.py        ->  tests/golden/c5_full_oracle.npz
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import enlsip_jl_b200 as E                                   # noqa: E402
from oracle import enlsip_oracle as O, problems as P         # noqa: E402

if __name__ == "__main__":
    d = E.synth.gen_single_index(16384, 4096, 1024, seed=5, ineq=True)
    t0 = time.time()
    r = O.solve(P.single_index(d["W"], d["y"], d["rho"], d["x0"], ineq=True, bounds=(-2.0, 2.0)), wallclock=False)
    dt = time.time() - t0
    trace = np.array([(t.t, t.rankA, t.rankJ2, t.dimA, t.dimJ2, t.code) for t in r.trace], dtype=np.int32)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "c5_full_oracle.npz"), x=r.x, f=r.f, exit_code=r.exit_code,
                        iterations=r.iterations, trace=trace, active=np.array(sorted(r.active), dtype=np.int32),
                        oracle_seconds=dt, cores=os.cpu_count())
    print(json.dumps({"exit_code": r.exit_code, "iterations": r.iterations, "f": r.f, "seconds": dt}))
