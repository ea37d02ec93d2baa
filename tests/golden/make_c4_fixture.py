"""Golden vector of BASELINE.json config 4 at a quarter of the named row count (m = 2^20, n = 256, 64 equalities):
one oracle solve on the full m x n Jacobian (numpy + SciPy LAPACK restatement of Enlsip.jl).  The named size
(m = 2^22) needs ~50 GB of host memory in the oracle; the GPU tests cover it through size-independent properties.

    python tests/golden/make_c4_fixture.py [log2_rows]     ->  tests/golden/c4_1M_oracle.npz
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
    lg = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    m = 1 << lg
    d = E.synth.gen_single_index(m, 256, 64, seed=4)
    t0 = time.time()
    r = O.solve(P.single_index(d["W"], d["y"], d["rho"], d["x0"]), wallclock=False)
    dt = time.time() - t0
    trace = np.array([(t.t, t.rankA, t.rankJ2, t.dimA, t.dimJ2, t.code) for t in r.trace], dtype=np.int32)
    xs = np.array([t.x_new for t in r.trace])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "c4_1M_oracle.npz" if lg == 20 else "c4_2p%d_oracle.npz" % lg),
                        x=r.x, f=r.f, exit_code=r.exit_code, iterations=r.iterations, trace=trace, x_iter=xs,
                        active=np.array(sorted(r.active), dtype=np.int32), oracle_seconds=dt, cores=os.cpu_count(), m=m)
    print(json.dumps({"m": m, "exit_code": r.exit_code, "iterations": r.iterations, "f": r.f, "seconds": dt}))
