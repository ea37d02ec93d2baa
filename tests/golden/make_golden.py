"""Generate the committed golden fixtures from the CPU oracle (run in the build container).

    python tests/golden/make_golden.py

Writes tests/golden/{c2_hs65,c3_gp_fd,c3_gp_analytic}.npz : per-problem oracle outputs for the first
problems of the C2 / C3 synthetic streams, plus the per-iteration discrete trace, and
tests/golden/reference_problems.json : oracle traces of the reference's own fixtures (HS65,
Osborne 2, chained Rosenbrock n=10/100, chained Wood n=20).
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import enlsip_jl_b200 as E                                   # noqa: E402  (synth only; no GPU needed)
from oracle import enlsip_oracle as O, problems as P         # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
MAXIT = 40


def pack(results, n, lmax):
    B = len(results)
    d = dict(x=np.zeros((B, n)), f=np.zeros(B), exit_code=np.zeros(B, np.int32), status=np.zeros(B, np.int32),
             iters=np.zeros(B, np.int32), active=np.zeros((B, lmax), np.int32), nfe=np.zeros(B, np.int32),
             njac=np.zeros(B, np.int32), x_pen=np.zeros((B, n)), last_step=np.zeros(B),
             trace=np.zeros((B, MAXIT, 8), np.int32), ntrace=np.zeros(B, np.int32))
    for b, r in enumerate(results):
        d["x"][b] = r.x; d["f"][b] = r.f; d["exit_code"][b] = r.exit_code; d["status"][b] = r.status
        d["iters"][b] = r.iterations; d["active"][b, :len(r.active)] = r.active
        d["nfe"][b] = r.nb_function_evaluations; d["njac"][b] = r.nb_jacobian_evaluations
        d["ntrace"][b] = len(r.trace)
        if len(r.trace) >= 2:
            d["x_pen"][b] = r.trace[-2].x_new
        if r.trace:
            d["last_step"][b] = 3.0 * r.trace[-1].p_norm   # alpha <= 3 (EF:2176): bound on the last move
        for k, tr in enumerate(r.trace[:MAXIT]):
            d["trace"][b, k] = (tr.t, tr.rankA, tr.rankJ2, tr.dimA, tr.dimJ2, tr.code, tr.index_del, tr.exit_code)
    return d


def main():
    B2, B3 = 256, 192
    x0 = E.synth.gen_hs65_batch(B2)
    res = [O.solve(P.hs65(x0[b]), wallclock=False) for b in range(B2)]
    np.savez_compressed(os.path.join(HERE, "c2_hs65.npz"), **pack(res, 3, 7))
    y, S, x0g, _ = E.synth.gen_gauss_peaks_batch(B3)
    for fd, name in ((True, "c3_gp_fd"), (False, "c3_gp_analytic")):
        res = [O.solve(P.gauss_peaks(y[b], S[b], x0g[b], fd=fd), wallclock=False) for b in range(B3)]
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **pack(res, 6, 13))
    ref = {}
    for prob, kw in ((P.hs65(), {}), (P.osborne2(), {}), (P.chained_rosenbrock(10), {}), (P.chained_rosenbrock(100), {}),
                     (P.chained_wood(20), dict(rel_tol=1e-5, x_tol=1e-3, c_tol=1e-6))):
        r = O.solve(prob, wallclock=False, **kw)
        ref[prob.name] = {"exit_code": r.exit_code, "status": r.status, "iterations": r.iterations, "f": r.f,
                          "x": [float(v) for v in r.x], "active": r.active,
                          "trace": [[tr.t, tr.rankA, tr.rankJ2, tr.dimA, tr.dimJ2, tr.code, tr.index_del, tr.exit_code]
                                    for tr in r.trace],
                          "f_trace": [tr.f_new for tr in r.trace]}
    json.dump(ref, open(os.path.join(HERE, "reference_problems.json"), "w"), indent=1)
    print("golden fixtures written")


if __name__ == "__main__":
    main()
