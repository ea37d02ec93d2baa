"""GPU development check of BASELINE.json config 5 (dense wide problem: inequalities + bounds, n up to 4096):
parity with the oracle at reduced size (device QRCP / M*Q of the compressed problem in use), then the named size."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import enlsip_jl_b200 as E                                   # noqa: E402
from oracle import enlsip_oracle as O, problems as P         # noqa: E402


def run(m, n, nb, seed, oracle=True, bounds=(-2.0, 2.0), max_iter=100):
    d = E.synth.gen_single_index(m, n, nb, seed=seed, ineq=True)
    lo, up = np.full(n, bounds[0]), np.full(n, bounds[1])
    mod = E.LargeCnlsModel("single_index", d["x0"], d, ineq=True, x_low=lo, x_upp=up)
    t0 = time.time()
    E.solve(mod, trace_cap=100, max_iter=max_iter)
    t1 = time.time()
    st = mod.stats()
    k = int(mod.iterations[0])
    print("engine m=%d n=%d nb=%d: ec=%d it=%d nact=%d f=%.12g  wall %.2fs (%.3f s / executed iteration)"
          % (m, n, nb, mod.exit_code[0], k, mod.nb_active[0], mod.obj_value[0], t1 - t0, (t1 - t0) / max(st["points"] - 1, 1)))
    print("   stats", {a: round(b, 2) for a, b in st.items()})
    tr = mod.trace[0]
    print("   t / rankA / rankJ2 / code per iteration:", [(int(e[1]), int(e[2]), int(e[3]), int(e[6])) for e in tr[:min(k + 1, 12)]])
    res = {"m": m, "n": n, "nb": nb, "exit_code": int(mod.exit_code[0]), "iterations": k, "wall_s": t1 - t0, "stats": st,
           "objective": float(mod.obj_value[0])}
    if oracle:
        t0 = time.time()
        r = O.solve(P.single_index(d["W"], d["y"], d["rho"], d["x0"], ineq=True, bounds=bounds), wallclock=False, max_iter=max_iter)
        t1 = time.time()
        same = int(mod.exit_code[0]) == r.exit_code and k == r.iterations
        for i, t in enumerate(r.trace):
            e = tr[i]
            same = same and (int(e[1]), int(e[2]), int(e[3]), int(e[4]), int(e[5]), int(e[6])) == (t.t, t.rankA, t.rankJ2, t.dimA, t.dimJ2, t.code)
        xpen = 0.0
        if len(r.trace) >= 2:
            i = len(r.trace) - 2
            xpen = np.linalg.norm(tr[i, 16:] - r.trace[i].x_new) / np.linalg.norm(r.trace[i].x_new)
        print("   oracle: ec=%d it=%d f=%.12g (%.2fs, %.3f s/it) | discrete-identical=%s frel=%.1e x_pen_rel=%.1e xrel=%.1e"
              % (r.exit_code, r.iterations, r.f, t1 - t0, (t1 - t0) / max(r.iterations, 1), same, abs(mod.obj_value[0] - r.f) / r.f, xpen,
                 np.linalg.norm(mod.sol[0] - r.x) / np.linalg.norm(r.x)))
        res.update(oracle_s=t1 - t0, oracle_iterations=r.iterations, identical=bool(same))
    mod.close()
    return res


if __name__ == "__main__":
    out = [run(4096, 512, 128, 31), run(3000, 384, 64, 32)]
    if "--full" in sys.argv:
        out.append(run(16384, 4096, 1024, 5, oracle="--oracle" in sys.argv))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "c5_check.json"), "w"), indent=1)
