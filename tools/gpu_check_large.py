"""GPU development check of the large-Jacobian regime (run under gpurun):
   TSQR R factor vs numpy QR, solve parity vs the oracle at small sizes, timings at the config-4 size."""
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import enlsip_jl_b200 as E                                   # noqa: E402
from oracle import enlsip_oracle as O, problems as P         # noqa: E402

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)


def peaks():
    exe = os.path.join(ROOT, "tools", "fp64_peaks")
    if os.path.exists(exe):
        out = subprocess.run([exe], capture_output=True, text=True, timeout=300).stdout.strip()
        print("fp64 peaks:", out)
        open(os.path.join(ROOT, "gpurun_out", "fp64_peaks.json"), "w").write(out + "\n")


def check_factor(m, n, nb, seed):
    d = E.synth.gen_single_index(m, n, nb, seed=seed)
    mod = E.LargeCnlsModel("single_index", d["x0"], d)
    R, bms, tms = mod.factor(d["x0"])
    th = P.det_tanh(d["W"] @ d["x0"])
    A = np.hstack([(1.0 - th * th)[:, None] * d["W"], (th - d["y"])[:, None]])
    Rn = np.linalg.qr(A, mode="r")
    e1 = np.abs(np.abs(R) - np.abs(Rn)).max() / np.abs(Rn).max()
    G = A.T @ A
    e2 = np.abs(R.T @ R - G).max() / np.abs(G).max()
    print("factor m=%d n=%d: | |R|-|R_numpy| | = %.2e   |R'R - A'A| = %.2e   build %.3f ms tsqr %.3f ms" % (m, n, e1, e2, bms, tms))
    mod.close()
    return e1, e2


def check_solve(m, n, nb, seed, ineq=False, bounds=None):
    d = E.synth.gen_single_index(m, n, nb, seed=seed, ineq=ineq)
    lo = None if bounds is None else np.full(n, bounds[0])
    up = None if bounds is None else np.full(n, bounds[1])
    mod = E.LargeCnlsModel("single_index", d["x0"], d, ineq=ineq, x_low=lo, x_upp=up)
    t0 = time.time()
    E.solve(mod, trace_cap=60)
    t1 = time.time()
    r = O.solve(P.single_index(d["W"], d["y"], d["rho"], d["x0"], ineq=ineq, bounds=bounds), wallclock=False)
    tr = mod.trace[0]
    same = int(mod.exit_code[0]) == r.exit_code and int(mod.iterations[0]) == r.iterations
    for k, t in enumerate(r.trace):
        e = tr[k]
        same = same and (int(e[1]), int(e[2]), int(e[3]), int(e[4]), int(e[5]), int(e[6])) == (t.t, t.rankA, t.rankJ2, t.dimA, t.dimJ2, t.code)
    xpen = 0.0
    if len(r.trace) >= 2:
        k = len(r.trace) - 2
        xpen = np.linalg.norm(tr[k, 16:] - r.trace[k].x_new) / np.linalg.norm(r.trace[k].x_new)
    print("solve m=%d n=%d ineq=%s: engine ec=%d it=%d f=%.12g (%.2fs) | oracle ec=%d it=%d f=%.12g | discrete-identical=%s "
          "frel=%.1e x_pen_rel=%.1e xrel=%.1e" % (m, n, ineq, mod.exit_code[0], mod.iterations[0], mod.obj_value[0], t1 - t0,
                                                     r.exit_code, r.iterations, r.f, same, abs(mod.obj_value[0] - r.f) / r.f, xpen,
                                                     np.linalg.norm(mod.sol[0] - r.x) / np.linalg.norm(r.x)))
    print("   stats", mod.stats())
    mod.close()


def full_size(m=1 << 22, n=256, nb=64, reps=3):
    import torch
    g = torch.Generator(device="cuda").manual_seed(4)
    W = torch.randn(m, n, dtype=torch.float64, device="cuda", generator=g) / np.sqrt(n)
    truth = torch.rand(n, dtype=torch.float64, device="cuda", generator=g) * 2 - 1
    y = torch.tanh(W @ truth) + 0.01 * torch.randn(m, dtype=torch.float64, device="cuda", generator=g)
    x0 = (truth * (1 + 0.05 * (torch.rand(n, dtype=torch.float64, device="cuda", generator=g) * 2 - 1))).cpu().numpy()
    tr = truth.cpu().numpy()
    rho = (tr[:4 * nb] ** 2).reshape(nb, 4).sum(axis=1)
    mod = E.LargeCnlsModel("single_index", x0, {"W": W, "y": y, "rho": rho})
    for i in range(reps):
        _, bms, tms = mod.factor(x0, want_R=False)
        print("full size m=%d n=%d: build %.2f ms  tsqr %.2f ms  -> %.2f TFLOP/s (2 m (n+1)^2)" % (m, n, bms, tms, 2.0 * m * (n + 1) ** 2 / tms / 1e9))
    t0 = time.time()
    E.solve(mod)
    t1 = time.time()
    st = mod.stats()
    print("full solve: ec=%d iterations=%d f=%.10g wall %.3fs ; stats %s" % (mod.exit_code[0], mod.iterations[0], mod.obj_value[0], t1 - t0, st))
    json.dump({"m": m, "n": n, "build_ms": bms, "tsqr_ms": tms, "solve_wall_s": t1 - t0, "iterations": int(mod.iterations[0]),
               "exit_code": int(mod.exit_code[0]), "stats": st}, open(os.path.join(ROOT, "gpurun_out", "large_full.json"), "w"))
    mod.close()


if __name__ == "__main__":
    peaks()
    check_factor(5000, 32, 8, 1)
    check_factor(70000, 64, 16, 2)
    check_factor(20000, 256, 64, 3)
    check_solve(2048, 32, 8, 4)
    check_solve(4096, 64, 16, 7)
    check_solve(2048, 32, 8, 5, ineq=True, bounds=(-2.0, 2.0))
    if "--full" in sys.argv:
        full_size()
