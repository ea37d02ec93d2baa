"""Shared parity criteria: engine outputs vs the committed oracle fixtures (tests/golden/*.npz).

Bars (see DESIGN.md "Parity"):
  analytic Jacobians  discrete outputs (status, exit code, iteration count, final working set and the
                      per-iteration (t, rankA, rankJ2, dimA, dimJ2, method, index_del, exit) trace) identical
                      except knife-edge cases (>= 97 % of the problems); objective within 1e-10 relative;
                      the iterate BEFORE the last step within 1e-10 relative; the final iterate within
                      1e-10 relative + the largest possible last move 3*||p_last|| (a converged solve ends with a
                      step of ~1e-9 whose steplength comes from a linesearch on rounding noise).
  forward differences the FD Jacobian (cnls_model.jl:65-82) amplifies a 1-ulp difference of x by
                      1/sqrt(eps); the oracle itself moves by ~5e-10 relative (and flips ~7 % of the raw
                      exit codes) when x0 changes by one ulp (tests/test_oracle.py::test_fd_noise_floor).
                      Bar: status identical >= 95 %, iteration count identical >= 90 %, objective within
                      1e-9, x within 1e-8 relative.
"""
import numpy as np

TRACE_HDR = 16


def compare(gold, eng, mode, n):
    """gold: npz fixture; eng: dict(x, f, exit_code, status, iters, active, trace[B,cap,16+n])."""
    gold = {k: gold[k] for k in getattr(gold, "files", gold)}     # an NpzFile decompresses an array on EVERY access
    B = gold["x"].shape[0]
    x, f = np.asarray(eng["x"])[:B], np.asarray(eng["f"])[:B]
    ec, st, it = (np.asarray(eng[k])[:B] for k in ("exit_code", "status", "iters"))
    act = np.asarray(eng["active"])[:B]
    tr = np.asarray(eng["trace"])[:B]
    same_status = st == gold["status"]
    same_iters = it == gold["iters"]
    same_all = same_status & same_iters & (ec == gold["exit_code"]) & np.all(act == gold["active"], axis=1)
    xn = np.linalg.norm(gold["x"], axis=1)
    xrel = np.linalg.norm(x - gold["x"], axis=1) / xn
    frel = np.abs(f - gold["f"]) / np.maximum(np.abs(gold["f"]), 1e-300)
    stats = dict(B=B, same_status=float(same_status.mean()), same_iters=float(same_iters.mean()),
                 same_all=float(same_all.mean()), xrel_max=float(xrel[same_all].max()),
                 frel_max=float(frel[same_all].max()))
    if mode == "analytic":
        assert same_status.all(), "termination status differs: %s" % np.nonzero(~same_status)[0][:10]
        cap = tr.shape[1]
        same_trace = np.zeros(B, dtype=bool)
        for b in np.nonzero(same_all)[0]:
            nt = min(int(gold["ntrace"][b]), cap, gold["trace"].shape[1])
            e = tr[b, :nt][:, [1, 2, 3, 4, 5, 6, 9, 10]].astype(np.int64)
            # knife-edge ties (e.g. HS65 is symmetric in x1/x2: two equal second-order multipliers, either
            # constraint may be the one deleted) count as mismatches here and must stay below 3 %
            same_trace[b] = np.array_equal(e, gold["trace"][b, :nt])
        stats["same_trace"] = float(same_trace.mean())
        assert same_trace.mean() >= 0.97, stats
        for b in np.nonzero(same_trace)[0]:
            assert frel[b] <= 1e-10, (b, frel[b])
            if gold["ntrace"][b] >= 2 and gold["ntrace"][b] <= cap:
                xp = tr[b, gold["ntrace"][b] - 2, TRACE_HDR:TRACE_HDR + n]
                rel = np.linalg.norm(xp - gold["x_pen"][b]) / np.linalg.norm(gold["x_pen"][b])
                assert rel <= 1e-10, ("iterate before the last step", b, rel)
            assert np.linalg.norm(x[b] - gold["x"][b]) <= 1e-10 * xn[b] + 1.05 * gold["last_step"][b] + 1e-300, (b, xrel[b])
    else:
        assert same_status.mean() >= 0.95, stats
        assert same_iters.mean() >= 0.90, stats
        ok = same_status & same_iters
        assert frel[ok].max() <= 1e-9, stats
        assert xrel[ok].max() <= 1e-8, stats
    return stats


def histogram(gold, eng, n, start=None):
    """Full mismatch histogram of a large sample: how many problems differ in status / raw exit
    code / iteration count / final working set / per-iteration trace, the distribution of the iteration-count
    differences, and quantiles of the relative errors of x and f over the problems whose discrete outputs agree."""
    gold = {k: gold[k] for k in getattr(gold, "files", gold)}     # an NpzFile decompresses an array on EVERY access
    B = gold["x"].shape[0]
    x, f = np.asarray(eng["x"])[:B], np.asarray(eng["f"])[:B]
    ec, st, it = (np.asarray(eng[k])[:B] for k in ("exit_code", "status", "iters"))
    act = np.asarray(eng["active"])[:B]
    same_status = st == gold["status"]
    same_exit = ec == gold["exit_code"]
    same_iters = it == gold["iters"]
    same_active = np.all(act == gold["active"], axis=1)
    agree = same_status & same_exit & same_iters & same_active
    h = {"problems": int(B), "status_mismatch": int((~same_status).sum()), "exit_code_mismatch": int((~same_exit).sum()),
         "iteration_count_mismatch": int((~same_iters).sum()), "working_set_mismatch": int((~same_active).sum()),
         "all_discrete_outputs_identical": int(agree.sum())}
    d = (it.astype(np.int64) - gold["iters"].astype(np.int64))
    vals, cnt = np.unique(d, return_counts=True)
    h["iteration_difference_histogram"] = {int(v): int(c) for v, c in zip(vals, cnt)}
    vals, cnt = np.unique(gold["exit_code"], return_counts=True)
    h["oracle_exit_codes"] = {int(v): int(c) for v, c in zip(vals, cnt)}
    if "trace" in eng and eng["trace"] is not None:
        tr = np.asarray(eng["trace"])[:B]
        cap = min(tr.shape[1], gold["trace"].shape[1])
        same_trace = np.zeros(B, dtype=bool)
        for b in np.nonzero(agree)[0]:
            nt = min(int(gold["ntrace"][b]), cap)
            same_trace[b] = np.array_equal(tr[b, :nt][:, [1, 2, 3, 4, 5, 6, 9, 10]].astype(np.int64), gold["trace"][b, :nt])
        h["trace_identical"] = int(same_trace.sum())
    ok = agree & (gold["exit_code"] > -90)
    xn = np.maximum(np.linalg.norm(gold["x"], axis=1), 1e-300)
    xrel = np.linalg.norm(x - gold["x"], axis=1) / xn
    frel = np.abs(f - gold["f"]) / np.maximum(np.abs(gold["f"]), 1e-300)
    q = [0.5, 0.9, 0.99, 0.999, 1.0]
    h["x_rel_error_quantiles"] = {str(k): float(np.quantile(xrel[ok], k)) for k in q} if ok.any() else {}
    h["f_rel_error_quantiles"] = {str(k): float(np.quantile(frel[ok], k)) for k in q} if ok.any() else {}
    h["x_rel_error_over_1e-10"] = int((xrel[ok] > 1e-10).sum())
    h["x_rel_error_over_1e-8"] = int((xrel[ok] > 1e-8).sum())
    # the final move of a converged solve is bounded by 3 ||p_last|| (flat merit function): excess over that bound
    excess = np.linalg.norm(x - gold["x"], axis=1) - (1e-10 * xn + 1.05 * gold["last_step"])
    h["x_beyond_last_step_bound"] = int((excess[ok] > 0).sum())
    return h
