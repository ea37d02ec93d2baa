"""GPU parity tests proper: the CUDA engine, called through the C ABI, against the oracle fixtures.

Run on the B200 box:  python -m pytest tests -m gpu -x -q
Nothing here reads /root/reference; the oracle (oracle/) is used only as the checker.
"""
import numpy as np
import pytest

from tests import parity

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def E():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import enlsip_jl_b200 as E
    E.capi.lib()
    return E


def _outputs(m):
    return dict(x=np.asarray(m.sol), f=np.asarray(m.obj_value), exit_code=np.asarray(m.exit_code), status=np.asarray(m.status_code),
                iters=np.asarray(m.iterations), active=np.asarray(m.active), trace=np.asarray(m.trace), counters=np.asarray(m.counters))


def test_det_exp_bit_parity(E):
    """The CUDA det_exp and oracle/detmath.c produce identical bits (the synthetic families are defined through it)."""
    from oracle import problems as P
    x = np.concatenate([np.random.default_rng(1).uniform(-750, 710, 1_000_000), np.linspace(-3, 3, 100001),
                        [0.0, -0.0, 709.78, -745.1, -744.0, -708.4, 1e-300, np.inf, -np.inf]])
    y = np.empty_like(x)
    E.capi.check(E.capi.lib().enlsipb200_det_exp(x.ctypes.data, y.ctypes.data, x.size, 0))
    assert np.array_equal(y.view(np.uint64), P.det_exp(x).view(np.uint64))


def test_c1_hs65_single_solve(E):
    """BASELINE config 1: the README / test/problems/HS65.jl solve."""
    from oracle import enlsip_oracle as O, problems as P
    m = E.CnlsModel("hs65", E.synth.HS65_X0[None, :].copy(), x_low=E.synth.HS65_LOW, x_upp=E.synth.HS65_UPP)
    assert E.total_nb_constraints(m) == 7 and E.status(m) == ["unsolved"]
    E.solve(m, trace_cap=40)
    o = O.solve(P.hs65(), wallclock=False)
    assert E.status(m) == ["found_first_order_stationary_point"] and int(m.exit_code[0]) == o.exit_code
    assert int(m.iterations[0]) == o.iterations and [int(v) for v in m.active[0] if v > 0] == o.active
    assert abs(E.sum_sq_residuals(m)[0] - 0.9535288567) < 1.5e-8          # docs/src/tutorial.md:128, 209-211
    assert abs(m.obj_value[0] - o.f) <= 1e-12 * o.f
    for k, tr in enumerate(o.trace):
        row = m.trace[0, k]
        assert (tr.t, tr.rankA, tr.rankJ2, tr.dimA, tr.dimJ2, tr.code, tr.exit_code) == tuple(int(v) for v in (row[1], row[2], row[3], row[4], row[5], row[6], row[10]))
        if k < len(o.trace) - 1:
            assert np.allclose(row[16:19], tr.x_new, rtol=1e-10, atol=0)
    assert m.trace[0, 0, 2] == 2 and m.trace[0, 0, 1] == 3               # starts rank deficient
    cv = E.constraints_values(m)[0]
    assert np.allclose(cv, np.concatenate([m.sol[0] - E.synth.HS65_LOW, E.synth.HS65_UPP - m.sol[0]]))


def test_c2_hs65_batch_vs_golden(E, golden_dir):
    gold = np.load(golden_dir + "/c2_hs65.npz")
    B = gold["x"].shape[0]
    m = E.CnlsModel("hs65", E.synth.gen_hs65_batch(B), x_low=E.synth.HS65_LOW, x_upp=E.synth.HS65_UPP)
    E.solve(m, trace_cap=40)
    out = _outputs(m)
    st = parity.compare(gold, out, "analytic", 3)
    print("C2 parity", st)
    assert (out["exit_code"] == -98).sum() == (gold["exit_code"] == -98).sum() > 0
    ok = (out["exit_code"] == gold["exit_code"]) & (out["iters"] == gold["iters"])
    assert np.array_equal(out["counters"][ok][:, 1], gold["njac"][ok])


@pytest.mark.parametrize("mode,fixture,jac", [("analytic", "c3_gp_analytic.npz", "analytic"), ("fd", "c3_gp_fd.npz", "forward_diff")])
def test_c3_gauss_peaks_vs_golden(E, golden_dir, mode, fixture, jac):
    gold = np.load(golden_dir + "/" + fixture)
    B = gold["x"].shape[0]
    y, S, x0, _ = E.synth.gen_gauss_peaks_batch(B)
    m = E.CnlsModel("gauss_peaks", x0, data={"y": y, "S": S}, x_low=E.synth.GP_LOW, x_upp=E.synth.GP_UPP, jacobian=jac)
    E.solve(m, trace_cap=40)
    st = parity.compare(gold, _outputs(m), mode, 6)
    print("C3 parity", mode, st)


@pytest.mark.parametrize("fixture,family,jac,n", [("c2_hs65_10k.npz", "hs65", "analytic", 3),
                                                   ("c3_gp_analytic_10k.npz", "gauss_peaks", "analytic", 6),
                                                   ("c3_gp_fd_10k.npz", "gauss_peaks", "forward_diff", 6)])
def test_10k_sample_histogram(E, golden_dir, fixture, family, jac, n):
    """10^4-problem oracle samples of C2 / C3 (tests/golden/make_golden_10k.py): the full
    histogram of status / exit-code / iteration / working-set / trace mismatches and of the x, f errors, printed and
    written to gpurun_out/ (committed under profiles/); bars as for the small fixtures."""
    import json
    import os
    from tests.test_hostport import check_10k_bars
    gold = np.load(golden_dir + "/" + fixture)
    B, start = gold["x"].shape[0], int(gold["start"])
    if family == "hs65":
        m = E.CnlsModel("hs65", E.synth.gen_hs65_batch(B, start=start), x_low=E.synth.HS65_LOW, x_upp=E.synth.HS65_UPP)
    else:
        y, S, x0, _ = E.synth.gen_gauss_peaks_batch(B, start=start)
        m = E.CnlsModel("gauss_peaks", x0, data={"y": y, "S": S}, x_low=E.synth.GP_LOW, x_upp=E.synth.GP_UPP, jacobian=jac)
    E.solve(m, trace_cap=40)
    h = parity.histogram(gold, _outputs(m), n)
    print("\n10k histogram (B200)", fixture, json.dumps(h))
    out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    json.dump(h, open(os.path.join(out_dir, "parity_hist_" + fixture.replace(".npz", ".json")), "w"), indent=1)
    check_10k_bars(h, 0 if jac == "analytic" else 1)


def test_c3_fresh_samples_vs_live_oracle(E):
    """Problems beyond the committed fixtures, checked against the oracle run on the box."""
    from oracle import enlsip_oracle as O, problems as P
    y, S, x0, _ = E.synth.gen_gauss_peaks_batch(24, start=1_000_000)
    m = E.CnlsModel("gauss_peaks", x0, data={"y": y, "S": S}, x_low=E.synth.GP_LOW, x_upp=E.synth.GP_UPP, jacobian="analytic")
    E.solve(m, trace_cap=40)
    for b in range(24):
        o = O.solve(P.gauss_peaks(y[b], S[b], x0[b], fd=False), wallclock=False)
        assert int(m.status_code[b]) == o.status and int(m.iterations[b]) == o.iterations
        assert [int(v) for v in m.active[b] if v > 0] == o.active
        assert abs(m.obj_value[b] - o.f) <= 1e-10 * abs(o.f)
        if len(o.trace) >= 2:
            xp = m.trace[b, len(o.trace) - 2, 16:22]
            assert np.linalg.norm(xp - o.trace[-2].x_new) <= 1e-10 * np.linalg.norm(xp)


def test_options_time_limit_max_iter_scaling(E):
    x0 = E.synth.gen_hs65_batch(64)
    m = E.CnlsModel("hs65", x0, x_low=E.synth.HS65_LOW, x_upp=E.synth.HS65_UPP)
    E.solve(m, time_limit=-1.0)                       # test/problems/chained_rosenbrock.jl:71-73 behaviour
    assert set(E.status(m)) == {"time_limit_exceeded"} and np.array_equal(m.sol, x0) and np.all(m.iterations == 1)
    E.solve(m, max_iter=3)
    st = np.asarray(m.status_code)
    assert set(E.status(m)) <= {"maximum_iterations_exceeded", "failed"} and np.mean(st == -2) > 0.8
    assert np.all(np.asarray(m.iterations)[st == -2] == 3)
    E.solve(m, scaling=True)
    assert np.mean(np.asarray(m.status_code) == 1) > 0.85


def test_batched_row_scaling_vs_live_oracle(E):
    """solve!(model; scaling = true) (structures.jl:160-178, EF:504-506, 532-534, 739) in the batched engine against
    the oracle's trace: HS65 (bounds + 1 inequality) and Gaussian peaks (1 equality + 12 bounds), analytic Jacobians."""
    from oracle import enlsip_oracle as O, problems as P
    x0 = E.synth.gen_hs65_batch(48, start=5000)
    m = E.CnlsModel("hs65", x0, x_low=E.synth.HS65_LOW, x_upp=E.synth.HS65_UPP)
    E.solve(m, trace_cap=60, scaling=True)
    y, S, x0g, _ = E.synth.gen_gauss_peaks_batch(24, start=2_000_000)
    g = E.CnlsModel("gauss_peaks", x0g, data={"y": y, "S": S}, x_low=E.synth.GP_LOW, x_upp=E.synth.GP_UPP, jacobian="analytic")
    E.solve(g, trace_cap=60, scaling=True)
    cases = [(m, lambda b: P.hs65(x0[b]), 3, 48), (g, lambda b: P.gauss_peaks(y[b], S[b], x0g[b], fd=False), 6, 24)]
    for mod, prob, n, B in cases:
        same = 0
        for b in range(B):
            o = O.solve(prob(b), wallclock=False, scaling=True)
            assert int(mod.status_code[b]) == o.status, (n, b, mod.exit_code[b], o.exit_code)
            tr_ok = int(mod.iterations[b]) == o.iterations and int(mod.exit_code[b]) == o.exit_code and \
                [int(v) for v in mod.active[b] if v > 0] == o.active
            if tr_ok:
                for k, tr in enumerate(o.trace[:60]):
                    row = mod.trace[b, k]
                    tr_ok = tr_ok and (tr.t, tr.rankA, tr.rankJ2, tr.dimA, tr.dimJ2, tr.code) == tuple(int(v) for v in row[1:7])
            if tr_ok:
                same += 1
                if o.exit_code > -90:
                    assert abs(mod.obj_value[b] - o.f) <= 1e-10 * max(abs(o.f), 1e-300), (n, b)
                    if 2 <= len(o.trace) <= 60:
                        xp = mod.trace[b, len(o.trace) - 2, 16:16 + n]
                        assert np.linalg.norm(xp - o.trace[-2].x_new) <= 1e-10 * np.linalg.norm(xp), (n, b)
        print("scaling=True parity: n=%d identical traces %d / %d" % (n, same, B))
        assert same >= 0.95 * B, (n, same, B)


@pytest.mark.parametrize("jac", ["analytic", "forward_diff"])
def test_evaluation_layer_bit_parity(E, jac):
    """enlsipb200_eval_batch (new_point!, EF:34-52; jac_forward_diff, cnls_model.jl:65-82) against the oracle's
    evaluation surface on the same inputs: r, J, c, A bit for bit (the families are defined through det_exp, and the
    forward differences use the same rounded operations in the same order)."""
    import torch
    from oracle import problems as P
    B = 96
    y, S, x0, _ = E.synth.gen_gauss_peaks_batch(B, start=3_000_000)
    dev = torch.device("cuda", 0)
    m = E.CnlsModel("gauss_peaks", torch.from_numpy(x0).to(dev), data={"y": torch.from_numpy(y).to(dev), "S": torch.from_numpy(S).to(dev)},
                    x_low=E.synth.GP_LOW, x_upp=E.synth.GP_UPP, jacobian=jac)
    out = {k: v.cpu().numpy() for k, v in E.evaluate(m).items()}
    for b in range(B):
        pb = P.gauss_peaks(y[b], S[b], x0[b], fd=(jac == "forward_diff"))
        assert np.array_equal(out["r"][b].view(np.uint64), pb.res(x0[b]).view(np.uint64)), b
        Jo = np.asarray(pb.jac_res(x0[b]))                         # m x n
        assert np.array_equal(out["J"][b].T.copy().view(np.uint64), np.ascontiguousarray(Jo).view(np.uint64)), b
        co, Ao = np.asarray(pb.cons(x0[b])), np.asarray(pb.jac_cons(x0[b]))
        assert np.array_equal(out["c"][b][:pb.l].view(np.uint64), co.view(np.uint64)), b
        assert np.array_equal(out["A"][b][:pb.l], Ao), b          # values (the reference's -I rows carry -0.0)
    assert m.launch_count() == 1


def test_step_kernel_from_materialised_inputs(E):
    """enlsipb200_step_batch (the batched step kernel that reads J, r, A, c from HBM) against the first
    iteration of the full solve and of the oracle: same working-set size and ranks, the Gauss-Newton direction
    p = (x1 - x0) / alpha to 1e-10."""
    import torch
    from oracle import enlsip_oracle as O, problems as P
    B = 64
    y, S, x0, _ = E.synth.gen_gauss_peaks_batch(B, start=4_000_000)
    dev = torch.device("cuda", 0)
    xd = torch.from_numpy(x0).to(dev)
    m = E.CnlsModel("gauss_peaks", xd, data={"y": torch.from_numpy(y).to(dev), "S": torch.from_numpy(S).to(dev)},
                    x_low=E.synth.GP_LOW, x_upp=E.synth.GP_UPP, jacobian="analytic")
    ev = E.evaluate(m, xd)
    st = {k: v.cpu().numpy() for k, v in E.gn_step(m, xd, ev).items()}
    E.solve(m, trace_cap=4)
    tr = m.trace.cpu().numpy()
    checked = 0
    for b in range(B):
        row = tr[b, 0]
        assert st["info"][b, 4] == 0
        assert (int(row[1]), int(row[2]), int(row[3])) == tuple(int(v) for v in st["info"][b, :3]), b
        if int(row[6]) == 1 and row[7] > 0:                      # first iteration took the Gauss-Newton direction
            p_full = (row[16:22] - x0[b]) / row[7]
            assert np.linalg.norm(st["p"][b] - p_full) <= 1e-9 * np.linalg.norm(p_full), b
            assert abs(np.linalg.norm(st["p"][b]) - row[8]) <= 1e-12 * row[8], b
            checked += 1
        if b < 8:
            o = O.solve(P.gauss_peaks(y[b], S[b], x0[b], fd=False), wallclock=False, max_iter=1)
            t0 = o.trace[0]
            assert (t0.t, t0.rankA, t0.rankJ2) == tuple(int(v) for v in st["info"][b, :3])
            if t0.code == 1:
                po = (t0.x_new - x0[b]) / t0.alpha
                assert np.linalg.norm(st["p"][b] - po) <= 1e-9 * np.linalg.norm(po), b
    assert checked >= B // 2


def test_device_buffers_and_determinism(E):
    """torch CUDA tensors (zero copy) give bit-identical results to host buffers; re-solving is idempotent; results do
    not depend on the position of a problem in the batch (the work queue hands problems to arbitrary warps)."""
    import torch
    B = 4096
    y, S, x0, _ = E.synth.gen_gauss_peaks_batch(B)
    mh = E.CnlsModel("gauss_peaks", x0, data={"y": y, "S": S}, x_low=E.synth.GP_LOW, x_upp=E.synth.GP_UPP, jacobian="forward_diff")
    E.solve(mh)
    dev = torch.device("cuda", 0)
    md = E.CnlsModel("gauss_peaks", torch.from_numpy(x0).to(dev), data={"y": torch.from_numpy(y).to(dev), "S": torch.from_numpy(S).to(dev)},
                     x_low=E.synth.GP_LOW, x_upp=E.synth.GP_UPP, jacobian="forward_diff")
    E.solve(md)
    torch.cuda.synchronize()
    assert np.array_equal(md.sol.cpu().numpy().view(np.uint64), np.asarray(mh.sol).view(np.uint64))
    assert np.array_equal(md.exit_code.cpu().numpy(), mh.exit_code) and np.array_equal(md.iterations.cpu().numpy(), mh.iterations)
    x1 = md.sol.clone()
    E.solve(md)
    torch.cuda.synchronize()
    assert torch.equal(x1, md.sol)
    perm = np.random.default_rng(3).permutation(B)
    mp = E.CnlsModel("gauss_peaks", x0[perm], data={"y": y[perm], "S": S[perm]}, x_low=E.synth.GP_LOW, x_upp=E.synth.GP_UPP, jacobian="forward_diff")
    E.solve(mp)
    assert np.array_equal(np.asarray(mp.sol).view(np.uint64), np.asarray(mh.sol)[perm].view(np.uint64))


def test_full_size_properties(E):
    """Size-independent properties at a large batch (1M HS65 = BASELINE config 2; 262144 Gaussian-peak fits)."""
    B = 1_000_000
    x0 = E.synth.gen_hs65_batch(B)
    m = E.CnlsModel("hs65", x0, x_low=E.synth.HS65_LOW, x_upp=E.synth.HS65_UPP)
    E.solve(m)
    ec, st = np.asarray(m.exit_code), np.asarray(m.status_code)
    assert set(np.unique(st)) <= {1, -1, -2, -11}
    conv = st == 1
    assert conv.mean() > 0.9
    assert np.all(np.abs(np.asarray(m.obj_value)[conv] - 0.9535288567) < 1e-6)
    lo, up = E.synth.HS65_LOW, E.synth.HS65_UPP
    xs = np.asarray(m.sol)[conv]
    assert np.all(xs >= lo - 1e-6) and np.all(xs <= up + 1e-6) and np.all(48 - (xs ** 2).sum(1) > -1e-6)
    assert 0.04 < np.mean(ec == -98) < 0.08          # the reference's endless working-set swap (EF:621-647)
    Bg = 262144
    y, S, x0g, truth = E.synth.gen_gauss_peaks_batch(Bg)
    g = E.CnlsModel("gauss_peaks", x0g, data={"y": y, "S": S}, x_low=E.synth.GP_LOW, x_upp=E.synth.GP_UPP, jacobian="forward_diff")
    E.solve(g)
    stg = np.asarray(g.status_code)
    assert np.mean(stg == 1) > 0.9
    xs = np.asarray(g.sol)[stg == 1]
    assert np.all(xs >= E.synth.GP_LOW - 1e-9) and np.all(xs <= E.synth.GP_UPP + 1e-9)
    h = xs[:, 0] / np.sqrt(xs[:, 1]) + xs[:, 3] / np.sqrt(xs[:, 4]) - S[stg == 1]
    assert np.max(np.abs(h)) < 1e-6                  # the equality constraint holds at the solutions
    f = np.asarray(g.obj_value)[stg == 1]
    assert np.median(f) < 128 * 0.01 ** 2 * 3        # residual level of the 0.01-sigma noise


def test_reference_suite_on_gpu(E):
    """test/problems/{osborne2, chained_rosenbrock (n=10), chained_wood (n=20)} + perturbed starts: identical
    termination, iteration counts, working sets and per-iteration method/rank/dimension trace; 1e-10 iterates."""
    from oracle import enlsip_oracle as O
    from tests.test_hostport import reference_suite_cases, check_against_oracle
    for fam, name, mk, xs, data, (lo, up), kw in reference_suite_cases():
        m = E.CnlsModel(name, np.ascontiguousarray(xs), data=data, x_low=lo, x_upp=up)
        E.solve(m, trace_cap=60, **kw)
        out = _outputs(m)
        for b in range(xs.shape[0]):
            o = O.solve(mk(xs[b]), wallclock=False, **kw)
            check_against_oracle(out, b, o, xs.shape[1], name)
        if name == "chained_wood20":
            assert np.any(out["trace"][:, :, 6] == 2)          # Newton steps taken
        assert set(E.status(m)) <= {"found_first_order_stationary_point", "failed", "maximum_iterations_exceeded"}
        print(name, "statuses", sorted(set(E.status(m))), "iters", list(np.asarray(m.iterations)), m.kernel_info())


def test_silent_false_prints_iteration_table(E, capsys):
    """solve!(model; silent=false) -> print_diagnosis (EF:2571-2580): one line per iteration with the reference's columns."""
    from oracle import enlsip_oracle as O, problems as P
    m = E.CnlsModel("hs65", E.synth.HS65_X0[None, :], x_low=E.synth.HS65_LOW, x_upp=E.synth.HS65_UPP)
    E.solve(m, silent=False)
    out = capsys.readouterr().out
    r = O.solve(P.hs65(), wallclock=False)
    rows = [ln for ln in out.splitlines() if ln[:4].strip().isdigit()]
    assert len(rows) == r.iterations == int(m.iterations[0])
    for ln, t in zip(rows, r.trace):
        cols = ln.split()
        assert abs(float(cols[1]) - t.f_new) <= 1e-6 * max(1.0, abs(t.f_new))
        assert abs(float(cols[4]) - t.alpha) <= 1e-2 * max(1e-3, abs(t.alpha))
    assert "found_first_order_stationary_point" in out


def test_host_buffer_pipeline_matches_device_path(E):
    """Host-buffer solves cut the batch into chunks whose uploads overlap the previous chunk's solves (B = 150 000 ->
    3 chunks, uneven tail): outputs must be bit-identical to one solve with device-resident inputs, and a second
    solve on the same handle (nothing pending any more) must reproduce them."""
    import torch
    B = 150_000
    y, S, x0, _ = E.synth.gen_gauss_peaks_batch(B)
    mh = E.CnlsModel("gauss_peaks", x0, data={"y": y, "S": S}, x_low=E.synth.GP_LOW, x_upp=E.synth.GP_UPP, jacobian="forward_diff")
    E.solve(mh)
    first = (np.asarray(mh.sol).copy(), np.asarray(mh.obj_value).copy(), np.asarray(mh.exit_code).copy(),
             np.asarray(mh.iterations).copy(), np.asarray(mh.active).copy())
    E.solve(mh)
    md = E.CnlsModel("gauss_peaks", torch.from_numpy(x0).cuda(), data={"y": torch.from_numpy(y).cuda(), "S": torch.from_numpy(S).cuda()},
                     x_low=E.synth.GP_LOW, x_upp=E.synth.GP_UPP, jacobian="forward_diff")
    E.solve(md)
    dev = (md.sol.cpu().numpy(), md.obj_value.cpu().numpy(), md.exit_code.cpu().numpy(), md.iterations.cpu().numpy(),
           md.active.cpu().numpy())
    second = (np.asarray(mh.sol), np.asarray(mh.obj_value), np.asarray(mh.exit_code), np.asarray(mh.iterations), np.asarray(mh.active))
    for a, b, c in zip(first, second, dev):
        assert np.array_equal(a, b, equal_nan=True) and np.array_equal(a, c, equal_nan=True)
    # new data through set_data travels with the next solve
    y2 = y[::-1].copy()
    mh.set_data(0, y2)
    E.solve(mh)
    md2 = E.CnlsModel("gauss_peaks", torch.from_numpy(x0).cuda(), data={"y": torch.from_numpy(y2).cuda(), "S": torch.from_numpy(S).cuda()},
                      x_low=E.synth.GP_LOW, x_upp=E.synth.GP_UPP, jacobian="forward_diff")
    E.solve(md2)
    assert np.array_equal(np.asarray(mh.obj_value), md2.obj_value.cpu().numpy(), equal_nan=True)
    assert not np.array_equal(np.asarray(mh.obj_value), first[1], equal_nan=True)


def test_empty_and_ragged_batches(E):
    """B = 0 is a no-op; batch sizes that do not fill a CTA (1, 5, 17, 33 problems at 16 problems per CTA) and sizes
    one short / one over a multiple of the grid give, problem by problem, the bits of the full batch."""
    y, S, x0, _ = E.synth.gen_gauss_peaks_batch(2400)
    full = E.CnlsModel("gauss_peaks", x0, data={"y": y, "S": S}, x_low=E.synth.GP_LOW, x_upp=E.synth.GP_UPP, jacobian="forward_diff")
    E.solve(full)
    for B in (0, 1, 5, 17, 33, 148 * 16 - 1, 148 * 16 + 1):
        m = E.CnlsModel("gauss_peaks", x0[:B], data={"y": y[:B], "S": S[:B]}, x_low=E.synth.GP_LOW, x_upp=E.synth.GP_UPP,
                        jacobian="forward_diff")
        E.solve(m)
        assert np.asarray(m.sol).shape == (B, 6)
        assert np.array_equal(np.asarray(m.sol).view(np.uint64), np.asarray(full.sol)[:B].view(np.uint64)), B
        assert np.array_equal(m.exit_code, full.exit_code[:B]) and np.array_equal(m.iterations, full.iterations[:B]), B
    x0h = E.synth.gen_hs65_batch(130)
    hf = E.CnlsModel("hs65", x0h, x_low=E.synth.HS65_LOW, x_upp=E.synth.HS65_UPP)
    E.solve(hf)
    for B in (0, 1, 63, 65):
        h = E.CnlsModel("hs65", x0h[:B], x_low=E.synth.HS65_LOW, x_upp=E.synth.HS65_UPP)
        E.solve(h)
        assert np.array_equal(np.asarray(h.sol).view(np.uint64), np.asarray(hf.sol)[:B].view(np.uint64)), B
