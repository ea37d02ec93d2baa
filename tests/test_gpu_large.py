"""GPU parity tests of the LARGE-JACOBIAN regime (BASELINE.json config 4), through the C ABI.

  * TSQR (Householder panels + FP64 tensor-core trailing updates) against LAPACK QR of the same [J | r];
  * whole solves against the oracle, which factors the full m x n Jacobian like the reference
    (src/enlsip_functions.jl:206-234): identical discrete trace, objective and iterates to 1e-10;
  * at the full size (m = 4M, n = 256) size-independent properties: R'R = [J r]'[J r], determinism,
    the solve converges and satisfies its constraints;
  * the row-sharded path on ONE GPU: two handles joined through NCCL are exercised in bench.py --gpus 2;
    here the stacked-R re-factorisation is checked against the single-shard factor.
Run on the B200 box:  python -m pytest tests -m gpu -x -q
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def E():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import enlsip_jl_b200 as E
    E.capi.lib()
    return E


def _aug(d, x):
    from oracle import problems as P
    th = P.det_tanh(d["W"] @ x)
    return np.hstack([(1.0 - th * th)[:, None] * d["W"], (th - d["y"])[:, None]])


@pytest.mark.parametrize("m,n,nb,seed", [(5000, 32, 8, 1), (70000, 64, 16, 2), (20000, 256, 64, 3), (33, 32, 8, 6),
                                         (4097, 96, 4, 8)])
def test_tsqr_factor_vs_lapack(E, m, n, nb, seed):
    d = E.synth.gen_single_index(m, n, nb, seed=seed)
    mod = E.LargeCnlsModel("single_index", d["x0"], d, m_global=max(m, 1000))
    R, _, _ = mod.factor(d["x0"])
    mod.close()
    A = _aug(d, d["x0"])
    Rn = np.linalg.qr(A, mode="r")
    scale = np.abs(Rn).max()
    assert np.abs(np.abs(R) - np.abs(Rn)).max() <= 1e-13 * scale          # rows of R are unique up to sign
    G = A.T @ A
    assert np.abs(R.T @ R - G).max() <= 1e-13 * np.abs(G).max()
    assert np.all(np.tril(R, -1) == 0.0)


@pytest.mark.parametrize("m,n,nb,seed,ineq,bounds", [(2048, 32, 8, 4, False, None), (4096, 64, 16, 7, False, None),
                                                      (2048, 32, 8, 5, True, (-2.0, 2.0)), (3000, 32, 3, 9, True, None),
                                                      (6000, 128, 32, 12, False, None)])
def test_large_solve_vs_oracle(E, m, n, nb, seed, ineq, bounds):
    from oracle import enlsip_oracle as O, problems as P
    from tests.test_large_host import compare_with_oracle
    d = E.synth.gen_single_index(m, n, nb, seed=seed, ineq=ineq)
    lo = None if bounds is None else np.full(n, bounds[0])
    up = None if bounds is None else np.full(n, bounds[1])
    mod = E.LargeCnlsModel("single_index", d["x0"], d, ineq=ineq, x_low=lo, x_upp=up)
    E.solve(mod, trace_cap=60)
    r = O.solve(P.single_index(d["W"], d["y"], d["rho"], d["x0"], ineq=ineq, bounds=bounds), wallclock=False)
    out = dict(x=mod.sol[0], f=mod.obj_value, exit_code=mod.exit_code, status=mod.status_code, iters=mod.iterations,
               nact=mod.nb_active, active=mod.active[0], trace=mod.trace[0])
    compare_with_oracle(out, r, n)
    assert mod.launch_count() > 0
    mod.close()


def test_large_device_buffers_and_determinism(E):
    import torch
    d = E.synth.gen_single_index(50000, 64, 16, seed=21)
    res = []
    for dev in (False, True, True):
        data = dict(d)
        if dev:
            data["W"] = torch.from_numpy(d["W"]).cuda()
            data["y"] = torch.from_numpy(d["y"]).cuda()
        mod = E.LargeCnlsModel("single_index", d["x0"], data)
        E.solve(mod)
        res.append((mod.sol.copy(), float(mod.obj_value[0]), int(mod.exit_code[0]), int(mod.iterations[0])))
        mod.close()
    for r in res[1:]:
        assert np.array_equal(r[0], res[0][0]) and r[1:] == res[0][1:]      # bit-identical: fixed reduction trees
    assert res[0][2] > 0


def test_large_full_size_properties(E):
    """m = 4M, n = 256, 64 equalities (BASELINE config 4), generated on the device."""
    import torch
    m, n, nb = 1 << 22, 256, 64
    g = torch.Generator(device="cuda").manual_seed(4)
    W = torch.randn(m, n, dtype=torch.float64, device="cuda", generator=g) / np.sqrt(n)
    truth = torch.rand(n, dtype=torch.float64, device="cuda", generator=g) * 2 - 1
    y = torch.tanh(W @ truth) + 0.01 * torch.randn(m, dtype=torch.float64, device="cuda", generator=g)
    x0 = (truth * (1 + 0.05 * (torch.rand(n, dtype=torch.float64, device="cuda", generator=g) * 2 - 1))).cpu().numpy()
    rho = (truth.cpu().numpy()[:4 * nb] ** 2).reshape(nb, 4).sum(axis=1)
    mod = E.LargeCnlsModel("single_index", x0, {"W": W, "y": y, "rho": rho})
    R, _, tms = mod.factor(x0)
    R2, _, _ = mod.factor(x0)
    assert np.array_equal(R, R2)
    # Gram identity with the device's own [J | r] (torch tanh differs from det_tanh by <= 2 ulp: tolerance 1e-11)
    xd = torch.from_numpy(x0).cuda()
    th = torch.tanh(W @ xd)
    s = 1 - th * th
    G = torch.empty(n + 1, n + 1, dtype=torch.float64, device="cuda")
    J = W * s[:, None]
    r = th - y
    G[:n, :n] = J.T @ J
    G[:n, n] = J.T @ r
    G[n, :n] = G[:n, n]
    G[n, n] = r @ r
    del J
    G = G.cpu().numpy()
    assert np.abs(R.T @ R - G).max() <= 1e-11 * np.abs(G).max()
    E.solve(mod)
    assert int(mod.status_code[0]) == 1 and int(mod.exit_code[0]) > 0
    x = mod.sol[0]
    assert np.abs((x[:4 * nb] ** 2).reshape(nb, 4).sum(axis=1) - rho).max() <= 1e-7      # equalities hold (eps_c = sqrt(eps))
    assert np.linalg.norm(x - truth.cpu().numpy()) <= 1e-2 * np.linalg.norm(x)          # noise 0.01, m/n = 16384
    assert abs(float(mod.obj_value[0]) / m - 1e-4) < 2e-6                                # sum r^2 ~ m * sigma^2
    mod.close()


@pytest.mark.parametrize("m,n,nb,seed", [(4096, 512, 128, 31), (3000, 384, 64, 32)])
def test_wide_problem_device_small_stage_vs_oracle(E, m, n, nb, seed):
    """BASELINE config 5 shape at reduced size (inequalities + bounds on every parameter, l = nb + 2n): the compressed
    small stage (blocked QRCP, compact-WY J*Q1, triangular solves: csrc/enl_small.cuh; src/enlsip_functions.jl:219-223,
    700) works on matrices beyond dgeqp3's blocking crossover."""
    from oracle import enlsip_oracle as O, problems as P
    from tests.test_large_host import compare_with_oracle
    d = E.synth.gen_single_index(m, n, nb, seed=seed, ineq=True)
    lo, up = np.full(n, -2.0), np.full(n, 2.0)
    mod = E.LargeCnlsModel("single_index", d["x0"], d, ineq=True, x_low=lo, x_upp=up)
    E.solve(mod, trace_cap=60)
    st = mod.stats()
    assert st["device_mulq"] > 0 and st["device_qrcp"] > 0                   # the device small stage really ran
    r = O.solve(P.single_index(d["W"], d["y"], d["rho"], d["x0"], ineq=True, bounds=(-2.0, 2.0)), wallclock=False)
    out = dict(x=mod.sol[0], f=mod.obj_value, exit_code=mod.exit_code, status=mod.status_code, iters=mod.iterations,
               nact=mod.nb_active, active=mod.active[0], trace=mod.trace[0])
    compare_with_oracle(out, r, n)
    mod.close()


def test_c5_full_size_vs_golden(E, golden_dir):
    """BASELINE config 5 at the named size: n = 4096, m = 16384, 1024 nonlinear inequalities + 8192 bounds, against
    the committed oracle solve of the same problem (tests/golden/c5_full_oracle.npz, made by make_c5_fixture.py:
    606 s on 8 cores): identical exit code, iteration count and per-iteration (t, rankA, rankJ2, dimA, dimJ2, method),
    identical final working set, objective to 1e-10; plus size-independent properties."""
    import os
    m, n, nb = 16384, 4096, 1024
    gold = np.load(os.path.join(golden_dir, "c5_full_oracle.npz"))
    d = E.synth.gen_single_index(m, n, nb, seed=5, ineq=True)
    lo, up = np.full(n, -2.0), np.full(n, 2.0)
    mod = E.LargeCnlsModel("single_index", d["x0"], d, ineq=True, x_low=lo, x_upp=up)
    E.solve(mod, trace_cap=40)
    st = mod.stats()
    assert st["device_qrcp"] > 0 and st["device_mulq"] > 0
    assert int(mod.exit_code[0]) == int(gold["exit_code"]) and int(mod.iterations[0]) == int(gold["iterations"])
    assert int(mod.status_code[0]) == 1
    k = gold["trace"].shape[0]
    assert np.array_equal(mod.trace[0][:k, 1:7].astype(np.int32), gold["trace"])
    assert abs(float(mod.obj_value[0]) - float(gold["f"])) <= 1e-10 * float(gold["f"])
    x = mod.sol[0]
    assert np.linalg.norm(x - gold["x"]) <= 1e-8 * np.linalg.norm(gold["x"])
    act = np.sort(mod.active[0][: int(mod.nb_active[0])])
    assert np.array_equal(act, gold["active"])
    g = d["rho"] - (x[:4 * nb] ** 2).reshape(nb, 4).sum(axis=1)
    assert g.min() >= -1e-7 and np.all(np.abs(x) <= 2.0 + 1e-12)
    assert np.all((act - 1) % 2 == 0)                          # only the tightened blocks (even k) can be active
    f1 = float(mod.obj_value[0])
    E.solve(mod)
    assert float(mod.obj_value[0]) == f1                       # deterministic
    mod.close()


def test_c4_1M_rows_vs_golden(E, golden_dir):
    """BASELINE config 4 at a quarter of the named row count (m = 2^20, n = 256, 64 equalities) against the
    committed oracle solve on the full m x n Jacobian (tests/golden/c4_1M_oracle.npz, make_c4_fixture.py: 149 s
    on 8 cores): identical exit code / iterations / per-iteration trace, objective 1e-10, every iterate 1e-10
    except the last (flat merit function)."""
    import os
    gold = np.load(os.path.join(golden_dir, "c4_1M_oracle.npz"))
    m, n, nb = int(gold["m"]), 256, 64
    d = E.synth.gen_single_index(m, n, nb, seed=4)
    mod = E.LargeCnlsModel("single_index", d["x0"], d)
    E.solve(mod, trace_cap=40)
    assert int(mod.exit_code[0]) == int(gold["exit_code"]) and int(mod.iterations[0]) == int(gold["iterations"])
    k = gold["trace"].shape[0]
    assert np.array_equal(mod.trace[0][:k, 1:7].astype(np.int32), gold["trace"])
    assert abs(float(mod.obj_value[0]) - float(gold["f"])) <= 1e-10 * float(gold["f"])
    for i in range(k - 1):
        xi = gold["x_iter"][i]
        assert np.linalg.norm(mod.trace[0][i, 16:16 + n] - xi) <= 1e-10 * np.linalg.norm(xi), i
    assert np.linalg.norm(mod.sol[0] - gold["x"]) <= 1e-8 * np.linalg.norm(gold["x"])
    assert np.array_equal(np.sort(mod.active[0][: int(mod.nb_active[0])]), gold["active"])
    mod.close()


def test_c4_named_size_vs_golden(E, golden_dir):
    """BASELINE config 4 at the NAMED size and on the data bench.py times (m = 2^22, n = 256, 64 equalities,
    synth.gen_single_index seed 4) against the committed oracle solve on the full 4M x 256 Jacobian
    (tests/golden/c4_2p22_oracle.npz, make_c4_fixture.py 22: 444 s on 8 cores): identical exit code / iteration count /
    per-iteration trace / working set, objective 1e-10, iterates 1e-10 before the last step (flat merit function)."""
    import os
    gold = np.load(os.path.join(golden_dir, "c4_2p22_oracle.npz"))
    m, n, nb = int(gold["m"]), 256, 64
    assert m == 1 << 22
    d = E.synth.gen_single_index(m, n, nb, seed=4)
    mod = E.LargeCnlsModel("single_index", d["x0"], d)
    E.solve(mod, trace_cap=40)
    assert int(mod.exit_code[0]) == int(gold["exit_code"]) and int(mod.iterations[0]) == int(gold["iterations"])
    k = gold["trace"].shape[0]
    assert np.array_equal(mod.trace[0][:k, 1:7].astype(np.int32), gold["trace"])
    assert abs(float(mod.obj_value[0]) - float(gold["f"])) <= 1e-10 * float(gold["f"])
    for i in range(k - 1):
        xi = gold["x_iter"][i]
        assert np.linalg.norm(mod.trace[0][i, 16:16 + n] - xi) <= 1e-10 * np.linalg.norm(xi), i
    assert np.linalg.norm(mod.sol[0] - gold["x"]) <= 1e-8 * np.linalg.norm(gold["x"])
    assert np.array_equal(np.sort(mod.active[0][: int(mod.nb_active[0])]), gold["active"])
    st = mod.stats()
    assert st["device_qrcp"] > 0 and st["device_mulq"] > 0
    mod.close()


@pytest.mark.parametrize("kw", [dict(max_iter=2), dict(scaling=True), dict(time_limit=-1.0)])
def test_large_options_vs_oracle(E, kw):
    """solve!(model; max_iter, scaling, time_limit) in the large regime (solver.jl:62-63): -2 / -11 statuses as in the
    reference (test/problems/chained_rosenbrock.jl:71-73 pins :time_limit_exceeded for time_limit = -1)."""
    from oracle import enlsip_oracle as O, problems as P
    from tests.test_large_host import compare_with_oracle
    d = E.synth.gen_single_index(1500, 32, 8, seed=4)
    mod = E.LargeCnlsModel("single_index", d["x0"], d)
    E.solve(mod, trace_cap=60, **kw)
    r = O.solve(P.single_index(d["W"], d["y"], d["rho"], d["x0"]), wallclock="time_limit" in kw, **kw)
    assert int(mod.exit_code[0]) == r.exit_code and int(mod.status_code[0]) == r.status
    assert int(mod.iterations[0]) == r.iterations
    if "time_limit" in kw:
        assert str(E.status(mod)[0]).lstrip(":") == "time_limit_exceeded"
        assert np.array_equal(mod.sol[0], d["x0"])          # x_opt stays x0
    else:
        out = dict(x=mod.sol[0], f=mod.obj_value, exit_code=mod.exit_code, status=mod.status_code, iters=mod.iterations,
                   nact=mod.nb_active, active=mod.active[0], trace=mod.trace[0])
        compare_with_oracle(out, r, 32)
    mod.close()


def test_rank_deficient_jacobian_vs_oracle(E):
    """pseudo_rank (EF:17-31) of the QRCP of J2 on the compressed problem: a Jacobian with two identical columns."""
    from oracle import enlsip_oracle as O, problems as P
    from tests.test_large_host import rank_deficient_problem, check_rank_deficient
    d = rank_deficient_problem()
    mod = E.LargeCnlsModel("single_index", d["x0"], d)
    E.solve(mod, trace_cap=60)
    r = O.solve(P.single_index(d["W"], d["y"], d["rho"], d["x0"]), wallclock=False)
    out = dict(x=mod.sol[0], f=mod.obj_value, exit_code=mod.exit_code, iters=mod.iterations, trace=mod.trace[0])
    check_rank_deficient(out, r)
    mod.close()


def test_large_solve_seed_sweep_vs_oracle(E):
    """Twelve more random instances (sizes, seeds, equality / inequality / bound mixes) against the live oracle: every
    discrete output identical, objective 1e-10, iterate before the last step 1e-10."""
    from oracle import enlsip_oracle as O, problems as P
    from tests.test_large_host import compare_with_oracle
    rng = np.random.default_rng(2024)
    done = 0
    for k in range(12):
        n = int(rng.choice([32, 64, 96]))
        m = int(rng.integers(1200, 5000))
        nb = int(rng.integers(1, n // 4 + 1))
        ineq = bool(k % 2)
        bounds = (-2.0, 2.0) if (ineq and k % 4 == 1) else None
        d = E.synth.gen_single_index(m, n, nb, seed=100 + k, ineq=ineq)
        lo = None if bounds is None else np.full(n, bounds[0])
        up = None if bounds is None else np.full(n, bounds[1])
        mod = E.LargeCnlsModel("single_index", d["x0"], d, ineq=ineq, x_low=lo, x_upp=up)
        E.solve(mod, trace_cap=100)
        r = O.solve(P.single_index(d["W"], d["y"], d["rho"], d["x0"], ineq=ineq, bounds=bounds), wallclock=False)
        out = dict(x=mod.sol[0], f=mod.obj_value, exit_code=mod.exit_code, status=mod.status_code, iters=mod.iterations,
                   nact=mod.nb_active, active=mod.active[0], trace=mod.trace[0])
        compare_with_oracle(out, r, n)
        mod.close()
        done += 1
    assert done == 12


@pytest.mark.parametrize("n,jac", [(1000, "analytic"), (1000, "forward_diff"), (334, "analytic"), (400, "forward_diff")])
def test_chained_rosenbrock_reference_size_vs_oracle(E, n, jac):
    """The reference's own large test problem, test/problems/chained_rosenbrock.jl:3-53, at ITS size: n = 1000,
    m = 1998 residuals, 998 nonlinear equalities (n + m >= 1000: Newton off, EF:2658 -> the large regime), as a general
    row family (csrc/enl_large_family.h) -- no multiple-of-32 restriction, [J | r] factored by the plain-Householder
    mode of the device QR, the whole small stage (998 active constraints of 1000 parameters) on the device.  Checked
    against the oracle: status, exit code, iteration count, active set, per-iteration trace identical, f to 1e-10."""
    from oracle import enlsip_oracle as O, problems as P
    from tests.test_large_host import compare_with_oracle
    pb = P.chained_rosenbrock(n, fd=(jac == "forward_diff"))
    mod = E.LargeCnlsModel("chained_rosenbrock", pb.x0, jacobian=jac)
    assert (mod.nb_parameters, mod.nb_residuals, mod.nb_eqcons, mod.nb_constraints) == (n, 2 * (n - 1), n - 2, n - 2)
    E.solve(mod, trace_cap=60)
    r = O.solve(pb, wallclock=False)
    assert int(mod.status_code[0]) == r.status == 1
    assert int(mod.exit_code[0]) == r.exit_code and int(mod.iterations[0]) == r.iterations
    assert [int(v) for v in mod.active[0] if v > 0] == r.active
    if jac == "analytic":
        out = dict(x=mod.sol[0], f=mod.obj_value, exit_code=mod.exit_code, status=mod.status_code, iters=mod.iterations,
                   nact=mod.nb_active, active=mod.active[0], trace=mod.trace[0])
        compare_with_oracle(out, r, n)
    else:      # forward differences: noise floor eps |r| / delta ~ 1e-8 in J
        assert abs(float(mod.obj_value[0]) - r.f) <= 1e-8 * r.f
        assert np.linalg.norm(mod.sol[0] - r.x) <= 1e-6 * np.linalg.norm(r.x)
    st = mod.stats()
    assert st["device_qrcp"] > 0 and st["launches"] > 0
    mod.close()


def test_tsqr_tma_trailing_kernel_matches():
    """The TMA-fed trailing updates (ENLSIP_TRAIL=3: producer warp, cp.async.bulk + mbarrier; ENLSIP_TRAIL=4: one
    cp.async.bulk.tensor.2d box per 16 x 32 tile through a tensor map, 8 or 16 math warps) against the default cp.async
    kernel: mode 3 returns its bits, mode 4 agrees to rounding.  The switch is read once per process, so the variant runs in a child process."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import sys, numpy as np; sys.path.insert(0, %r); import enlsip_jl_b200 as E\n"
            "d = E.synth.gen_single_index(70000, 128, 32, seed=2)\n"
            "mod = E.LargeCnlsModel('single_index', d['x0'], d, m_global=70000)\n"
            "R, _, _ = mod.factor(d['x0']); sys.stdout.buffer.write(R.tobytes())\n" % root)
    outs = []
    for mode, nw in (("2", "8"), ("3", "8"), ("4", "8"), ("4", "16")):
        env = dict(os.environ, ENLSIP_TRAIL=mode, ENLSIP_TMAP_NW=nw)
        outs.append(subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, check=True).stdout)
    a, b, c8, c16 = (np.frombuffer(o, dtype=np.float64) for o in outs)
    assert a.size == 129 * 129 and np.array_equal(a.view(np.uint64), b.view(np.uint64))
    # ENLSIP_TRAIL=4 (tensor-map TMA, 128-byte swizzle): the k order inside a DMMA chain follows the swizzled layout,
    # so the sums are the same up to rounding, not to the bit
    for c in (c8, c16):
        assert c.size == a.size and np.abs(c - a).max() <= 1e-12 * np.abs(a).max()


@pytest.mark.parametrize("m,n,nb,seed", [(3000, 32, 8, 41), (5000, 64, 16, 42)])
def test_single_index_forward_difference_jacobian_vs_oracle(E, m, n, nb, seed):
    """The FD-Jacobian variant of the tall family (SURVEY.md 8d, C4 inputs; jac_forward_diff, cnls_model.jl:65-82): the
    residual Jacobian by forward differences inside the build kernel (one det_tanh per entry, no second pass over W)
    against the oracle with the same differencing: same status / iteration count / working set, f and x to the FD noise
    floor.  The analytic solve of the same problem lands on the same point to 1e-6."""
    from oracle import enlsip_oracle as O, problems as P
    d = E.synth.gen_single_index(m, n, nb, seed=seed)
    mod = E.LargeCnlsModel("single_index", d["x0"], d, jacobian="forward_diff")
    E.solve(mod, trace_cap=60)
    r = O.solve(P.single_index(d["W"], d["y"], d["rho"], d["x0"], fd_res=True), wallclock=False)
    assert int(mod.status_code[0]) == r.status == 1
    assert abs(int(mod.iterations[0]) - r.iterations) <= 1
    assert sorted(int(v) for v in mod.active[0] if v > 0) == sorted(r.active)
    assert abs(float(mod.obj_value[0]) - r.f) <= 1e-8 * r.f
    assert np.linalg.norm(mod.sol[0] - r.x) <= 1e-6 * np.linalg.norm(r.x)
    ma = E.LargeCnlsModel("single_index", d["x0"], d)
    E.solve(ma)
    assert np.linalg.norm(mod.sol[0] - ma.sol[0]) <= 1e-6 * np.linalg.norm(ma.sol[0])
    mod.close(); ma.close()
