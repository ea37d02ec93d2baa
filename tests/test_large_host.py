"""Large-Jacobian regime on the CPU: the host ENLSIP driver over the compressed [R_J | Q'r] problem
(enlsip.jl_b200/csrc/enl_large_host.h, compiled with the CPU backend oracle/hostport/largeport.cpp)
against the oracle, which works on the full m x n Jacobian exactly like the reference
(src/enlsip_functions.jl:206-234, 2638-2880).

Also the N > 1 path: a world-size-2 gloo run of the row-sharded TSQR (local R factors all-gathered and
re-factored on every rank, linesearch sums all-reduced) must reproduce the single-process
solve.  The GPU build replaces the callbacks by ncclAllGather / ncclAllReduce (tests/test_gpu_large.py).
"""
import ctypes
import math
import os
import socket
import sys

import numpy as np
import pytest

import __graft_entry__ as ge

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
AG = ctypes.CFUNCTYPE(None, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double), ctypes.c_longlong)
AR = ctypes.CFUNCTYPE(None, ctypes.POINTER(ctypes.c_double), ctypes.c_int)


def run_large(d, ineq=False, bounds=None, rows=None, m_global=None, world=1, ag=None, ar=None, trace_cap=60, nthreads=2,
              max_iter=100, scaling=False):
    """largeport_solve_sharded on the dict produced by synth.gen_single_index."""
    lib = ctypes.CDLL(ge.build_largeport())
    W = np.ascontiguousarray(d["W"]); y = np.ascontiguousarray(d["y"]); rho = np.ascontiguousarray(d["rho"])
    x0 = np.ascontiguousarray(d["x0"])
    rows_, n = W.shape
    m = rows_ if m_global is None else m_global
    lo = np.full(n, -np.inf) if bounds is None else np.full(n, float(bounds[0]))
    up = np.full(n, np.inf) if bounds is None else np.full(n, float(bounds[1]))
    l = rho.size + (0 if bounds is None else 2 * n)
    se = math.sqrt(np.finfo(float).eps)
    out = dict(x=np.zeros(n), f=np.zeros(1), exit_code=np.zeros(1, np.int32), status=np.zeros(1, np.int32),
               iters=np.zeros(1, np.int32), nact=np.zeros(1, np.int32), active=np.zeros(max(l, 1), np.int32),
               trace=np.zeros((trace_cap, 16 + n)))
    vp, ci, ll, cd = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_double
    lib.largeport_solve_sharded.argtypes = [ci, ll, ll, ci, AG, AR, ci, ci] + [vp] * 6 + [ci, ci, cd, cd, cd] + [vp] * 8 + [ci, ci]
    p = lambda a: a.ctypes.data_as(vp)
    rc = lib.largeport_solve_sharded(n, m, rows_, world, ag if ag else AG(), ar if ar else AR(), rho.size, 1 if ineq else 0,
                                     p(W), p(y), p(rho), p(lo), p(up), p(x0), max_iter, 1 if scaling else 0, se, se, se, p(out["x"]),
                                     p(out["f"]), p(out["exit_code"]), p(out["status"]), p(out["iters"]), p(out["nact"]),
                                     p(out["active"]), p(out["trace"]), trace_cap, nthreads)
    assert rc == 0
    return out


def compare_with_oracle(out, r, n):
    """discrete trace identical; objective 1e-10; iterate before the last step 1e-10 (see tests/parity.py)."""
    assert int(out["exit_code"][0]) == r.exit_code and int(out["iters"][0]) == r.iterations
    assert int(out["status"][0]) == r.status
    for k, t in enumerate(r.trace):
        e = out["trace"][k]
        assert (int(e[1]), int(e[2]), int(e[3]), int(e[4]), int(e[5]), int(e[6])) == \
               (t.t, t.rankA, t.rankJ2, t.dimA, t.dimJ2, t.code), k
    assert abs(out["f"][0] - r.f) <= 1e-10 * max(1.0, abs(r.f))
    assert sorted(out["active"][: int(out["nact"][0])].tolist()) == sorted(r.active)
    if len(r.trace) >= 2:
        k = len(r.trace) - 2
        assert np.linalg.norm(out["trace"][k, 16:16 + n] - r.trace[k].x_new) <= 1e-10 * np.linalg.norm(r.trace[k].x_new)
    # the last step of a converged solve is decided by a merit function flat to rounding
    assert np.linalg.norm(out["x"] - r.x) <= 1e-8 * np.linalg.norm(r.x)


@pytest.mark.parametrize("m,n,nb,seed,ineq,bounds", [(1500, 32, 8, 4, False, None), (2048, 64, 16, 7, False, None),
                                                      (1200, 32, 8, 5, True, (-2.0, 2.0)), (1100, 32, 3, 9, True, None)])
def test_large_host_driver_vs_oracle(m, n, nb, seed, ineq, bounds):
    import enlsip_jl_b200 as E
    from oracle import enlsip_oracle as O, problems as P
    d = E.synth.gen_single_index(m, n, nb, seed=seed, ineq=ineq)
    out = run_large(d, ineq=ineq, bounds=bounds)
    r = O.solve(P.single_index(d["W"], d["y"], d["rho"], d["x0"], ineq=ineq, bounds=bounds), wallclock=False)
    compare_with_oracle(out, r, n)


@pytest.mark.parametrize("kw", [dict(max_iter=2), dict(scaling=True)])
def test_large_host_driver_options_vs_oracle(kw):
    """solve!(model; max_iter, scaling) (solver.jl:62-63): iteration cap -> exit code -2 (:maximum_iterations_exceeded),
    row scaling of the active constraints (structures.jl:168-175) -> same trace machinery."""
    import enlsip_jl_b200 as E
    from oracle import enlsip_oracle as O, problems as P
    d = E.synth.gen_single_index(1500, 32, 8, seed=4)
    out = run_large(d, **kw)
    r = O.solve(P.single_index(d["W"], d["y"], d["rho"], d["x0"]), wallclock=False, **kw)
    compare_with_oracle(out, r, 32)
    if "max_iter" in kw:
        assert int(out["exit_code"][0]) == -2 and int(out["status"][0]) == -2


def rank_deficient_problem():
    """Two identical columns of W (outside the constrained blocks): J and J2 have rank n - 1 at every x."""
    import enlsip_jl_b200 as E
    d = E.synth.gen_single_index(1500, 32, 4, seed=13)
    d["W"][:, 25] = d["W"][:, 21]
    return d


def check_rank_deficient(out, r):
    """rankJ2 = n - t - 1 is found identically; the objective and every discrete output agree.  x itself is not
    unique: QRCP may pivot on either of the two identical columns (an exact tie of column norms, broken by rounding),
    which moves x along e_21 - e_25 only; everything identifiable (the other coordinates, x_21 + x_25) agrees."""
    n = 32
    assert int(out["exit_code"][0]) == r.exit_code and int(out["iters"][0]) == r.iterations
    for k, t in enumerate(r.trace):
        e = out["trace"][k]
        assert (int(e[1]), int(e[2]), int(e[3]), int(e[4]), int(e[5]), int(e[6])) == (t.t, t.rankA, t.rankJ2, t.dimA, t.dimJ2, t.code)
        assert t.rankJ2 == n - t.t - 1
    assert abs(out["f"][0] - r.f) <= 1e-10 * r.f
    dx = np.asarray(out["x"]).reshape(-1) - r.x
    assert np.abs(np.delete(dx, [21, 25])).max() <= 1e-7
    assert abs(dx[21] + dx[25]) <= 1e-7


def test_rank_deficient_jacobian_vs_oracle():
    from oracle import enlsip_oracle as O, problems as P
    d = rank_deficient_problem()
    out = run_large(d)
    r = O.solve(P.single_index(d["W"], d["y"], d["rho"], d["x0"]), wallclock=False)
    check_rank_deficient(out, r)


def test_single_index_shards_are_position_independent():
    import enlsip_jl_b200 as E
    full = E.synth.gen_single_index(200000, 8, 2, seed=4)
    part = E.synth.gen_single_index(200000, 8, 2, seed=4, start=70000, rows=90000)
    assert np.array_equal(full["W"][70000:160000], part["W"]) and np.array_equal(full["y"][70000:160000], part["y"])
    assert np.array_equal(full["x0"], part["x0"]) and np.array_equal(full["rho"], part["rho"])


# ------------------------------------------------------------------------------------------------------------
# world-size-2 gloo run of the row-sharded solve
# ------------------------------------------------------------------------------------------------------------
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, m, n, nb, seed, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import enlsip_jl_b200 as E
    from tests.test_large_host import run_large, AG, AR
    rows = m // world
    d = E.synth.gen_single_index(m, n, nb, seed=seed, start=rank * rows, rows=rows if rank < world - 1 else m - rank * rows)
    ncalls = [0, 0]

    def allgather(send, recv, count):
        ncalls[0] += 1
        s = torch.from_numpy(np.ctypeslib.as_array(send, shape=(count,)).copy())
        parts = [torch.zeros(count, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(parts, s)
        np.ctypeslib.as_array(recv, shape=(world * count,))[:] = torch.cat(parts).numpy()

    def allreduce(buf, count):
        ncalls[1] += 1
        a = np.ctypeslib.as_array(buf, shape=(count,))
        t = torch.from_numpy(a.copy())
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        a[:] = t.numpy()

    out = run_large(d, m_global=m, world=world, ag=AG(allgather), ar=AR(allreduce), nthreads=1)
    digest = torch.from_numpy(np.concatenate([out["x"], out["f"], out["exit_code"].astype(float), out["iters"].astype(float)]))
    allv = [torch.zeros_like(digest) for _ in range(world)]
    dist.all_gather(allv, digest)
    if rank == 0:
        q.put((out, [v.numpy() for v in allv], ncalls))
    dist.destroy_process_group()


def test_row_sharded_solve_two_ranks_gloo():
    import torch.multiprocessing as mp
    import enlsip_jl_b200 as E
    m, n, nb, seed, world = 3000, 32, 8, 11, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, m, n, nb, seed, q)) for r in range(world)]
    for p in procs:
        p.start()
    out, digests, ncalls = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # replicas are bit-identical (every rank re-factors the same stacked R and sees the same reduced sums)
    assert np.array_equal(digests[0], digests[1])
    assert ncalls[0] >= 2 and ncalls[1] >= 4      # one all-gather per new point, all-reduces in the linesearch
    single = run_large(E.synth.gen_single_index(m, n, nb, seed=seed), nthreads=1)
    assert int(out["exit_code"][0]) == int(single["exit_code"][0]) and int(out["iters"][0]) == int(single["iters"][0])
    k = int(single["iters"][0])
    assert np.array_equal(out["trace"][:k, 1:7], single["trace"][:k, 1:7])      # t, ranks, dims, method per iteration
    assert abs(out["f"][0] - single["f"][0]) <= 1e-12 * abs(single["f"][0])
    if k >= 2:
        assert np.linalg.norm(out["trace"][k - 2, 16:] - single["trace"][k - 2, 16:]) <= 1e-11 * np.linalg.norm(single["trace"][k - 2, 16:])
    assert np.linalg.norm(out["x"] - single["x"]) <= 1e-8 * np.linalg.norm(single["x"])
