"""Multi-GPU path on CPU: world_size-2 gloo run of the sharding logic bench.py uses (no data-path collective).

Each rank generates its contiguous shard of the C3 stream, solves it with the host build of the solver core and
the ranks only exchange a max-reduced time and gathered result digests -- the same plumbing as the NCCL run.
"""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, B, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import enlsip_jl_b200 as E
    from tests.test_hostport import run
    y, S, x0, _ = E.synth.gen_gauss_peaks_batch(B, start=rank * B)
    out = run(1, x0, y, S, E.synth.GP_LOW, E.synth.GP_UPP, 1, trace_cap=1, nthreads=1)
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)                     # the only collective: max over ranks of the time
    gathered = [torch.zeros(B, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(gathered, torch.from_numpy(out["f"].copy()))
    if rank == 0:
        q.put((float(t.item()), torch.cat(gathered).numpy()))
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_rank():
    import enlsip_jl_b200 as E
    from tests.test_hostport import run
    B, world = 12, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, B, q)) for r in range(world)]
    for p in procs:
        p.start()
    tmax, f_sharded = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert tmax == 2.0
    y, S, x0, _ = E.synth.gen_gauss_peaks_batch(world * B)
    full = run(1, x0, y, S, E.synth.GP_LOW, E.synth.GP_UPP, 1, trace_cap=1, nthreads=1)
    assert np.array_equal(full["f"], f_sharded)                  # shards reproduce the global stream bit for bit


def test_shard_stream_is_position_independent():
    import enlsip_jl_b200 as E
    y, S, x0, _ = E.synth.gen_gauss_peaks_batch(10)
    y2, S2, x02, _ = E.synth.gen_gauss_peaks_batch(4, start=5)
    assert np.array_equal(y[5:9], y2) and np.array_equal(S[5:9], S2) and np.array_equal(x0[5:9], x02)
    a = E.synth.gen_hs65_batch(8)
    assert np.array_equal(a[3:8], E.synth.gen_hs65_batch(5, start=3))
