"""Row-sharded TSQR check (run under torchrun, one rank per GPU):
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py [rows]
Every rank factors its shard of a C4-type problem; rank 0 also factors and solves the whole problem alone and compares
R (up to the signs of its rows), the iteration count, the objective and x."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import enlsip_jl_b200 as E                                   # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl")
m = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 18
n, nb = 256, 64
rows = m // world
row0 = rank * rows
if rank == world - 1:
    rows = m - row0
d = E.synth.gen_single_index(m, n, nb, seed=4, start=row0, rows=rows)
mod = E.LargeCnlsModel("single_index", d["x0"], {"W": d["W"], "y": d["y"], "rho": d["rho"]}, m_global=m, device=local)
mod.join(rank, world)
R, _, t_ms = mod.factor(d["x0"])
for _ in range(3):
    _, _, t_ms = mod.factor(d["x0"], want_R=False)
E.solve(mod)
its, f, x = int(mod.iterations[0]), float(mod.obj_value[0]), mod.sol[0].copy()
st = mod.stats()
mod.close()
dist.barrier()
if rank == 0:
    dall = E.synth.gen_single_index(m, n, nb, seed=4)
    one = E.LargeCnlsModel("single_index", dall["x0"], dall, device=local)
    R1, _, t1 = one.factor(dall["x0"])
    E.solve(one)
    sg = np.sign(np.diag(R)) * np.sign(np.diag(R1))
    dR = np.abs(R * sg[:, None] - R1).max() / np.abs(R1).max()
    print("ranks %d rows %d mode %s: tsqr %.3f ms per factorisation (1 GPU: %.3f); max |R - R_1gpu| / max|R| = %.2e; iterations %d vs %d; "
          "f rel diff %.2e; x rel diff %.2e" % (world, m, os.environ.get("ENLSIP_TSQR_DIST", "tree"), t_ms, t1, dR, its,
                                                  int(one.iterations[0]), abs(f - float(one.obj_value[0])) / abs(f),
                                                  np.linalg.norm(x - one.sol[0]) / np.linalg.norm(x)), flush=True)
    assert dR <= 1e-10 and its == int(one.iterations[0])
    assert abs(f - float(one.obj_value[0])) <= 1e-10 * abs(f)
    one.close()
dist.barrier()
dist.destroy_process_group()
