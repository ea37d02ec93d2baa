"""Row-sharded large regime on several GPUs of one node (one process per GPU, NCCL): the cross-rank level of the TSQR tree
(csrc/enl_tsqr.cuh, TsqrDist) and the stacked second stage against the single-GPU solve.  Needs at least two visible
GPUs; skipped otherwise (the single-GPU box of the round-end test run).  tools/dist_check.py does the work:
R up to row signs, iteration count, objective and x of the sharded solve against rank 0 solving the whole problem alone."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("mode", ["tree", "stack"])
def test_row_sharded_factorisation_matches_single_gpu(mode):
    n = _gpus()
    if n < 2:
        pytest.skip("needs two GPUs")
    world = 2 if n < 4 else 4
    env = dict(os.environ, ENLSIP_TSQR_DIST=mode)
    port = 29600 + (os.getpid() % 200) + (0 if mode == "tree" else 1)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tools", "dist_check.py"), str(3 * 65536 + 4096)]
    p = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    assert ("mode %s" % mode) in p.stdout, p.stdout[-2000:]
