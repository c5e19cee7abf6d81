"""GPU, needs >= 2 devices: the sharded whole-genome path under torch.distributed.run
(NCCL), checked on hardware -- identical joined q table on every rank, equal to the
single-process join, same peaks (tests/multirank_worker.py).  Skipped on a one-GPU box;
the CPU side of the exchange is covered by tests/test_multigraph_gloo.py."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_sharded_whole_genome_join_on_all_gpus():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("one GPU: the multi-rank join needs at least two")
    world = 8 if n >= 8 else (4 if n >= 4 else 2)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", "29631",
           os.path.join(ROOT, "tests", "multirank_worker.py"), "64"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert "MULTIRANK OK world=%d" % world in out.stdout
