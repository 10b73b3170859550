"""N>1 path: rows sharded over ranks, K-length partial sums allreduced."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

REPO = Path(__file__).resolve().parents[1]


def _run(backend, nproc, port, extra_env=None):
    # the oracle's OpenMP regions ask for omp_get_num_procs() threads each (modandbase.cpp:464): cap them, or nproc ranks
    # oversubscribe the host nproc-fold and spin in each other's barriers
    per_rank = str(max(2, (os.cpu_count() or 2) // nproc))
    env = dict(os.environ, OMP_NUM_THREADS="2", OMP_THREAD_LIMIT=per_rank, OMP_WAIT_POLICY="passive", **(extra_env or {}))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(REPO / "tests" / "mgpu_worker.py"), backend]
    return subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=600)


def test_row_sharding_world2_gloo(oracle):
    """CPU, gloo, world_size 2: the oracle on row blocks + allreduce reproduces the full-data results."""
    res = _run("gloo", 2, 29541)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    assert "mgpu_worker[gloo] world=2 OK" in res.stdout


@pytest.mark.gpu
def test_row_sharding_nccl(gpu):
    """GPU: one process per B200, ncclAllReduce inside the library; needs >= 2 GPUs on the box."""
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("single-GPU box")
    res = _run("nccl", n, 29542)  # every visible GPU (8 on the scaling box)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    # the same on the terms-specialised kernels: Phi^T reduces over CTAs AND ranks in its own tail, optcg keeps its
    # vectors in HBM (the sum of squared residuals rides on the fused allreduce)
    res = _run("nccl", n, 29543, {"OB_SPEC": "1"})
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
