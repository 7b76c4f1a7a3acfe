"""The REAL multi-rank path (one process per GPU, NCCL rendezvous, CUDA-IPC peer memory, epoch flags) under pytest:
launches tools/dist_check.py with torchrun on 2 ranks, once with the fused peer-memory exchange and once with the NCCL
send/recv transport (CMC_P2P=0).  Every rank compares its planes, the residual, the checksums and the GetLayer output with
the same case solved as one slab on its own GPU.  Needs 2 visible GPUs (skipped otherwise: ranks that wait on each other's
flags must not share one device, B200_PROFILING.md)."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _gpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("p2p", ["1", "0"])
def test_two_ranks_against_single_gpu(p2p):
    if _gpus() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    env = dict(os.environ, CMC_P2P=p2p)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533" if p2p == "1" else "29534", str(ROOT / "tools" / "dist_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env, cwd=str(ROOT))
    assert r.returncode == 0 and "DIST CHECK PASSED" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
    want = "fused-stores-peer-memory" if p2p == "1" else "nccl"
    assert f"exchange {want}" in r.stdout, r.stdout[-2000:]
