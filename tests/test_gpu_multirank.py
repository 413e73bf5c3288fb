"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): vectors and CSR rows sharded
over the ranks, halo exchange + rank-ordered scalar-pack combination, bit-exact against the
oracle.  The CPU (gloo) tests of the host-side sharding logic are in test_sharding_cpu.py."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.parametrize("world,no_peer", [(2, 0), (2, 1), (4, 0), (8, 0)])
def test_sharded_runs_bit_exact(world, no_peer):
    """no_peer = 0: halo pushes over peer memory (CUDA IPC) fused into the kernels;
    no_peer = 1: the NCCL point-to-point exchange path.  Both must match the oracle bit for bit."""
    if _ngpus() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29511 + world + 10 * no_peer),
           os.path.join(ROOT, "tests", "multirank_worker.py")]
    env = dict(os.environ, CGO_NO_PEER=str(no_peer))
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert p.returncode == 0 and "MULTIRANK_OK" in p.stdout, p.stdout[-3000:] + p.stderr[-3000:]
    assert f"PEER_MEMORY={1 - no_peer}" in p.stdout, p.stdout[-2000:]
