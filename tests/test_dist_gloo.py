"""World-size-2 gloo run of the host transport used by the multi-GPU path (handle exchange, id broadcast,
slab scatter/gather, box <-> slab all-to-all) -- CPU only."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


import pytest


@pytest.mark.parametrize("world", [2, 4])
def test_host_transport(world):
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", CUDA_VISIBLE_DEVICES="")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr",
           "127.0.0.1", "--master-port", str(29531 + world), os.path.join(ROOT, "tests", "gloo_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert res.stdout.count("GLOO_WORKER_OK") == world


def test_invalid_transport_combination_is_rejected():
    from petibm_b200.dist import Comm

    with pytest.raises(ValueError):
        Comm(0, 2, 0, reduce="p2p", halo="memcpy")
    Comm(0, 1, 0, reduce="p2p", halo="memcpy")   # single rank: nothing to order
