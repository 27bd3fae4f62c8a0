"""Multi-process (gloo, CPU) coverage of the N>1 host path: communicator wrapper, Split into the
pencil sub-communicators, per-rank geometry."""
import os
import socket
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("nproc", [2, 4])
def test_torchcomm_gloo(nproc):
    import torch
    if torch.cuda.is_available():
        pytest.skip("CPU-only test (the worker asserts that transforms refuse to run without a GPU)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc),
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(HERE, "dist_worker.py")]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=300, env=env)
    text = out.stdout.decode("utf-8", "replace")
    assert out.returncode == 0, text[-4000:]
    assert text.count("WORKER_OK") == nproc, text[-4000:]
