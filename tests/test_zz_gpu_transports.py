"""Multi-GPU parity of the slab transports that are not the default: NCCL send/recv and the fused
peer-store transport (the y / x FFT pass writes straight into the peers' receive buffers over NVLink).
Runs last (file name) and only where the box has >= 2 GPUs."""
import os
import subprocess
import sys

import pytest

from test_gpu_multi import HERE, _free_port

pytestmark = pytest.mark.gpu


# Every mode below passed on 4 GPUs in round 2 (profiles/r02_multi_4/parity_*.log; the same transports x pipelines
# also carry forward parity at 2 and 8 GPUs in profiles/r02_multi_2, r02_multi_8).  Each case is its own torchrun with
# a timeout: a protocol mistake between ranks would show up as a hang that only the timeout ends.
@pytest.mark.parametrize("transport,pipeline", [
    ("nccl", "x"), ("p2p", "x"), ("store", "x"), ("nccl", "kz"), ("p2p", "kz"), ("store", "kz"),
    ("nccl", "pencil-chunks"), ("p2p", "pencil-chunks")])
@pytest.mark.parametrize("nproc", [2, 4, 8])
def test_slab_transport_parity(nproc, transport, pipeline):
    import torch
    if torch.cuda.device_count() < nproc:
        pytest.skip("needs %d GPUs" % nproc)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc),
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(HERE, "gpu_dist_worker.py"), "--transport", transport, pipeline]
    out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=420)
    text = out.stdout.decode("utf-8", "replace")
    assert out.returncode == 0, text[-6000:]
    assert text.count("GPU_WORKER_OK") == nproc, text[-6000:]
