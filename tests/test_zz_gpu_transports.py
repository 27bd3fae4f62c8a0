"""Multi-GPU parity of the slab transports that are not the default: NCCL send/recv and the fused
peer-store transport (the y / x FFT pass writes straight into the peers' receive buffers over NVLink).
Runs last (file name) and only where the box has >= 2 GPUs."""
import os
import subprocess
import sys

import pytest

from test_gpu_multi import HERE, _free_port

pytestmark = pytest.mark.gpu


# "store" was written after round 1's GPU minutes were spent (emulator-checked only): a device failure
# of this opt-in transport is reported as xfail, a pass as XPASS.
@pytest.mark.parametrize("transport", ["nccl", pytest.param("store", marks=pytest.mark.xfail(
    strict=False, reason="opt-in transport; first device run pending"))])
@pytest.mark.parametrize("nproc", [2, 4, 8])
def test_slab_transport_parity(nproc, transport):
    import torch
    if torch.cuda.device_count() < nproc:
        pytest.skip("needs %d GPUs" % nproc)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc),
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(HERE, "gpu_dist_worker.py"), "--transport", transport]
    out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=900)
    text = out.stdout.decode("utf-8", "replace")
    assert out.returncode == 0, text[-6000:]
    assert text.count("GPU_WORKER_OK") == nproc, text[-6000:]
