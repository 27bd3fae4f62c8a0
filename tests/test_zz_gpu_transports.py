"""Multi-GPU parity of the slab transports that are not the default: NCCL send/recv and the fused
peer-store transport (the y / x FFT pass writes straight into the peers' receive buffers over NVLink).
Runs last (file name) and only where the box has >= 2 GPUs."""
import os
import subprocess
import sys

import pytest

from test_gpu_multi import HERE, _free_port

pytestmark = pytest.mark.gpu


# "store", the "kz" pipeline and "p2p" for pencil / line were written after round 1's GPU minutes were
# spent (emulator-checked only).  A protocol mistake between ranks would show up as a hang that only the
# subprocess timeout ends, so these opt-in modes run on request only (B200FFT_EXPERIMENTAL=1, set by
# scripts/gpu_round2_multi.sh) and do not put the default suite's wall clock at risk.
_PENDING = pytest.mark.skipif(not os.environ.get("B200FFT_EXPERIMENTAL"),
                              reason="opt-in mode, first device run pending: set B200FFT_EXPERIMENTAL=1")


@pytest.mark.parametrize("transport,pipeline", [
    ("nccl", "x"), pytest.param("p2p", "x", marks=_PENDING), pytest.param("store", "x", marks=_PENDING), pytest.param("nccl", "kz", marks=_PENDING),
    pytest.param("p2p", "kz", marks=_PENDING), pytest.param("store", "kz", marks=_PENDING),
    pytest.param("nccl", "pencil-chunks", marks=_PENDING), pytest.param("p2p", "pencil-chunks", marks=_PENDING)])
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
