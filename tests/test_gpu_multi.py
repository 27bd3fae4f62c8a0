"""Multi-rank parity on the device: one process per GPU (copy engines + NCCL) where the box has >= 2 GPUs, and
-- on ANY box with a GPU -- 2 / 4 / 8 rank processes sharing GPU 0 over the peer-mapped transports."""
import os
import signal
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("nproc", [2, 4, 8])
def test_multi_gpu_parity(nproc):
    import torch
    if torch.cuda.device_count() < nproc:
        pytest.skip("needs %d GPUs" % nproc)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc),
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(HERE, "gpu_dist_worker.py")]
    out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=900)
    text = out.stdout.decode("utf-8", "replace")
    assert out.returncode == 0, text[-6000:]
    assert text.count("GPU_WORKER_OK") == nproc, text[-6000:]


def run_group(cmd, timeout):
    """Run a torchrun command in its own process group; on a timeout the whole group goes (a hung rank must not
    keep the GPU busy behind the test's back).  Returns (returncode or None on timeout, output text)."""
    proc = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, start_new_session=True)
    try:
        out, _ = proc.communicate(timeout=timeout)
        return proc.returncode, out.decode("utf-8", "replace")
    except subprocess.TimeoutExpired:
        try:
            os.killpg(proc.pid, signal.SIGKILL)
        except ProcessLookupError:
            pass
        out, _ = proc.communicate()
        return None, out.decode("utf-8", "replace")


def shared_gpu_run(nproc, timeout=300):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc),
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(HERE, "gpu_dist_worker.py"), "--share-gpu"]
    rc, text = run_group(cmd, timeout)
    assert rc is not None, "timed out:\n" + text[-6000:]
    assert rc == 0, text[-6000:]
    assert text.count("GPU_WORKER_OK") == nproc, text[-6000:]


@pytest.mark.parametrize("nproc", [2, 4])
def test_multi_rank_shared_gpu(nproc):
    """The distributed transforms with all ranks on GPU 0 (separate processes, CUDA IPC between them): every class /
    alignment / communication layout / dealias mode against the oracle, the goldens of the unmodified reference with
    this rank count, both known answers -- through the copy-engine and fused-store transports (gpu_dist_worker.py:
    shared_gpu).  Needs one GPU, so the single-GPU driver run covers slab P > 1, pencil and line too.  Both passed on a
    B200 (profiles/r02_shared_gpu/); the 8-rank run (4x2 and 2x4 pencil grids) is tests/test_zz_shared_gpu_8.py."""
    shared_gpu_run(nproc)
