"""Device runs of kernel variants that are compiled but not yet the default (B200FFT_VARIANT).  Each runs
in a subprocess with a timeout, after every other GPU test (file name), so that a fault in an experimental
kernel cannot disturb the parity suite of the default path."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


# Written after round 1's GPU minutes were spent: checked in the CPU emulator only so far, so a device
# failure is reported as xfail (and a pass as XPASS) instead of failing the suite of the default path.
@pytest.mark.xfail(strict=False, reason="opt-in kernel variant; first device run pending")
@pytest.mark.parametrize("what", ["cluster", "rowbar", "l2", "kzblock"])
def test_variant_on_device(what):
    out = subprocess.run([sys.executable, os.path.join(HERE, "gpu_variant_worker.py"), what],
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=300)
    text = out.stdout.decode("utf-8", "replace")
    assert out.returncode == 0 and "VARIANT_WORKER_OK" in text, text[-6000:]
