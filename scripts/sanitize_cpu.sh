#!/bin/bash
# AddressSanitizer + UBSan pass over everything that runs on the CPU: the host build of the C-ABI layer
# (tests/emu/host_shim.cpp: csrc/b200fft.cu + plan programs + kernel phase bodies in the emulator) under the
# emulator / plan / schedule / host-shim test suites and the Python layer driving it.  Takes a few minutes.  (Found the dangling Step references
# of the plan builders that std::deque now rules out.)
set -e
cd "$(dirname "$0")/.."
LIB=/tmp/libb200fft_shim_asan.so
g++ -std=c++17 -O1 -g -fsanitize=address,undefined -shared -fPIC -x c++ -I tests/emu/cuda_shim tests/emu/host_shim.cpp -o $LIB -ldl -pthread
cat > /tmp/b200fft_asan_driver.py <<PY
import sys, ctypes
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import emu_util, host_shim_util
from mpifft4py_b200 import _lib
L = ctypes.CDLL("$LIB")
_lib.declare(L)
emu_util._lib = L
host_shim_util._shim = L
import pytest
files = ["tests/test_emu_plans.py", "tests/test_emu_random_plans.py", "tests/test_emu_schedule.py", "tests/test_passes.py",
         "tests/test_c2r_staging.py", "tests/test_c2c.py", "tests/test_host_shim.py", "tests/test_host_shim_multi.py",
         "tests/test_host_shim_golden.py", "tests/test_zz_ns_known_answer.py",
         # the Python layer on top of the host build (tests/cpu_engine.py), thread-ranks included
         "tests/test_ref_procedures_oracle.py", "tests/test_worker_lists_cpu.py", "tests/test_serial_functions_cpu.py",
         "tests/test_reference_suite_unmodified.py", "tests/test_engine_run.py"]
sys.exit(pytest.main(["-x", "-q", "-p", "no:cacheprovider", "-m", "not gpu"] + files))
PY
LD_PRELOAD="$(g++ -print-file-name=libasan.so) $(g++ -print-file-name=libubsan.so)" ASAN_OPTIONS=detect_leaks=0 python -u /tmp/b200fft_asan_driver.py
