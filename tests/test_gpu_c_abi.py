"""The C ABI consumed from plain C (no Python in the call path): tests/c_abi/smoke.c is compiled
against include/icnn_b200.h + libicnn_b200.so and run on the GPU."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def test_c_consumer(tmp_path):
    exe = str(tmp_path / "smoke")
    lib = os.path.join(ROOT, "icnn_b200")
    subprocess.check_call(["gcc", "-O1", "-std=c99", "-I", os.path.join(ROOT, "include"), "-I", "/usr/local/cuda/include",
                           os.path.join(ROOT, "tests", "c_abi", "smoke.c"), "-o", exe,
                           "-L", lib, "-l:libicnn_b200.so", "-L", "/usr/local/cuda/lib64", "-lcudart", "-lm",
                           "-Wl,-rpath," + lib + ":/usr/local/cuda/lib64"])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "C ABI OK" in out.stdout
