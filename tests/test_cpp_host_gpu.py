"""The compiled-language host layer (include/cubecl_b200.hpp): builds on CPU, runs its checks on the GPU box."""
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
EXE = ROOT / "cubecl_b200" / "lib" / "sum_things_cpp"


def test_cpp_example_is_built():
    from cubecl_b200 import build
    build.build()
    assert EXE.exists()


@pytest.mark.gpu
def test_cpp_host_layer_end_to_end():
    r = subprocess.run([str(EXE)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "[15, 15, 15, 15]" in r.stdout and "cmma golden ok" in r.stdout and "deferred error surfaced" in r.stdout
    assert "block-scaled matmul ok" in r.stdout
