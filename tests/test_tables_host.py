"""The device-side table / flat-plan builder (monocle3_b200/csrc/tables.cu) is made of
__host__ __device__ per-element functions: tests/native/host_tables_check.cu runs exactly those
on the CPU (serial prefix sums in place of the scan kernels) and compares every item, slot,
row pointer and flat-plan piece with the host builders they replace.  No GPU needed."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("nvcc") is None, reason="nvcc not on PATH")
def test_device_table_bodies_match_host_builders(tmp_path):
    exe = str(tmp_path / "host_tables_check")
    subprocess.run(["nvcc", "-std=c++17", "-O1", "-gencode", "arch=compute_100a,code=sm_100a", "-I",
                    os.path.join(ROOT, "monocle3_b200", "csrc"), "-o", exe,
                    os.path.join(ROOT, "tests", "native", "host_tables_check.cu")], check=True, capture_output=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-3000:]
    assert "ok: device table bodies agree" in r.stdout
