"""CPU check of the register FFTs and packed-pair helpers of meyda_b200/csrc/mb_fft.cuh: the header compiled for the host
(scalar fall-back of the f32x2 helpers) against a float64 DFT.  No GPU, no oracle."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_register_fft_and_packed_helpers_on_the_host(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not found")
    exe = str(tmp_path / "fft_host_check")
    r = subprocess.run([nvcc, "-std=c++17", "-O1", "-o", exe, os.path.join(ROOT, "tests", "host", "fft_host_check.cu")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout
    got = {k: float(v) for k, v in (line.split() for line in out.strip().splitlines())}
    assert set(got) == {"fft32", "fft16", "fft8", "fft4", "fft2", "cmul", "split"}
    for name, err in got.items():
        assert err < 2e-6, (name, err)  # float32 arithmetic against float64: a few ulp of the peak
