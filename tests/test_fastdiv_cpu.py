"""make_fastdiv (csrc/ptx.cuh): the multiply-high division constants of the convolution / pooling kernels, checked on
the host against integer division (tests/cpp/fastdiv_check.cu is compiled with nvcc, no GPU needed)."""
import os
import shutil
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def test_fastdiv_matches_integer_division(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = str(tmp_path / "fastdiv_check")
    subprocess.run([nvcc, "-std=c++17", "-Wno-deprecated-gpu-targets", "-o", exe,
                    os.path.join(HERE, "cpp", "fastdiv_check.cu")], check=True, capture_output=True, timeout=300)
    out = subprocess.run([exe], check=True, capture_output=True, text=True, timeout=120).stdout
    assert out.startswith("ok "), out
