"""The cost-space work split every row-streaming kernel uses (csrc/stream_common.cuh::Split, plan_split, cost_to_row) checked on
the host: a small program compiled with nvcc (host code only; no GPU is touched) walks 20 000 random geometries."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_split_partitions_the_rows(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = tmp_path / "split_check"
    inc = [f"-I{os.path.join(ROOT, 'blind_image_denoising_b200', 'csrc')}", f"-I{os.path.join(ROOT, 'include')}"]
    r = subprocess.run([nvcc, "-std=c++17", "-O1", "-gencode", "arch=compute_100a,code=sm_100a", "--expt-relaxed-constexpr", *inc,
                        os.path.join(ROOT, "tests", "csrc", "split_check.cu"), "-o", str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.startswith("ok"), r.stdout + r.stderr
