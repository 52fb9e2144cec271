"""The attention kernel's work decomposition (csrc/attention_sched.cuh) checked exhaustively on the CPU: the same
host/device functions the kernel and its combine pass use, compiled as host code (tests/native/)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_attention_schedule_partitions_every_step(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = tmp_path / "attention_schedule_check"
    subprocess.run([nvcc, "-O2", "-std=c++17", "-arch=sm_100a", "-I", os.path.join(ROOT, "self_forcing_b200", "csrc"), "-o", str(exe),
                    os.path.join(ROOT, "tests", "native", "attention_schedule_check.cu")], check=True, capture_output=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.strip().splitlines()[-1].startswith("OK"), r.stdout[-2000:]
