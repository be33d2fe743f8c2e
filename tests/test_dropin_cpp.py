"""The C++ drop-in classes (uv-slam_b200/host): same Evaluate() signatures as the reference's cost
functions and the optimization() call surface, built with g++ against the C ABI library."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "test_dropin")


def _build():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "uv-slam_b200", "host")])
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-Wall", "-o", EXE, os.path.join(ROOT, "tests", "cpp", "test_dropin.cpp"),
                           "-L" + os.path.join(ROOT, "uv-slam_b200", "host"), "-luvs_host",
                           "-L" + os.path.join(ROOT, "uv-slam_b200", "csrc"), "-luvs_b200",
                           "-Wl,-rpath," + os.path.join(ROOT, "uv-slam_b200", "host"),
                           "-Wl,-rpath," + os.path.join(ROOT, "uv-slam_b200", "csrc")])


def test_dropin_builds_and_fails_loudly_without_gpu():
    import torch
    _build()
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu test")
    p = subprocess.run([EXE], capture_output=True, text=True)
    assert p.returncode == 3 and "no CPU fallback" in p.stdout


@pytest.mark.gpu
def test_dropin_known_answers_on_gpu():
    _build()
    p = subprocess.run([EXE], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0 and "drop-in OK" in p.stdout, p.stdout + p.stderr
