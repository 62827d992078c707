"""The header-only C++ mirror of the reference interface (include/proxsuite_b200.hpp) and the
example that re-states examples/cpp/first_example_dense.cpp on it: it must compile with a plain
g++ against the C-ABI library, fail loudly without a GPU (exit code 3, no CPU fallback), and - on a
GPU box - reproduce the reference's known answer x* = [1, 0.5, -1] (test/src/cvxpy.py:24-46) and
solve a BatchQP through solve_in_parallel."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "proxsuite_b200")


def build_example(tmp_path):
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else shutil.which("g++")
    assert gxx, "g++ is required"
    exe = os.path.join(str(tmp_path), "first_example_dense")
    cmd = [gxx, "-std=c++17", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"),
           os.path.join(ROOT, "examples", "first_example_dense.cpp"), "-L" + LIBDIR, "-lpqp_b200", "-Wl,-rpath," + LIBDIR, "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def has_gpu():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def test_cpp_mirror_compiles_and_fails_loudly_without_gpu(tmp_path):
    exe = build_example(tmp_path)
    if has_gpu():
        pytest.skip("GPU present: covered by the gpu-marked test")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 3, (r.returncode, r.stdout, r.stderr)
    assert "no CPU fallback" in r.stdout


@pytest.mark.gpu
def test_cpp_mirror_example_on_gpu(tmp_path):
    exe = build_example(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert "64 / 64 solved" in r.stdout
