"""Builds and runs the C++ host-side mirror of the reference's operator interface (include/dn_backend.hpp) through
tests/cpp/test_backend.cpp: on CPU only against the oracle + the documented known answers, and on the GPU box
against libdeepnet_b200.so with host-vs-CUDA comparison (the C++ analogue of Tensor.Test/CudaTests.fs)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build(tmp_path):
    from oracle import host_tensor
    host_tensor.build()
    exe = str(tmp_path / "test_backend")
    lib = os.path.join(ROOT, "deepnet_b200", "lib")
    ora = os.path.join(ROOT, "oracle", "_build")
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "test_backend.cpp"), "-o", exe,
                           "-L", lib, "-ldeepnet_b200", "-L", ora, "-ldn_oracle",
                           f"-Wl,-rpath,{lib}", f"-Wl,-rpath,{ora}", "-Wl,-rpath,/usr/local/cuda/lib64",
                           "-L/usr/local/cuda/lib64"])
    return exe


def test_cpp_mirror_on_oracle(tmp_path):
    out = subprocess.run([build(tmp_path), "oracle"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "ALL OK" in out.stdout, out.stdout + out.stderr


@pytest.mark.gpu
def test_cpp_mirror_cuda_vs_oracle(tmp_path):
    out = subprocess.run([build(tmp_path), "both"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "ALL OK (both)" in out.stdout, out.stdout + out.stderr
