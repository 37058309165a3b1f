"""Compiles and runs tests/cpp/test_boundary.cpp: the C++ adapters of include/mgpu_adapters.h (GpuIndex_c::MultiQuery,
GpuMatchQueue_c) driven the way the reference's RTN.WeightBoundary gtest drives CSphIndex::MultiQuery."""
import os
import subprocess

import pytest

from conftest import has_gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    from manticoresearch_b200 import build as b
    lib = b.build()
    exe = str(tmp_path / "test_boundary")
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "test_boundary.cpp"),
                    "-o", exe, lib, b.WRITER_LIB, "-Wl,-rpath," + os.path.dirname(lib)], check=True)
    return exe


@pytest.mark.skipif(has_gpu(), reason="checks the no-GPU failure mode of the C++ adapters")
def test_cpp_adapters_fail_loudly_without_gpu(tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe, str(tmp_path), "--expect-no-device"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "NO_DEVICE" in r.stdout and "no CPU fallback" in r.stdout


@pytest.mark.gpu
def test_cpp_boundary_like_rtn_weight_boundary(tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe, str(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "boundary OK" in r.stdout
