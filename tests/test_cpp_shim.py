"""include/helfem_b200.hpp -- the header-only C++ shims with the reference's class and member names --
compiles against the C ABI with a plain column-major matrix type and behaves like the reference's classes
(errors before compute_tei, wrong-sized matrices, no CPU fallback).  tests/cpp/shim_check.cpp."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else shutil.which("g++")
    if cxx is None:
        pytest.skip("no host C++ compiler")
    import helfem_b200 as hb
    hb.lib()
    exe = str(tmp_path / "shim_check")
    libdir = os.path.join(ROOT, "helfem_b200")
    subprocess.check_call([cxx, "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "shim_check.cpp"), "-o", exe, "-L", libdir,
                           "-lhelfemqc_b200", "-Wl,-rpath," + libdir])
    return exe


def test_cpp_shim_host_only(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("host-only behaviour is checked on a machine without a GPU")
    out = subprocess.run([_build(tmp_path)], capture_output=True, text=True)
    assert out.returncode == 0 and "OK host-only" in out.stdout, out.stdout + out.stderr


@pytest.mark.gpu
def test_cpp_shim_on_gpu(tmp_path):
    out = subprocess.run([_build(tmp_path), "gpu"], capture_output=True, text=True)
    assert out.returncode == 0 and "OK gpu" in out.stdout, out.stdout + out.stderr
