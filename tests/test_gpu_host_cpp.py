"""The C++ host mirror of the reference's ECS-facing physics surface (gears_b200/host/gears_physics.hpp): compile a
test program written like the reference's own unit tests against the C ABI, run it on the GPU."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "host_cpp", "test_update_physics.cpp")


def _compile(tmp_path):
    import __graft_entry__ as g
    g.build()
    exe = str(tmp_path / "test_update_physics")
    lib_dir = os.path.join(ROOT, "gears_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-o", exe, SRC, "-L" + lib_dir,
                           "-lgears_physics_cuda", "-Wl,-rpath," + lib_dir])
    return exe


def test_host_cpp_compiles_and_links(tmp_path):
    """CPU: the header mirrors the C ABI it calls (compiles warning-free and links against the built library)."""
    assert os.path.exists(_compile(tmp_path))


@pytest.mark.gpu
def test_host_cpp_update_physics(tmp_path):
    r = subprocess.run([_compile(tmp_path)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "HOST_CPP_OK" in r.stdout, r.stdout + r.stderr
