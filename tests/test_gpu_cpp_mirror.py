"""GPU: the header-only C++ host mirror (include/helio_voxel_cuda.hpp) compiles with plain g++ against
the C ABI and reproduces the reference's sphere fixture."""
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def _build(tmp_path):
    exe = tmp_path / "test_cpp_mirror"
    lib_dir = ROOT / "helio_b200"
    subprocess.run(["g++", "-std=c++17", "-O1", "-I", str(ROOT / "include"), str(ROOT / "tests" / "cpp" / "test_cpp_mirror.cpp"),
                    "-L", str(lib_dir), "-lhelio_voxel_cuda", f"-Wl,-rpath,{lib_dir}", "-o", str(exe)], check=True)
    return exe


def test_cpp_mirror_compiles_without_cuda_headers(tmp_path):
    _build(tmp_path)


@pytest.mark.gpu
def test_cpp_mirror_matches_the_reference_sphere(tmp_path):
    out = subprocess.run([str(_build(tmp_path))], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "OK cpp mirror" in out.stdout


def test_cpp_bounded_extraction_publisher_replays_the_reference_tests(tmp_path):
    """Host-only part of the C++ mirror: runs on the CPU."""
    exe = tmp_path / "test_cpp_publisher"
    lib_dir = ROOT / "helio_b200"
    subprocess.run(["g++", "-std=c++17", "-O1", "-I", str(ROOT / "include"), str(ROOT / "tests" / "cpp" / "test_cpp_publisher.cpp"),
                    "-L", str(lib_dir), "-lhelio_voxel_cuda", f"-Wl,-rpath,{lib_dir}", "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0 and "OK cpp publisher" in out.stdout, out.stdout + out.stderr
