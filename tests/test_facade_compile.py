"""CPU: the drop-in facade headers compile and link against libfastlio_b200.so."""
import os
import subprocess

from fast_lio_b200 import build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_facades_compile_and_link(tmp_path):
    exe = tmp_path / "facade_check"
    cmd = ["/usr/bin/g++", "-std=c++14", "-Wall", "-Wno-unused",
           "-I", os.path.join(ROOT, "include"),                    # the facades
           "-I", os.path.join(ROOT, "tests", "facade"),            # stands in for the reference's include/ (mock esekfom.hpp)
           "-I", os.path.join(ROOT, "oracle", "shim"),             # pcl/point_types.h + Eigen::aligned_allocator shim
           os.path.join(ROOT, "tests", "facade", "facade_check.cpp"),
           "-o", str(exe), build.LIB, "-Wl,-rpath," + os.path.dirname(build.LIB)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
