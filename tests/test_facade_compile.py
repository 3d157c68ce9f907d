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


def test_c_abi_header_is_plain_c_and_the_example_links(tmp_path):
    """include/fastlio_b200.h must be consumable from C (cgo / ctypes / a C application): examples/pipeline.c walks one scan
    through every stage with -std=c99 -pedantic.  Without a GPU the program must fail loudly at fl_map_create."""
    exe = tmp_path / "pipeline_example"
    cmd = ["/usr/bin/gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "examples", "pipeline.c"), build.LIB, "-lm", "-o", str(exe), "-Wl,-rpath," + os.path.dirname(build.LIB)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    from fast_lio_b200 import api
    if api.device_count() == 0:
        run = subprocess.run([str(exe)], capture_output=True, text=True)
        assert run.returncode != 0 and "no CUDA device" in run.stderr
