"""Compiles tests/cpp/tenc_dropin_test.cpp (the C++ drop-in class used the way HM uses it) with g++ and runs it on the GPU."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_dropin_class_matches_oracle(oracle, tmp_path):
    exe = str(tmp_path / "tenc_dropin_test")
    pkg = os.path.join(ROOT, "hm-opencl_b200")
    cmd = ["g++", "-std=c++11", "-O1", "-DHMME_STANDALONE", "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(pkg, "host"),
           "-I" + os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests/cpp/tenc_dropin_test.cpp"), os.path.join(pkg, "host/TEncOpenCL.cpp"),
           "-L" + pkg, "-lhmme_b200", "-L" + os.path.join(ROOT, "oracle"), "-lhmme_oracle", "-Wl,-rpath," + pkg, "-Wl,-rpath," + os.path.join(ROOT, "oracle"),
           "-o", exe]
    subprocess.run(cmd, check=True)
    r = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    print(r.stdout)
    assert r.returncode == 0 and "PASS" in r.stdout, r.stdout[-2000:]


def test_cpp_group_over_every_visible_gpu_matches_oracle(oracle, tmp_path):
    """tests/cpp/group_test.cpp: hmme_group_* driven from plain C++ over every visible GPU (2 or more on the driver's box)."""
    exe = str(tmp_path / "group_test")
    pkg = os.path.join(ROOT, "hm-opencl_b200")
    cmd = ["g++", "-std=c++11", "-O1", "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests/cpp/group_test.cpp"),
           "-L" + pkg, "-lhmme_b200", "-L" + os.path.join(ROOT, "oracle"), "-lhmme_oracle", "-Wl,-rpath," + pkg, "-Wl,-rpath," + os.path.join(ROOT, "oracle"),
           "-o", exe]
    subprocess.run(cmd, check=True)
    r = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    print(r.stdout)
    assert r.returncode == 0 and "PASS" in r.stdout, r.stdout[-2000:]
