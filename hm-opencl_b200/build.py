"""Builds libhmme_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python hm-opencl_b200/build.py [--force]

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "hmme_b200.cu")
SRCS = [SRC, os.path.join(HERE, "csrc", "hmme_group.cu")]
DEPS = SRCS + sorted(os.path.join(HERE, "csrc", f) for f in os.listdir(os.path.join(HERE, "csrc")) if f.endswith(".cuh")) + [
    os.path.join(os.path.dirname(HERE), "include", "hmme_b200.h")]
OUT = os.path.join(HERE, "libhmme_b200.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-shared",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def stale():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(d) > t for d in DEPS)


def build(force=False, quiet=True):
    if not (force or stale()):
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + ["-o", OUT] + SRCS + ["-ldl"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    log = os.path.join(HERE, "build.log")
    with open(log, "w") as f:
        f.write(" ".join(cmd) + "\n" + r.stdout)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("nvcc failed building libhmme_b200.so (see %s)" % log)
    if not quiet:
        print(r.stdout)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, quiet=False))
