"""Build libneat_b200.so (C ABI, no torch dependency) in-tree with nvcc for sm_100a.

    python -m neat_b200.build            # incremental
    python -m neat_b200.build --force
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libneat_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
SOURCES = ["api.cu", "plan.cpp", "junction.cpp"]


def _newer_than(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(os.path.dirname(HERE), "include", "neat_b200.h"))
    if not force and not _newer_than(OUT, deps):
        return OUT
    cmd = [NVCC, "-O3", "-std=c++17", "-lineinfo", "-shared", "-Xcompiler", "-fPIC,-Wall", *ARCH,
           "-Xptxas", "-v" if verbose else "-warn-spills",
           "-o", OUT] + [os.path.join(CSRC, s) for s in SOURCES]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed")
    if verbose:
        sys.stderr.write(r.stdout + r.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
