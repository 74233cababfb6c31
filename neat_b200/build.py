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


def build(force=False, verbose=False, variant=None, defines=()):
    """variant / defines: an A/B build `libneat_b200.<variant>.so` with extra -D flags (scripts/whatif.py loads it through
    NEAT_LIB_VARIANT); the product library is always the plain build."""
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(os.path.dirname(HERE), "include", "neat_b200.h"))
    out = os.path.join(HERE, "libneat_b200.%s.so" % variant) if variant else OUT
    if not force and not variant and not _newer_than(out, deps):
        return out
    cmd = [NVCC, "-O3", "-std=c++17", "-lineinfo", "-shared", "-Xcompiler", "-fPIC,-Wall", *ARCH, *defines,
           "-Xptxas", "-v" if verbose else "-warn-spills",
           "-o", out] + [os.path.join(CSRC, s) for s in SOURCES]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed")
    if verbose:
        sys.stderr.write(r.stdout + r.stderr)
    return out


if __name__ == "__main__":
    var = sys.argv[sys.argv.index("--variant") + 1] if "--variant" in sys.argv else None
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, variant=var,
                defines=[a for a in sys.argv[1:] if a.startswith("-D")]))
