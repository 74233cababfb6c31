"""TEST INFRASTRUCTURE ONLY -- builds the reference's own CUDA extension (`hawp.base._C`, the repo's one CUDA kernel:
third-party/hawp/hawp/base/csrc/{binding.cpp,linesegment.cu}) from the sources WHERE THEY LIE under /root/reference,
exactly as third-party/hawp/hawp/base/csrc/__init__.py does (torch.utils.cpp_extension.load of those two files), but
cross-compiled for sm_100a and with the output in oracle/_ref/ (git-ignored, travels to the GPU box with the snapshot).
The `-m gpu` test tests/test_hawp_oracle.py::test_gpu_encodels_vs_reference_kernel loads the built module on the B200
and compares neat_encodels with the REAL reference kernel.  No reference source is copied into this repository.

Second artefact (round 2): the reference's own PYTHON path for the step -- the seven unmodified files VolSDFNetwork /
VolSDFLoss import (code/model/{density,embedder,ray_sampler}.py, code/model/networks/{neat_wfr_rend_a,loss_wfr}.py,
code/utils/{rend_util,general}.py) -- archived byte for byte into oracle/_ref/neat_ref_code.zip (git-ignored build output,
like the .so; imported through zipimport by oracle/ref_shim.py).  It is what `bench.py --impl reference`, `cpu_baseline`
and `gpu_eager_baseline` time on the GPU box, where /root/reference does not exist: the UNMODIFIED reference, not a port.

    python oracle/build_ref.py            # needs /root/reference; a no-op message otherwise"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
NAME = "hawp_ref_C"
REF_CSRC = os.path.join(os.environ.get("NEAT_REFERENCE_ROOT", "/root/reference"), "third-party", "hawp", "hawp", "base", "csrc")


REF_CODE = os.path.join(os.environ.get("NEAT_REFERENCE_ROOT", "/root/reference"), "code")
CODE_ZIP = os.path.join(OUT, "neat_ref_code.zip")
CODE_FILES = ["model/density.py", "model/embedder.py", "model/ray_sampler.py", "model/networks/neat_wfr_rend_a.py",
              "model/networks/loss_wfr.py", "utils/rend_util.py", "utils/general.py"]


def stage_code():
    """Archive the reference's step files (unmodified) into oracle/_ref/neat_ref_code.zip; returns the path or None."""
    import zipfile
    srcs = [os.path.join(REF_CODE, f) for f in CODE_FILES]
    if not all(os.path.exists(s) for s in srcs):
        return CODE_ZIP if os.path.exists(CODE_ZIP) else None
    if os.path.exists(CODE_ZIP) and all(os.path.getmtime(CODE_ZIP) >= os.path.getmtime(s) for s in srcs):
        return CODE_ZIP
    os.makedirs(OUT, exist_ok=True)
    with zipfile.ZipFile(CODE_ZIP, "w", zipfile.ZIP_DEFLATED) as z:
        for d in ("model/", "model/networks/", "utils/"):     # explicit directory entries: namespace packages in a zip
            z.writestr(zipfile.ZipInfo(d), "")
        for f, s in zip(CODE_FILES, srcs):
            z.write(s, f)
    return CODE_ZIP


def built_path():
    p = os.path.join(OUT, NAME + ".so")
    return p if os.path.exists(p) else None


def load_built():
    """Import the prebuilt module (GPU box: /root/reference is absent, only oracle/_ref/ travelled)."""
    import importlib.util
    import torch  # noqa: F401  (the extension links against libtorch)
    p = built_path()
    if p is None:
        return None
    try:
        spec = importlib.util.spec_from_file_location(NAME, p)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    except (ImportError, OSError) as e:   # built against another torch / image: the comparison is skipped, not failed
        print("oracle/_ref/%s.so does not load here: %s" % (NAME, e))
        return None
    return mod


def build(verbose=False):
    stage_code()
    srcs = [os.path.join(REF_CSRC, "binding.cpp"), os.path.join(REF_CSRC, "linesegment.cu")]
    if not all(os.path.exists(s) for s in srcs):
        print("reference sources not present (%s): nothing built" % REF_CSRC)
        return built_path()
    p = built_path()
    if p and all(os.path.getmtime(p) >= os.path.getmtime(s) for s in srcs):
        return p
    os.makedirs(OUT, exist_ok=True)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")          # no GPU here: name the target instead of probing
    from torch.utils.cpp_extension import load
    load(name=NAME, sources=srcs, build_directory=OUT, verbose=verbose,
         extra_cuda_cflags=["-gencode", "arch=compute_100a,code=sm_100a"])
    return built_path()


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv))
