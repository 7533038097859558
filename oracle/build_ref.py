"""Build the REFERENCE's own native extension (number_system/mx/cpp/*.{cpp,cu}) from the sources
where they lie under /root/reference, into oracle/_ref/ (git-ignored, travels with gpurun).

    python oracle/build_ref.py

This does not run the reference's build system (custom_extensions.py); it hands the same four
source files to torch.utils.cpp_extension with an explicit sm_100a arch.  The result is TEST
INFRASTRUCTURE: a second oracle for the MX core (quantize_mx_func_cpp / quantize_elemwise_func_cpp
on the CPU, the *_cuda entry points on the GPU box).  No reference source is copied into the repo.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_CPP = "/root/reference/number_system/mx/cpp"
OUT = os.path.join(HERE, "_ref")
NAME = "msq_ref_ext"


def build(verbose=False):
    so = os.path.join(OUT, NAME + ".so")
    if not os.path.isdir(REF_CPP):
        return so if os.path.exists(so) else None          # GPU box: use the prebuilt file
    if os.path.exists(so):
        return so
    os.makedirs(OUT, exist_ok=True)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0")   # no visible GPU in the authoring container
    from torch.utils.cpp_extension import load
    load(name=NAME, sources=[os.path.join(REF_CPP, f) for f in ("funcs.cpp", "mx.cu", "elemwise.cu", "reduce.cu")],
         build_directory=OUT, extra_cuda_cflags=["-gencode", "arch=compute_100a,code=sm_100a"], verbose=verbose)
    return so


def load_ref():
    """import the prebuilt extension (None if it was never built)"""
    so = os.path.join(OUT, NAME + ".so")
    if not os.path.exists(so):
        return None
    import importlib.util
    import torch  # noqa: F401  (the extension links against libtorch)
    spec = importlib.util.spec_from_file_location(NAME, so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv))
