"""ORACLE (test infrastructure): build the reference's own CUDA extensions.

Compiles /root/reference/cuda/{resample2d_package,block_extractor,
local_attn_reshape}/*.{cc,cu} for sm_100a into oracle/_ref/ (git-ignored,
travels to the GPU box with gpurun).  The sources are read where they lie;
the only change is the one-token torch-2.x fix `.type()` -> `.scalar_type()`
inside AT_DISPATCH_FLOATING_TYPES, applied on the fly to a
scratch copy under /tmp that is deleted afterwards.  No reference source is
written into this repository.

The resulting modules (resample2d_cuda, block_extractor_cuda,
local_attn_reshape_cuda) are the on-device reference the GPU parity tests
use to pin both the C oracle and the product kernels; they need a GPU to
run, so nothing is executed here.

Usage: python oracle/build_ref.py   (no-op when /root/reference is absent)
"""
import os
import re
import shutil
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("FFWM_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")

EXTS = {
    "resample2d_cuda": ("cuda/resample2d_package", ["resample2d_cuda.cc", "resample2d_kernel.cu"], ["resample2d_kernel.cuh"]),
    "block_extractor_cuda": ("cuda/block_extractor", ["block_extractor_cuda.cc", "block_extractor_kernel.cu"], ["block_extractor_kernel.cuh"]),
    "local_attn_reshape_cuda": ("cuda/local_attn_reshape", ["local_attn_reshape_cuda.cc", "local_attn_reshape_kernel.cu"], ["local_attn_reshape_kernel.cuh"]),
}


def built(name):
    return os.path.exists(os.path.join(OUT, name + ".so"))


def main(force=False):
    if not os.path.isdir(REF):
        print("build_ref: %s not present, keeping prebuilt oracle/_ref" % REF)
        return 0
    if not force and all(built(n) for n in EXTS):
        return 0
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    os.environ["CC"] = "/usr/bin/gcc"
    os.environ["CXX"] = "/usr/bin/g++"
    from torch.utils import cpp_extension
    os.makedirs(OUT, exist_ok=True)
    for name, (sub, srcs, hdrs) in EXTS.items():
        if built(name) and not force:
            continue
        scratch = tempfile.mkdtemp(prefix="ffwm_ref_")
        try:
            for f in srcs + hdrs:
                text = open(os.path.join(REF, sub, f)).read()
                text = re.sub(r"\.type\(\)", ".scalar_type()", text)
                open(os.path.join(scratch, f), "w").write(text)
            bdir = os.path.join(scratch, "build")
            os.makedirs(bdir)
            cpp_extension.load(
                name=name, sources=[os.path.join(scratch, s) for s in srcs],
                build_directory=bdir, verbose=False, is_python_module=False,
                extra_cuda_cflags=["-gencode", "arch=compute_100a,code=sm_100a", "-w"],
                extra_cflags=["-w"], with_cuda=True)
            shutil.copy(os.path.join(bdir, name + ".so"), os.path.join(OUT, name + ".so"))
            print("build_ref: built", name)
        finally:
            shutil.rmtree(scratch, ignore_errors=True)
    return 0


if __name__ == "__main__":
    sys.exit(main(force="--force" in sys.argv))
