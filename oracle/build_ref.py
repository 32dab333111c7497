#!/usr/bin/env python
"""Build the UNMODIFIED reference tinygemm extension for sm_100a into oracle/_ref/.

Test infrastructure only. Compiles the reference's own sources where they lie
under /root/reference/tinygemm_lib (nothing is copied into this repo); the only
outputs are oracle/_ref/tinygemm.so and its ninja build directory, both
git-ignored. The module keeps its original import name (`tinygemm`) and op
namespace, so it must only ever be imported in a *separate process* from
any4_b200's own `torch.ops.tinygemm.*` registration (see oracle/ref_runner.py).

The reference's setup.py is not used: it hard-codes TORCH_CUDA_ARCH_LIST
'8.0;8.6;9.0' with no PTX (tinygemm_lib/setup.py:11), which cannot run on B200.
"""
import os
import shutil
import sys

REF = os.environ.get("ANY4_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")

SOURCES = [
    "TinyGemm.cpp",
    "TinyGemm_bf16.cu",
    "TinyGemm_int4.cu",
    "TinyGemm_int8.cu",
    "TinyGemmConvertA.cu",
    "TinyGemmConvertB.cu",
    "TinyGemmDequantize.cu",
]


def build(verbose: bool = False) -> str:
    src_dir = os.path.join(REF, "tinygemm_lib")
    if not os.path.isdir(src_dir):
        raise FileNotFoundError(f"reference sources not found at {src_dir}")
    so = os.path.join(OUT, "tinygemm.so")
    if os.path.exists(so):
        return so
    os.makedirs(os.path.join(OUT, "build"), exist_ok=True)
    os.environ["TORCH_CUDA_ARCH_LIST"] = "10.0a"
    os.environ.setdefault("MAX_JOBS", str(os.cpu_count() or 8))
    from torch.utils.cpp_extension import load

    load(
        name="tinygemm",
        sources=[os.path.join(src_dir, s) for s in SOURCES],
        extra_cflags=["-O3"],
        extra_cuda_cflags=["-O3", "--use_fast_math"],
        build_directory=os.path.join(OUT, "build"),
        is_python_module=False,
        verbose=verbose,
    )
    shutil.copy(os.path.join(OUT, "build", "tinygemm.so"), so)
    return so


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv))
