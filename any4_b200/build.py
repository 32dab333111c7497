"""In-tree build of the native code (nvcc, sm_100a only; no JIT cache, no CPU fallback).

Two artefacts, both under any4_b200/lib/ (git-ignored, shipped to the GPU box by gpurun):

  libtinygemm_b200.so   the CUDA kernels behind the C ABI of include/tinygemm_b200.h
                        (no torch dependency; this is the drop-in boundary)
  tinygemm_ops.so       the torch custom-op layer (torch.ops.tinygemm.*) that re-creates the
                        reference's operator surface on top of that C ABI

`python -m any4_b200.build` builds both; `build_all()` is what __graft_entry__.build() calls.
"""
import hashlib
import os
import shutil
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
_VARIANT = os.environ.get("TG_BUILD_VARIANT", "")  # e.g. "_trace": a second build next to the product one (ANY4_B200_LIB_DIR)
LIB = os.path.join(HERE, "lib" + _VARIANT)
OBJ = os.path.join(HERE, "build" + _VARIANT)
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

CUDA_HOME = os.environ.get("CUDA_HOME", "/usr/local/cuda")
NVCC = os.path.join(CUDA_HOME, "bin", "nvcc")

KERNEL_SOURCES = ["capi.cu", "convert.cu", "gemv_w4_b.cu", "gemv_w4_tc.cu", "gemv_w4_a.cu", "gemv_generic.cu", "decode_ops.cu", "quantize.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
    "-I", INCLUDE,
] + os.environ.get("TG_NVCC_EXTRA", "").split()  # e.g. -DTG_W4_TRACE for scripts/trace_kernel.py

CAPI_LIB = os.path.join(LIB, "libtinygemm_b200.so")
OPS_LIB = os.path.join(LIB, "tinygemm_ops.so")


def _digest(paths, extra=""):
    h = hashlib.sha256(extra.encode())
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(p.encode())
            h.update(f.read())
    return h.hexdigest()


def _run(cmd, log):
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    with open(log, "w") as f:
        f.write(" ".join(cmd) + "\n" + proc.stdout)
    if proc.returncode != 0:
        sys.stderr.write(proc.stdout)
        raise RuntimeError(f"build step failed: {' '.join(cmd[:3])} ... (log: {log})")
    return proc.stdout


def _headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hs.append(os.path.join(INCLUDE, "tinygemm_b200.h"))
    return hs


def build_capi(force=False, verbose=False):
    """nvcc -> libtinygemm_b200.so (one object per .cu, compiled in parallel)."""
    os.makedirs(LIB, exist_ok=True)
    os.makedirs(OBJ, exist_ok=True)
    srcs = [os.path.join(CSRC, s) for s in KERNEL_SOURCES]
    stamp = os.path.join(OBJ, "capi.stamp")
    want = _digest(srcs + _headers(), " ".join(NVCC_FLAGS))
    if not force and os.path.exists(CAPI_LIB) and os.path.exists(stamp) and open(stamp).read() == want:
        return CAPI_LIB
    if not os.path.exists(NVCC):
        raise RuntimeError(f"nvcc not found at {NVCC}: the CUDA library cannot be built (there is no CPU fallback)")
    procs = []
    objs = []
    for s in srcs:
        o = os.path.join(OBJ, os.path.basename(s) + ".o")
        objs.append(o)
        log = o + ".log"
        cmd = [NVCC, *NVCC_FLAGS, "-c", s, "-o", o]
        procs.append((cmd, log, subprocess.Popen(cmd, stdout=open(log, "w"), stderr=subprocess.STDOUT)))
    failed = False
    for cmd, log, pr in procs:
        if pr.wait() != 0:
            failed = True
            sys.stderr.write(open(log).read())
        elif verbose:
            sys.stdout.write(open(log).read())
    if failed:
        raise RuntimeError("nvcc failed")
    _run([NVCC, "-shared", "-cudart", "shared", "-o", CAPI_LIB, *objs,
          "-Xlinker", "-rpath", "-Xlinker", os.path.join(CUDA_HOME, "lib64")], os.path.join(OBJ, "link_capi.log"))
    with open(stamp, "w") as f:
        f.write(want)
    return CAPI_LIB


def build_ops(force=False, verbose=False):
    """g++ -> tinygemm_ops.so: TORCH_LIBRARY registration, links libtinygemm_b200.so via $ORIGIN."""
    import torch
    from torch.utils import cpp_extension as ce

    os.makedirs(LIB, exist_ok=True)
    os.makedirs(OBJ, exist_ok=True)
    src = os.path.join(CSRC, "torch_ops.cpp")
    stamp = os.path.join(OBJ, "ops.stamp")
    want = _digest([src, os.path.join(INCLUDE, "tinygemm_b200.h")], torch.__version__)
    if not force and os.path.exists(OPS_LIB) and os.path.exists(stamp) and open(stamp).read() == want:
        return OPS_LIB
    inc = []
    for p in ce.include_paths(device_type="cuda") if "device_type" in ce.include_paths.__code__.co_varnames else ce.include_paths(cuda=True):
        inc += ["-I", p]
    inc += ["-I", sysconfig.get_paths()["include"], "-I", INCLUDE, "-I", os.path.join(CUDA_HOME, "include")]
    torch_lib = os.path.join(os.path.dirname(torch.__file__), "lib")
    abi = int(torch._C._GLIBCXX_USE_CXX11_ABI)
    cmd = [
        "g++", "-O2", "-std=c++17", "-fPIC", "-shared", f"-D_GLIBCXX_USE_CXX11_ABI={abi}",
        "-DTORCH_API_INCLUDE_EXTENSION_H", *inc, src, "-o", OPS_LIB,
        "-L", LIB, "-ltinygemm_b200", "-L", torch_lib, "-ltorch", "-ltorch_cpu", "-lc10", "-lc10_cuda", "-ltorch_cuda",
        "-Wl,-rpath,$ORIGIN", f"-Wl,-rpath,{torch_lib}",
    ]
    out = _run(cmd, os.path.join(OBJ, "ops.log"))
    if verbose:
        sys.stdout.write(out)
    with open(stamp, "w") as f:
        f.write(want)
    return OPS_LIB


def build_all(force=False, verbose=False):
    return build_capi(force, verbose), build_ops(force, verbose)


def clean():
    for d in (LIB, OBJ):
        shutil.rmtree(d, ignore_errors=True)


if __name__ == "__main__":
    if "clean" in sys.argv:
        clean()
    else:
        print(build_all(force="-f" in sys.argv, verbose="-v" in sys.argv))
