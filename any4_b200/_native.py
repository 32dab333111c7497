"""Loader for the native libraries (built in-tree by any4_b200/build.py).

* `capi()`   -> ctypes handle on libtinygemm_b200.so, the C ABI of include/tinygemm_b200.h
* `load_ops()` registers `torch.ops.tinygemm.*` by loading tinygemm_ops.so

There is deliberately no fallback: if the libraries are missing the import fails and says how
to build them; nothing in the package computes on the CPU.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_DIR = os.environ.get("ANY4_B200_LIB_DIR") or os.path.join(_HERE, "lib")  # override: A/B-testing two builds
CAPI_PATH = os.path.join(LIB_DIR, "libtinygemm_b200.so")
OPS_PATH = os.path.join(LIB_DIR, "tinygemm_ops.so")

_capi = None
_ops_loaded = False

# every symbol include/tinygemm_b200.h declares (checked by tests/test_capi_cpu.py)
CAPI_SYMBOLS = (
    "tg_set_option", "tg_last_error", "tg_version", "tg_launch_count", "tg_reset_launch_count",
    "tg_convert_to_A", "tg_convert_from_A", "tg_convert_to_B", "tg_convert_from_B",
    "tg_convert_to_Aint4", "tg_convert_to_Aint8", "tg_convert_to_Bint4", "tg_convert_to_Bint8",
    "tg_gemm_w4_rm", "tg_gemm_w4_rm_ws", "tg_gemm_w4_rm_workspace_bytes", "tg_repack_Aint4_to_Bint4", "tg_gemm_w4_rm_hostio", "tg_gemm_w4_rm_sharded", "tg_gemm_w4_rm_exchange", "tg_gemm_w4_rm_exchange_silu_pairs", "tg_gemm_w8_rm", "tg_gemm_w16_rm",
    "tg_gemm_tc_workspace_bytes", "tg_gemm_w4_tc", "tg_gemm_w8_tc", "tg_gemm_w16_tc",
    "tg_dequant_int4",
    "tg_quantize_any4_rows",
    "tg_decode_add_rmsnorm", "tg_decode_silu_mul", "tg_decode_rope_attention", "tg_gemm_w4_rm_silu_pairs",
)


def _missing(path):
    return RuntimeError(
        f"{path} not found: the CUDA extension has not been built. Run `python -m any4_b200.build` "
        "(needs nvcc; sm_100a only). There is no CPU fallback."
    )


def capi():
    """ctypes.CDLL of the C-ABI library with argument types declared."""
    global _capi
    if _capi is not None:
        return _capi
    if not os.path.exists(CAPI_PATH):
        raise _missing(CAPI_PATH)
    lib = ctypes.CDLL(CAPI_PATH, mode=ctypes.RTLD_GLOBAL)
    vp, i64, i32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int
    lib.tg_last_error.restype = ctypes.c_char_p
    lib.tg_version.restype = ctypes.c_char_p
    lib.tg_launch_count.restype = ctypes.c_uint64
    lib.tg_reset_launch_count.restype = None
    lib.tg_set_option.argtypes = [i32, i32]
    lib.tg_set_option.restype = i32
    lib.tg_convert_to_A.argtypes = [vp, vp, i64, i64, vp]
    lib.tg_convert_from_A.argtypes = [vp, vp, i64, i64, vp]
    lib.tg_convert_to_B.argtypes = [vp, vp, i64, i64, i32, vp]
    lib.tg_convert_from_B.argtypes = [vp, vp, i64, i64, i32, vp]
    for name in ("tg_convert_to_Aint4", "tg_convert_to_Aint8", "tg_convert_to_Bint4", "tg_convert_to_Bint8"):
        getattr(lib, name).argtypes = [vp, vp, i64, i64, i32, vp]
    lib.tg_gemm_w4_rm.argtypes = [vp, vp, vp, vp, vp, vp, i64, i64, i64, i32, i32, i32, i32, i32, vp]
    if hasattr(lib, "tg_repack_Aint4_to_Bint4"):
        lib.tg_repack_Aint4_to_Bint4.argtypes = [vp, vp, i64, i64, i32, i32, vp]
        lib.tg_gemm_w4_rm_workspace_bytes.argtypes = [i64, i64, i64, i32]
        lib.tg_gemm_w4_rm_workspace_bytes.restype = ctypes.c_size_t
        lib.tg_gemm_w4_rm_ws.argtypes = [vp, vp, vp, vp, vp, vp, i64, i64, i64, i32, i32, i32, i32, i32, vp, ctypes.c_size_t, vp]
    if hasattr(lib, "tg_gemm_w4_rm_hostio"):
        lib.tg_gemm_w4_rm_hostio.argtypes = [vp, vp, vp, vp, vp, vp, vp, i64, i64, i64, i32, i32, i32, i32, i32, vp]
    if hasattr(lib, "tg_gemm_w4_rm_sharded"):  # (absent only in older builds loaded through ANY4_B200_LIB_DIR)
        lib.tg_gemm_w4_rm_sharded.argtypes = [vp, i32, i64, vp, vp, vp, vp, vp, i64, i64, i64, i32, i32, i32, i32, vp]
    if hasattr(lib, "tg_gemm_w4_rm_exchange"):
        u32 = ctypes.c_uint32
        lib.tg_gemm_w4_rm_exchange.argtypes = [vp, vp, i32, u32, i32, i64, vp, vp, vp, vp, vp, i64, i64, i64, i32, i32, i32, i32, vp]
        if hasattr(lib, "tg_gemm_w4_rm_exchange_silu_pairs"):
            lib.tg_gemm_w4_rm_exchange_silu_pairs.argtypes = lib.tg_gemm_w4_rm_exchange.argtypes
    if hasattr(lib, "tg_quantize_any4_rows"):
        lib.tg_quantize_any4_rows.argtypes = [vp, vp, i64, i64, i32, i32, i32, ctypes.c_float, vp, vp, vp, vp, vp, vp, i32, vp]
    lib.tg_gemm_w8_rm.argtypes = [vp, vp, vp, vp, i64, i64, i64, i32, i32, i32, i32, vp]
    lib.tg_gemm_w16_rm.argtypes = [vp, vp, vp, i64, i64, i64, i32, i32, i32, vp]
    lib.tg_gemm_tc_workspace_bytes.argtypes = [i64, i64, i64]
    lib.tg_gemm_tc_workspace_bytes.restype = ctypes.c_size_t
    lib.tg_gemm_w4_tc.argtypes = [vp, vp, vp, vp, vp, vp, i64, i64, i64, i32, i32, i32, i32, i32, i32, vp, vp]
    lib.tg_gemm_w8_tc.argtypes = [vp, vp, vp, vp, i64, i64, i64, i32, i32, i32, i32, i32, vp, vp]
    lib.tg_gemm_w16_tc.argtypes = [vp, vp, vp, i64, i64, i64, i32, i32, i32, i32, vp, vp]
    lib.tg_dequant_int4.argtypes = [vp, vp, i64, vp]
    if hasattr(lib, "tg_decode_add_rmsnorm"):
        f32 = ctypes.c_float
        lib.tg_decode_add_rmsnorm.argtypes = [vp, vp, vp, vp, i64, f32, i32, vp]
        lib.tg_decode_silu_mul.argtypes = [vp, vp, i64, i32, vp]
        lib.tg_gemm_w4_rm_silu_pairs.argtypes = [vp, vp, vp, vp, vp, vp, i64, i64, i64, i32, i32, i32, i32, vp]
        lib.tg_decode_rope_attention.argtypes = [vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, f32, i32, vp]
    for name in CAPI_SYMBOLS:
        if not hasattr(lib, name):
            continue
        fn = getattr(lib, name)
        if name.startswith(("tg_convert", "tg_gemm_w", "tg_dequant", "tg_decode")):
            fn.restype = i32
    _capi = lib
    return lib


def load_ops():
    """Register torch.ops.tinygemm.* (idempotent)."""
    global _ops_loaded
    if _ops_loaded:
        return
    import torch

    if not os.path.exists(OPS_PATH):
        raise _missing(OPS_PATH)
    capi()  # make sure the dependency is resolvable even without the rpath
    torch.ops.load_library(OPS_PATH)
    _ops_loaded = True


def last_error():
    return capi().tg_last_error().decode()
