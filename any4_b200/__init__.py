"""any4_b200: B200-native (sm_100a) tinygemm - the weight-only small-batch GEMV path of
facebookresearch/any4 (tinygemm_lib/ + modules.py), rebuilt from scratch.

Layout of the package
  csrc/            CUDA kernels + the C ABI (include/tinygemm_b200.h) + the torch op layer
  build.py         in-tree nvcc build
  _native.py       loaders (ctypes C ABI, torch.ops.tinygemm.*)
  functional.py    mirror of tinygemm_lib/functional.py (16 wrappers)
  utils.py         mirror of tinygemm_lib/utils.py (host-side quantizers)
  modules.py       mirror of modules.py (Int4Linear / Int8Linear / Any4Linear) + row-sharded variant
Importing the package does not touch the GPU; `any4_b200.functional` / `modules` load the ops.
"""
__all__ = ["functional", "utils", "modules"]
