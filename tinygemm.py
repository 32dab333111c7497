"""`import tinygemm` registers torch.ops.tinygemm.* - the import the reference's modules.py,
quantize.py (`import_or_skip("tinygemm")`) and tests perform (reference: TinyGemm.cpp:13-15 is an
empty pybind module whose import side effect is the op registration).  Here the ops come from the
B200 library in any4_b200/lib/."""
from any4_b200 import _native as _native

_native.load_ops()
__doc__ = "tinygemm: low-bit CUDA GEMM library (B200-native build)"
