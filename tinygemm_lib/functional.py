"""`tinygemm_lib.functional` as the reference exposes it; implementation in any4_b200.functional."""
import tinygemm  # noqa: F401  (registers torch.ops.tinygemm.*)
from any4_b200.functional import *  # noqa: F401,F403
from any4_b200.functional import valid_tinygemm_kernel_call  # noqa: F401
