"""`tinygemm_lib.utils` as the reference exposes it; implementation in any4_b200.utils."""
from any4_b200.utils import *  # noqa: F401,F403
