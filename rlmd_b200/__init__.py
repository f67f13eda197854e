"""
rlmd_b200 - B200-native (sm_100a) engine for rlmd's multiplicative Monte-Carlo
hot path.  Python shims over a C-ABI CUDA library (include/rlmd_b200.h):

    rlmd_b200.lev_exp       same names/signatures as the reference's lev.lev_exp
    rlmd_b200.engine        the engine-level calls the shims are made of
"""
from . import _lib  # noqa: F401  (raises if the CUDA library is not built)

__all__ = ["_lib"]
