"""dqmc_b200 — B200-native (sm_100a) hot path of determinant quantum Monte Carlo for the O(3) spin-fermion
model: local-update sweep, Green's-function wrap and UDT stabilization behind the interface of carstenbauer/dqmc.

`DQMC`, `Params` and `Lattice` mirror the reference's types; all arithmetic of the path runs in
libdqmc_b200.so (hand-written CUDA, C ABI in include/dqmc_b200.h).  There is no CPU fallback.
"""
from .model import Params, Lattice  # noqa: F401
from .dqmc import DQMC, UniformStream  # noqa: F401
from .lib import DqmcError, LIB_PATH, SIGNATURES  # noqa: F401
