"""ctypes binding of libdqmc_b200.so (the C ABI declared in include/dqmc_b200.h).

There is no fallback: if the shared library has not been built (``python -c 'import __graft_entry__ as g;
g.build()'``) importing the product path raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdqmc_b200.so")

OP_HOP_HALF_B, OP_HOP_A, OP_HOP_HALF_INV_B, OP_HOP_INV_A, OP_MU, OP_MU_INV, OP_HOP_HALF_A, OP_HOP_HALF_INV_A = range(8)
B_LEFT, B_RIGHT, B_INV_LEFT, B_INV_RIGHT, B_DAGGER_LEFT = range(5)


class DqmcParams(C.Structure):
    _fields_ = [("L", C.c_int32), ("flv", C.c_int32), ("opdim", C.c_int32), ("slices", C.c_int32),
                ("safe_mult", C.c_int32), ("edrun", C.c_int32), ("all_checks", C.c_int32), ("device", C.c_int32),
                ("delay", C.c_int32), ("reserved", C.c_int32),
                ("delta_tau", C.c_double), ("lambda_", C.c_double), ("r", C.c_double), ("c", C.c_double),
                ("u", C.c_double)]


class DqmcError(RuntimeError):
    pass


_lib = None

_P = C.c_void_p
_D = C.POINTER(C.c_double)
_I64 = C.POINTER(C.c_int64)
_I32 = C.POINTER(C.c_int32)

# name -> (restype, argtypes); every symbol include/dqmc_b200.h declares
SIGNATURES = {
    "dqmc_create": (C.c_int, [C.POINTER(_P), C.POINTER(DqmcParams)]),
    "dqmc_destroy": (C.c_int, [_P]),
    "dqmc_last_error": (C.c_char_p, [_P]),
    "dqmc_set_operator": (C.c_int, [_P, C.c_int, C.c_int64, C.c_int64, _I64, _I64, _P, C.c_int]),
    "dqmc_set_neighbors": (C.c_int, [_P, _I64]),
    "dqmc_set_hsfield": (C.c_int, [_P, _D]),
    "dqmc_get_hsfield": (C.c_int, [_P, _D]),
    "dqmc_set_greens": (C.c_int, [_P, _D]),
    "dqmc_get_greens": (C.c_int, [_P, _D]),
    "dqmc_get_state": (C.c_int, [_P, _I32, _I32]),
    "dqmc_set_state": (C.c_int, [_P, C.c_int32, C.c_int32]),
    "dqmc_build_stack": (C.c_int, [_P]),
    "dqmc_propagate": (C.c_int, [_P, _I32, _I32]),
    "dqmc_wrap_greens": (C.c_int, [_P, _D, C.c_int32, C.c_int32]),
    "dqmc_multiply_B": (C.c_int, [_P, C.c_int, C.c_int32, _D]),
    "dqmc_calculate_greens_from": (C.c_int, [_P, _D, _D, _D, _D, _D, _D, _D]),
    "dqmc_logdet": (C.c_int, [_P, _D]),
    "dqmc_decompose_udt": (C.c_int, [_P, _D, _D, _D, _D]),
    "dqmc_local_updates": (C.c_int, [_P, C.c_double, _D, C.c_int64, _I64, _I64, _D]),
    "dqmc_sweep": (C.c_int, [_P, C.c_int32, C.c_double, _D, C.c_int64, _I64, _I64, _D]),
    "dqmc_set_uniforms": (C.c_int, [_P, _D, C.c_int64]),
    "dqmc_calc_boson_action": (C.c_int, [_P, _D]),
    "dqmc_measure_chi_dynamic": (C.c_int, [_P, _D]),
    "dqmc_global_update": (C.c_int, [_P, C.c_double, _D, C.c_double, _D, _I32, _I32]),
    "dqmc_measure_tdgfs": (C.c_int, [_P]),
    "dqmc_get_tdgf": (C.c_int, [_P, C.c_int, C.c_int32, _D]),
    "dqmc_free_tdgfs": (C.c_int, [_P]),
    "dqmc_inv_sum_udts": (C.c_int, [_P, _D, _D, _D, _D, _D, _D, _D]),
    "dqmc_timers": (C.c_int, [_P, _D, C.c_int32]),
    "dqmc_set_timing": (C.c_int, [_P, C.c_int32]),
    "dqmc_checks": (C.c_int, [_P, _D, _I64]),
    "dqmc_sync": (C.c_int, [_P]),
    "dqmc_bench_kernel": (C.c_int, [_P, C.c_int, C.c_int, _D]),
    "dqmc_test_zgemm": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _D, _D, C.c_int, _D, C.c_int, _D, _D,
                                  C.c_int]),
    "dqmc_test_qr_paired": (C.c_int, [_P, _D, _D, _D, _D, _D, C.c_int32]),
    "dqmc_test_udt": (C.c_int, [_P, _D, _D, _D, _D, C.c_int32]),
    "dqmc_lu_profile": (C.c_int, [_P, C.c_int32, _I64]),
    "dqmc_qr_profile": (C.c_int, [_P, C.c_int32, _I64]),
    "dqmc_kernel_launches": (C.c_int64, [_P]),
}


def load():
    """Load libdqmc_b200.so and attach the prototypes.  Raises if it is missing (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DqmcError(f"{LIB_PATH} not built: run __graft_entry__.build() (nvcc, sm_100a). "
                        "dqmc_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def dptr(a):
    return a.ctypes.data_as(_D)


def cplx_in(a, shape):
    """Column-major complex128 host array laid out like Julia's Array{ComplexF64}."""
    a = np.asarray(a, dtype=np.complex128)
    assert a.shape == tuple(shape), (a.shape, shape)
    return np.asfortranarray(a)


def cplx_buf(shape):
    return np.zeros(shape, dtype=np.complex128, order="F")
