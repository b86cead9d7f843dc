"""CPU checks of the executable specification of the half-matrix stabilization path (oracle/experiments/paired_panel_spec.py):
the paired Householder QR in the device's conventions is a QR (Q unitary, Q R = X, exact quaternion structure of the compact-WY T),
and a UDT stack built with it reproduces the oracle's Green's function (reference algorithm: zgeqp3 stacks + calculate_greens,
src/linalg.jl:20-39, src/stack.jl:338-369) -- the property the GPU tests then hold the CUDA kernels to."""
import numpy as np

import oracle
from oracle.experiments.paired_panel_spec import full_from_left_int, paired_qr_device, panel_device_dataflow, psi


def test_panel_is_a_block_reflector_with_quaternion_T():
    rs = np.random.RandomState(5)
    n, h = 96, 48
    XL = (rs.randn(n, h) + 1j * rs.randn(n, h)) * np.sort(np.logspace(20, -20, h))[::-1][None, :]
    R, V, T, d = panel_device_dataflow(XL[:, :16])
    Q = np.eye(n) - V @ T @ V.conj().T
    assert np.abs(Q.conj().T @ Q - np.eye(n)).max() < 1e-14
    assert (np.abs(Q.conj().T @ XL[:, :16] - R) / np.linalg.norm(XL[:, :16], axis=0)).max() < 1e-14
    # T(2c+1, 2j+1) = conj(T(2c, 2j)), T(2c, 2j+1) = -conj(T(2c+1, 2j)): exact (the kernel derives the odd columns this way)
    assert np.array_equal(T[1::2, 1::2], T[0::2, 0::2].conj())
    assert np.array_equal(T[0::2, 1::2], -T[1::2, 0::2].conj())
    # partner columns of V, moduli of the quaternion diagonal
    assert all(np.array_equal(V[:, 2 * c + 1], psi(V[:, 2 * c])) for c in range(16))
    mod = np.sqrt(np.abs(R[2 * np.arange(16), np.arange(16)]) ** 2 + np.abs(R[2 * np.arange(16) + 1, np.arange(16)]) ** 2)
    assert np.allclose(d, mod, rtol=1e-14)


def test_blocked_paired_qr_and_zero_column():
    rs = np.random.RandomState(6)
    n, h = 128, 64
    XL = (rs.randn(n, h) + 1j * rs.randn(n, h)) * np.sort(np.logspace(30, -30, h))[::-1][None, :]
    XL[:, 40] = 0.0                                  # a zero column: tau = 0, R column 0, the factorization goes on
    XL[2 * 7:2 * 7 + 2, 7] = 0.0                     # a zero quaternion on the diagonal: direction (1, 0) is used
    rhs = np.zeros((n, h), dtype=complex)
    rhs[2 * np.arange(h), np.arange(h)] = 1.0
    RL, QHL, dabs = paired_qr_device(XL, rhs)
    QH, Rf, Xf = full_from_left_int(QHL), full_from_left_int(RL), full_from_left_int(XL)
    assert np.abs(QH @ QH.conj().T - np.eye(n)).max() < 1e-13
    cn = np.linalg.norm(Xf, axis=0)
    cn[cn == 0.0] = 1.0
    assert (np.abs(QH.conj().T @ Rf - Xf) / cn[None, :]).max() < 1e-13
    assert dabs[40] < 1e-13 * dabs.max()


def test_half_matrix_greens_matches_the_reference_algorithm():
    from oracle.experiments.half_matrix_greens import main
    assert main(L=4, M=100) < 1e-11                  # beta = 10: |G_half - G_ref| / |G|
