"""Half-matrix (antiunitary-symmetric) stabilization path: the paired Householder QR (`qr_panel_paired_kernel`, `qr_factor_paired`)
against its executable specification `oracle/experiments/paired_panel_spec.py`, and the paired decompose_udt! against the
properties the reference's UDT has (linalg.jl:20-39: U unitary, U D T = X, D graded) plus the symmetry of every factor.

Tolerances: a Householder QR is unique only up to rounding-level differences that the grading of the columns amplifies, so the
comparison with the NumPy specification is columnwise relative (1e-11); the properties are at rounding level (1e-12).
"""
import numpy as np
import pytest

from oracle.experiments.paired_panel_spec import full_from_left_int, paired_qr_device

pytestmark = pytest.mark.gpu


def _mk(L):
    from dqmc_b200 import DQMC, Params
    return DQMC(Params(L=L, slices=20, safe_mult=10, Bfield=False, lambda_=0.5), device=0)


def _sym_full(AL):
    """natural-order symmetric matrix from its left half [A; -conj(B)]."""
    n, h = AL.shape
    top, bot = AL[:h], AL[h:]
    return np.block([[top, -np.conj(bot)], [bot, np.conj(top)]])


@pytest.mark.parametrize("L,span,lookahead", [(4, 0, False), (4, 30, True), (8, 0, True), (8, 60, True), (12, 40, True)])
def test_qr_paired_vs_spec(L, span, lookahead):
    mc = _mk(L)
    n, h = mc.n, mc.n // 2
    rs = np.random.RandomState(100 + L + span)
    XL = (rs.randn(n, h) + 1j * rs.randn(n, h)) * np.sort(np.logspace(span, -span, h))[::-1][None, :]
    rhs = np.zeros((n, h), dtype=complex)
    rhs[2 * np.arange(h), np.arange(h)] = 1.0
    R, QHL, V, Tf, dabs = mc.test_qr_paired(XL, rhs, lookahead=lookahead)
    assert np.all(np.isfinite(R)) and np.all(np.isfinite(QHL)) and np.all(np.isfinite(V)) and np.all(np.isfinite(Tf))
    # properties
    QH = full_from_left_int(QHL)
    Rf = full_from_left_int(R)
    Xf = full_from_left_int(XL)
    e_unit = np.abs(QH @ QH.conj().T - np.eye(n)).max()
    e_rec = (np.abs(QH.conj().T @ Rf - Xf) / np.linalg.norm(Xf, axis=0)[None, :]).max()
    below = max(np.abs(R[2 * j + 2:, j]).max() if 2 * j + 2 < n else 0.0 for j in range(h))
    mod = np.sqrt(np.abs(R[2 * np.arange(h), np.arange(h)]) ** 2 + np.abs(R[2 * np.arange(h) + 1, np.arange(h)]) ** 2)
    e_d = np.abs(dabs[:h] - mod).max() / mod.max()
    # specification
    Rs, QHs, ds = paired_qr_device(XL, rhs)
    cn = np.linalg.norm(XL, axis=0)[None, :]
    e_R = (np.abs(R - Rs) / cn).max()
    e_Q = np.abs(QHL - QHs).max()
    print(f"\npaired QR n={n} grading 1e+-{span}: |QQ^H-1| {e_unit:.1e}, |QR-X|/|col| {e_rec:.1e}, below-diagonal {below:.1e}, "
          f"dabs {e_d:.1e}; vs spec: R {e_R:.1e}, Q^H {e_Q:.1e}")
    assert below == 0.0
    assert e_unit < 1e-12 and e_rec < 1e-12 and e_d < 1e-13
    assert np.array_equal(dabs[:h], dabs[h:])
    assert e_R < 1e-11 and e_Q < 1e-11


@pytest.mark.parametrize("L,span", [(4, 20), (8, 50), (16, 50)])
def test_udt_paired_properties(L, span):
    mc = _mk(L)
    n, h = mc.n, mc.n // 2
    rs = np.random.RandomState(7 + L)
    scales = np.logspace(span, -span, h)[rs.permutation(h)]
    AL = (rs.randn(n, h) + 1j * rs.randn(n, h)) * scales[None, :]
    X = _sym_full(AL)
    U, D, T = mc.test_udt(X, paired=True)
    U0, D0, T0 = mc.test_udt(X, paired=False)
    e_unit = np.abs(U.conj().T @ U - np.eye(n)).max()
    rec = (U * D[None, :]) @ T
    e_rec = (np.abs(rec - X) / np.linalg.norm(X, axis=0)[None, :]).max()
    e_sym = max(np.abs(U - _sym_full(U[:, :h])).max(), np.abs(T - _sym_full(T[:, :h])).max() / np.abs(T).max())
    print(f"\npaired UDT n={n}: |U^H U-1| {e_unit:.1e}, |UDT-X|/|col| {e_rec:.1e}, symmetry {e_sym:.1e}, "
          f"log10 D range {np.log10(D.max() / D.min()):.1f} (sort-once QR: {np.log10(D0.max() / D0.min()):.1f}), "
          f"cond T {np.linalg.cond(T):.1e} (sort-once: {np.linalg.cond(T0):.1e})")
    assert e_unit < 1e-12 and e_rec < 1e-12 and e_sym < 1e-14
    assert np.log10(D.max() / D.min()) > 0.9 * 2 * span          # D carries the grading (linalg.jl:29-33)
    assert np.array_equal(D[:h], D[h:])
