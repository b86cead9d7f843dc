"""Numerical experiment behind DESIGN.md's choice of stabilization algebra.  Test infrastructure.

Compares, at beta=40 (M=400, safe_mult=10), the reference's calculate_greens (two column-pivoted
QRs + LU solve + inverse, stack.jl:338-369) with GPU-friendlier formulations fed by the same or by
differently-built UDT stacks:

  loh      G = Ur Drp^-1 [Dlp^-1 Ul^† Ur Drp^-1 + Dlm Tl Tr^† Drm]^-1 Dlp^-1 Ul^†   (one LU, no QR)
  presort  UDT by sorting columns by norm once, then unpivoted Householder QR
  nopiv    UDT by unpivoted Householder QR

usage: python -m oracle.experiments.stab_variants [L] [lambda]
"""
import sys
import time

import numpy as np
import scipy.linalg as sla

from oracle import Params, OracleDQMC
from oracle import dqmc as odq


def udt_geqp3(A):
    return odq.decompose_udt(A)


def udt_presort(A):
    nrm = np.linalg.norm(A, axis=0)
    piv = np.argsort(-nrm, kind="stable")
    Q, R = sla.qr(A[:, piv], mode="full", check_finite=False)
    D = np.abs(np.real(np.diag(R)))
    T = np.empty_like(R)
    T[:, piv] = R / D[:, None]
    return Q, D, T


def udt_nopiv(A):
    Q, R = sla.qr(A, mode="full", check_finite=False)
    D = np.abs(np.real(np.diag(R)))
    return Q, D, R / D[:, None]


def greens_loh(Ul, Dl, Tl, Ur, Dr, Tr):
    Dlp, Dlm = np.maximum(Dl, 1.0), np.minimum(Dl, 1.0)
    Drp, Drm = np.maximum(Dr, 1.0), np.minimum(Dr, 1.0)
    inner = (Ul.conj().T @ Ur) / Dlp[:, None] / Drp[None, :] + (Dlm[:, None] * (Tl @ Tr.conj().T)) * Drm[None, :]
    rhs = Ul.conj().T / Dlp[:, None]
    X = sla.solve(inner, rhs, check_finite=False)
    return (Ur / Drp[None, :]) @ X


def chain(mc, udt, slices, dagger):
    n = mc.n
    U = np.eye(n, dtype=complex); D = np.ones(n); T = np.eye(n, dtype=complex)
    sm = mc.p.safe_mult
    for k, s in enumerate(slices):
        U = mc.multiply_daggered_B_left(s, U) if dagger else mc.multiply_B_left(s, U)
        if (k + 1) % sm == 0:
            U = U * D[None, :]
            U, D, Tn = udt(U)
            T = Tn @ T
    return U, D, T


def main():
    L = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    lam = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
    M = 400
    p = Params(L=L, slices=M, delta_tau=0.1, safe_mult=10, lam=lam)
    mc = OracleDQMC(p)
    rs = np.random.RandomState(1234)
    mc.hsfield = rs.rand(3, L * L, M)
    n = mc.n
    for c in (200, 100, 10):   # G(c): left = B(c-1)..B(0), right = B(c)^†..B(M-1)^†
        t0 = time.time()
        res = {}
        for name, udt in (("geqp3", udt_geqp3), ("presort", udt_presort), ("nopiv", udt_nopiv)):
            L_ = chain(mc, udt, list(range(0, c)), False)
            R_ = chain(mc, udt, list(range(M - 1, c - 1, -1)), True)
            res[name] = (L_, R_)
        mc.Ul, mc.Dl, mc.Tl = res["geqp3"][0]
        mc.Ur, mc.Dr, mc.Tr = res["geqp3"][1]
        Gref = mc.calculate_greens().copy()
        sc = np.max(np.abs(Gref))
        print(f"L={L} lam={lam} slice={c}: D range {mc.Dl.min():.1e}..{mc.Dl.max():.1e}  ({time.time()-t0:.1f}s)")
        for name in ("geqp3", "presort", "nopiv"):
            (Ul, Dl, Tl), (Ur, Dr, Tr) = res[name]
            G1 = greens_loh(Ul, Dl, Tl, Ur, Dr, Tr)
            mc.Ul, mc.Dl, mc.Tl, mc.Ur, mc.Dr, mc.Tr = Ul, Dl, Tl, Ur, Dr, Tr
            G2 = mc.calculate_greens()
            print(f"   stack={name:8s}  loh: {np.max(np.abs(G1-Gref))/sc:.2e}   ref-calc_greens: {np.max(np.abs(G2-Gref))/sc:.2e}")


if __name__ == "__main__":
    main()
