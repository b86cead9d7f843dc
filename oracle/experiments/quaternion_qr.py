"""Test infrastructure (oracle side): prototype of a UDT decomposition that works on HALF of every matrix by using the
antiunitary flavour symmetry (see antiunitary_symmetry.py), as preparation for the device path.  Not used by the product.

A matrix with the symmetry, X = [[A, B], [-conj(B), conj(A)]] (n = 2h), commutes with the antilinear map phi(v) = J conj(v),
J = [[0, 1], [-1, 0]]: its column c+h is -phi(column c), orthogonal to column c and of the same norm.  Regard the row pair
(i, i+h) as one quaternion.  A *paired* Householder step with u = x + (|x| / |q_1|) (x_1 e_1 + x_{1+h} e_{1+h}) and

    H = 1 - (2 / |u|^2) (u u^H + phi(u) phi(u)^H)              (real tau, two mutually orthogonal reflector vectors)

commutes with phi, maps x onto the rows (1, 1+h) and therefore maps the partner column -phi(x) onto the same row pair: one
step eliminates TWO columns, needs the dot products u^H c and phi(u)^H c = -u^T J c of the left-half columns c only, and the
whole factorization has h = n/2 sequential steps instead of n and half the flops.  R is upper triangular as a quaternion
matrix (2x2 blocks [[a, b], [-conj(b), conj(a)]] on the diagonal), D = quaternion modulus of its diagonal.

This script checks, on slice-matrix chains of the oracle (L = 4, beta = 4, graded over ~10 orders of magnitude):
  * Q (rebuilt from its left half) is unitary, Q R = X P to rounding, D is sorted/graded like the pivoted QR's,
  * G computed from a stack built with this UDT agrees with the oracle's G (LAPACK zgeqp3 path) to ~1e-13.

    python -m oracle.experiments.quaternion_qr
"""
import numpy as np

import oracle


def phi(v):
    h = v.shape[0] // 2
    w = np.conj(v)
    return np.concatenate([w[h:], -w[:h]], axis=0)


def full_from_left(XL):
    """[[A], [-conj(B)]] (n x h)  ->  [[A, B], [-conj(B), conj(A)]]: column c+h = -phi(column c)."""
    return np.concatenate([XL, -phi(XL)], axis=1)


def sym_residual(X):
    h = X.shape[0] // 2
    return np.abs(X[:, h:] + phi(X[:, :h])).max()


def paired_udt(X):
    """UDT of a symmetric X from its left half.  Returns (QL, D_h, TL): left halves of U and T, and the h distinct scales
    (each belongs to a column pair).  Column 'pivoting' = one sort of the left-half columns by norm."""
    n = X.shape[0]
    h = n // 2
    XL = X[:, :h].copy()
    perm = np.argsort(-np.linalg.norm(XL, axis=0), kind="stable")
    Wk = XL[:, perm].copy()                      # working copy, left-half columns in sorted order
    us, taus = [], []
    for j in range(h):
        rows = np.r_[j:h, h + j:n]               # active rows: quaternion rows j..h-1
        x = Wk[rows, j]
        m = h - j
        nx = np.linalg.norm(x)
        q1 = np.hypot(abs(x[0]), abs(x[m]))
        u = x.copy()
        if nx > 0 and q1 > 0:
            s = nx / q1
            u[0] += s * x[0]
            u[m] += s * x[m]
            tau = 2.0 / np.vdot(u, u).real
        else:
            tau = 0.0
        pu = phi(u)
        C = Wk[rows, j:]
        C -= tau * (np.outer(u, u.conj() @ C) + np.outer(pu, pu.conj() @ C))
        Wk[rows, j:] = C
        Wk[rows[1:m], j] = 0.0                    # eliminated entries: exact zeros, as in a LAPACK R (they are rounding
        Wk[rows[m + 1:], j] = 0.0                 # noise of the size of the LARGE scales and would wreck D^-1 R)
        us.append((rows, u, tau))
    # R as a quaternion upper triangle: left-half columns of the complex R (rows j and j+h of column c >= j)
    RL = Wk                                       # n x h; entries below the quaternion diagonal are ~0
    D = np.sqrt(np.abs(RL[np.arange(h), np.arange(h)]) ** 2 + np.abs(RL[h + np.arange(h), np.arange(h)]) ** 2)
    # Q left half = H_1 ... H_h applied to the first h unit vectors
    QL = np.zeros((n, h), dtype=complex)
    QL[np.arange(h), np.arange(h)] = 1.0
    for rows, u, tau in reversed(us):
        pu = phi(u)
        C = QL[rows, :]
        C -= tau * (np.outer(u, u.conj() @ C) + np.outer(pu, pu.conj() @ C))
        QL[rows, :] = C
    # T = D^-1 R P^T as a symmetric matrix: its left half has, in sorted column order, the columns D^-1 R; the rows come in
    # quaternion pairs (j, j+h) scaled by the same D_j
    Dfull = np.concatenate([D, D])
    Rfull_sorted = full_from_left(RL)             # n x n, columns in sorted order (left block) and their partners
    Tsorted = Rfull_sorted / Dfull[:, None]
    T = np.empty_like(Tsorted)
    T[:, perm] = Tsorted[:, :h]
    T[:, h + perm] = Tsorted[:, h:]
    return QL, D, T[:, :h]


def main(L=4, M=40, lam=0.5):
    om = oracle.OracleDQMC(oracle.Params(L=L, slices=M, safe_mult=10, Bfield=True, lam=lam))
    om.init(np.random.RandomState(3).rand(3, L * L, M))
    n = om.n
    h = n // 2
    # a slice-matrix chain as the stack sees it: B_10 ... B_1, then (chain . U D) again
    X = np.eye(n, dtype=complex)
    for s in range(10):
        X = om.multiply_B_left(s, X)
    print(f"n={n}: symmetry residual of the chain {sym_residual(X):.1e}, cond {np.linalg.cond(X):.1e}")
    Uo, Do, To = oracle.dqmc.decompose_udt(X.copy())
    QL, D, TL = paired_udt(X)
    Q, T = full_from_left(QL), full_from_left(TL)
    Dfull = np.concatenate([D, D])
    print(f"  |Q^H Q - 1| = {np.abs(Q.conj().T @ Q - np.eye(n)).max():.1e}")
    rec = (Q * Dfull[None, :]) @ T
    print(f"  |Q D T - X| / |col| = {(np.abs(rec - X) / np.linalg.norm(X, axis=0)[None, :]).max():.1e}")
    print(f"  log10 D range {np.log10(Dfull.max() / Dfull.min()):.1f} (zgeqp3: {np.log10(Do.max() / Do.min()):.1f});  cond(T) = {np.linalg.cond(T):.1e} (zgeqp3: {np.linalg.cond(To):.1e})")

    # The device path's stabilization with the stack UDT replaced by the paired one: left and right
    # stacks over the whole imaginary-time axis, G from the Loh-split formula, against the reference algorithm (zgeqp3 stack
    # + the reference's calculate_greens).  Harness of stab_variants.py.
    from oracle.experiments.stab_variants import chain, greens_loh, udt_geqp3, udt_presort

    def udt_paired(Y):
        assert sym_residual(Y) < 1e-9 * np.abs(Y).max(), "matrix handed to decompose_udt is not symmetric"
        QLl, Dh, TLl = paired_udt(Y)
        return full_from_left(QLl), np.concatenate([Dh, Dh]), full_from_left(TLl)

    mc = oracle.OracleDQMC(oracle.Params(L=L, slices=M, safe_mult=10, Bfield=True, lam=lam))
    mc.hsfield = np.random.RandomState(3).rand(3, L * L, M)
    for c in sorted({M // 20 * 10, M // 40 * 10, 10} - {0}):          # multiples of safe_mult
        res = {}
        for name, udt in (("geqp3", udt_geqp3), ("presort", udt_presort), ("paired", udt_paired)):
            res[name] = (chain(mc, udt, list(range(0, c)), False), chain(mc, udt, list(range(M - 1, c - 1, -1)), True))
        mc.Ul, mc.Dl, mc.Tl = res["geqp3"][0]
        mc.Ur, mc.Dr, mc.Tr = res["geqp3"][1]
        Gref = mc.calculate_greens().copy()
        sc = np.abs(Gref).max()
        line = f"  M={M} slice={c}: log10 D range {np.log10(mc.Dl.max() / mc.Dl.min()):.0f};  |G_loh(stack) - G_ref| / |G|:"
        for name in ("geqp3", "presort", "paired"):
            (Ul, Dl, Tl), (Ur, Dr, Tr) = res[name]
            G1 = greens_loh(Ul, Dl, Tl, Ur, Dr, Tr)
            line += f"  {name} {np.abs(G1 - Gref).max() / sc:.1e}"
            if name == "paired":
                line += f" (symmetry residual {sym_residual(G1):.1e}, cond T_l {np.linalg.cond(Tl):.1e} vs presort {np.linalg.cond(res['presort'][0][2]):.1e})"
        print(line)


if __name__ == "__main__":
    main()                       # beta = 4
    main(M=400)                  # beta = 40: 40 stack blocks, D spans > 100 orders of magnitude
    main(L=6, M=200, lam=1.0)    # n = 144, stronger coupling
