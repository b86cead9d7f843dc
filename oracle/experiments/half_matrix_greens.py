"""Test infrastructure (oracle side): the stabilized Green's function computed on HALF matrices only (antiunitary symmetry),
the last piece of the executable specification for the half-matrix device path (after quaternion_qr.py and
quaternion_qr_blocked.py).  Not used by the product.

Every symmetric matrix S = [[A, B], [-conj(B), conj(A)]] is represented by its left half S_L (n x h, natural row order);
`full(S_L)` is only used where an operand really is needed in full (the left operand of a product).  Operations:

  mul(S_full, T_L)         -> (S T)_L                     half the flops of a full product
  adj_L(S_L)               -> (S^H)_L                     (S^H)_L = conj([A, B])^T, and B = -conj(S_L[h:, :]) ...
  scale rows / columns     with PAIRED diagonal scales (d_i = d_{i+h}), which keep the symmetry
  paired QR of S_L         (quaternion_qr_blocked.blocked_paired_qr, rows pair-interleaved inside) with the right-hand side
                           transformed on the fly, then the back substitution with the quaternion-upper-triangular R:
                           for i = h-1..0:  x_i = r_ii^-1 (y_i - sum_{j>i} r_ij x_j),  r_ij = 2x2 blocks [[a, b], [-conj(b), conj(a)]]

and the device path's formula

  G = U_r D_r+^-1 [ D_l+^-1 U_l^H U_r D_r+^-1 + D_l- T_l T_r^H D_r- ]^-1 D_l+^-1 U_l^H.

Check (python -m oracle.experiments.half_matrix_greens): left/right stacks over beta = 40 built with the paired UDT, G from
halves only, against the reference algorithm (zgeqp3 stacks + the reference's calculate_greens).
"""
import numpy as np

import oracle
from oracle.experiments.quaternion_qr import full_from_left as full, paired_udt, sym_residual
from oracle.experiments.quaternion_qr_blocked import blocked_paired_qr, interleave_rows


def adj_L(SL):
    """Left half of S^H from the left half of S."""
    h = SL.shape[1]
    A, mBc = SL[:h], SL[h:]                       # S = [[A, B], [-conj(B), conj(A)]],  SL = [A; -conj(B)]
    B = -np.conj(mBc)
    # S^H = [[A^H, -B^T], [B^H, A^T]]  ->  left half [A^H; B^H]
    return np.concatenate([A.conj().T, B.conj().T], axis=0)


def quat_block(e, o):
    """2x2 complex block of the quaternion (e, o) acting on a pair-interleaved row pair: column (e, o), partner column -phi."""
    return np.array([[e, -np.conj(o)], [o, np.conj(e)]])


def solve_staircase(Rint, Yint):
    """Rint: n x h, pair-interleaved rows, quaternion upper triangular (column j non-zero in rows 0..2j+1), representing the
    symmetric matrix whose pair-interleaved COLUMNS are (column j, partner).  Yint: n x k right-hand side (pair-interleaved
    rows).  Returns X (n x k, pair-interleaved rows) with R_full X = Y."""
    n, h = Rint.shape
    X = Yint.copy()
    for i in range(h - 1, -1, -1):
        rii = quat_block(Rint[2 * i, i], Rint[2 * i + 1, i])
        X[2 * i:2 * i + 2] = np.linalg.solve(rii, X[2 * i:2 * i + 2])
        if i:
            # column pair i of R_full above the diagonal: rows 0..2i-1; the partner column is -phi(column) in interleaved rows
            c = Rint[:2 * i, i]
            pc = np.empty_like(c)
            pc[0::2] = -np.conj(c[1::2])
            pc[1::2] = np.conj(c[0::2])
            X[:2 * i] -= np.outer(c, X[2 * i]) + np.outer(pc, X[2 * i + 1])
    return X


def greens_half(UlL, Dl, TlL, UrL, Dr, TrL):
    """All arguments are left halves / paired scales (length h).  Returns G_L."""
    n, h = UlL.shape
    Dlp, Dlm = np.maximum(Dl, 1.0), np.minimum(Dl, 1.0)
    Drp, Drm = np.maximum(Dr, 1.0), np.minimum(Dr, 1.0)
    two = lambda d: np.concatenate([d, d])
    UlH = full(adj_L(UlL))
    innerL = (UlH @ UrL) / two(Dlp)[:, None] / Drp[None, :] + (two(Dlm)[:, None] * (full(TlL) @ adj_L(TrL))) * Drm[None, :]
    rhsL = adj_L(UlL) / two(Dlp)[:, None]
    # paired QR of inner with the right-hand side carried along (pair-interleaved rows inside)
    p = interleave_rows(n)
    inv = np.argsort(p)
    Rint, Yint = innerL[p].copy(), rhsL[p].copy()
    blocked_paired_qr(Rint, Yint, NP=16)
    for j in range(h):
        Rint[2 * j + 2:, j] = 0.0
    # the solution X = inner^-1 rhs is symmetric; its rows come back pair-interleaved in the COLUMN-pair order of inner,
    # i.e. interleaved row 2i <-> natural row i, 2i+1 <-> i+h
    XL = solve_staircase(Rint, Yint)[inv]
    return (full(UrL) / two(Drp)[None, :]) @ XL


def main(L=4, M=400, lam=0.5):
    from oracle.experiments.stab_variants import chain, udt_geqp3

    def udt_paired_full(Y):
        QL, Dh, TL = paired_udt(Y)
        return full(QL), np.concatenate([Dh, Dh]), full(TL)

    mc = oracle.OracleDQMC(oracle.Params(L=L, slices=M, safe_mult=10, Bfield=True, lam=lam))
    mc.hsfield = np.random.RandomState(3).rand(3, L * L, M)
    h = mc.n // 2
    worst = 0.0
    for c in sorted({M // 20 * 10, M // 40 * 10, 10} - {0}):          # multiples of safe_mult: U of a finished UDT is unitary
        ref = (chain(mc, udt_geqp3, list(range(0, c)), False), chain(mc, udt_geqp3, list(range(M - 1, c - 1, -1)), True))
        mc.Ul, mc.Dl, mc.Tl = ref[0]
        mc.Ur, mc.Dr, mc.Tr = ref[1]
        Gref = mc.calculate_greens().copy()
        (Ul, Dl, Tl), (Ur, Dr, Tr) = (chain(mc, udt_paired_full, list(range(0, c)), False),
                                      chain(mc, udt_paired_full, list(range(M - 1, c - 1, -1)), True))
        GL = greens_half(Ul[:, :h], Dl[:h], Tl[:, :h], Ur[:, :h], Dr[:h], Tr[:, :h])
        G = full(GL)
        worst = max(worst, np.abs(G - Gref).max() / np.abs(Gref).max())
        print(f"L={L} M={M} slice={c}: log10 D range {np.log10(Dl.max() / Dl.min()):.0f};  |G_half - G_ref| / |G| = "
              f"{np.abs(G - Gref).max() / np.abs(Gref).max():.1e};  symmetry residual of G_ref {sym_residual(Gref):.1e}")
    return worst


if __name__ == "__main__":
    main()
    main(L=6, M=200, lam=1.0)
