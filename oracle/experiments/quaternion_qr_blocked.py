"""Test infrastructure (oracle side): executable specification of the BLOCKED paired Householder QR for the half-matrix
device path (next step after quaternion_qr.py; not used by the product yet).  It fixes the conventions a CUDA version needs:

* storage: the left half of a symmetric matrix, X_L (n x h, n = 2h), with the rows in PAIR-INTERLEAVED order
  (quaternion row i = complex rows 2i, 2i+1 = natural rows i, i+h), so that the active part of every step is a contiguous
  trailing block, as in an ordinary QR;
* in that order  phi(v)[2i] = conj(v[2i+1]),  phi(v)[2i+1] = -conj(v[2i]);
* step j works on column j from row t = 2j: reflector pair (w, phi(w)) normalised to w[t] = 1, w[t+1] = 0 (so it is stored in
  the eliminated entries of column j, LAPACK style; phi(w) is generated on the fly), real tau = 2 / |w|^2;
  the quaternion diagonal of R (rows t, t+1 of column j) stays in place;
* a panel is NP pair-steps = 2 NP reflectors; compact WY form  H_1 ... H_NP = 1 - V T V^H  with V = [w_1, phi(w_1), w_2, ...]
  and the usual forward/columnwise larft recurrence (tau real, tau_{2k} = tau_{2k+1});
* trailing update and right-hand sides:  C <- C - V (T^H (V^H C))  on the remaining left-half columns, rows >= 2 * (panel start).

Checks (python -m oracle.experiments.quaternion_qr_blocked), on symmetric matrices graded over up to 60 orders of magnitude:
Q unitary, Q R = X, and Q^H applied to a right-hand side through the panels equals Q^H from the explicit Q (all ~1e-15).
"""
import numpy as np

from oracle.experiments.quaternion_qr import full_from_left


def interleave_rows(n):
    h = n // 2
    p = np.empty(n, dtype=int)
    p[0::2] = np.arange(h)
    p[1::2] = h + np.arange(h)
    return p                      # X_int = X_nat[p]


def phi_int(v):
    w = np.empty_like(v)
    w[0::2] = np.conj(v[1::2])
    w[1::2] = -np.conj(v[0::2])
    return w


def paired_reflector(x):
    """x: trailing part of a column (pair-interleaved, even length).  Returns (w, tau, d0, d1): H = 1 - tau (w w^H + phi(w) phi(w)^H)
    maps x onto its first quaternion entry (d0, d1); w[0] = 1, w[1] = 0."""
    nx = np.linalg.norm(x)
    q = np.hypot(abs(x[0]), abs(x[1]))
    if nx == 0.0:
        w = np.zeros_like(x); w[0] = 1.0
        return w, 0.0, 0.0, 0.0
    if q == 0.0:
        # first quaternion entry exactly zero: any unit quaternion direction works; take (1, 0)
        u = x.copy(); u[0] += nx
        d0, d1 = -nx, 0.0
    else:
        s = nx / q
        u = x.copy(); u[0] += s * x[0]; u[1] += s * x[1]
        d0, d1 = -s * x[0], -s * x[1]
    a, b = u[0], u[1]
    N = abs(a) ** 2 + abs(b) ** 2
    w = (np.conj(a) * u + b * phi_int(u)) / N          # right-multiplication by the inverse quaternion: w[0] = 1, w[1] = 0
    w[0], w[1] = 1.0, 0.0
    tau = 2.0 / np.vdot(w, w).real
    return w, tau, d0, d1


def blocked_paired_qr(XL_int, rhs_int=None, NP=16):
    """In place on XL_int (n x h, pair-interleaved rows): on return column j holds the quaternion-upper-triangular R in rows
    0..2j+1 and w_j[2:] below.  rhs_int (n x k, same row order) <- Q^H rhs.  Returns the list of (j0, V, T) panels."""
    n, h = XL_int.shape
    panels = []
    for j0 in range(0, h, NP):
        npair = min(NP, h - j0)
        r0 = 2 * j0
        m = n - r0
        V = np.zeros((m, 2 * npair), dtype=complex)
        taus = np.zeros(2 * npair)
        # ---- panel factorization (the latency-bound part: npair sequential steps)
        for k in range(npair):
            j, t = j0 + k, 2 * (j0 + k)
            w, tau, d0, d1 = paired_reflector(XL_int[t:, j].copy())
            pw = phi_int(w)
            if k + 1 < npair:                          # remaining panel columns
                C = XL_int[t:, j + 1:j0 + npair]
                C -= tau * (np.outer(w, w.conj() @ C) + np.outer(pw, pw.conj() @ C))
            XL_int[t, j], XL_int[t + 1, j] = d0, d1
            XL_int[t + 2:, j] = w[2:]                   # w stored in the eliminated entries (w[0] = 1, w[1] = 0 implicit)
            V[t - r0:, 2 * k] = w
            V[t - r0:, 2 * k + 1] = pw
            taus[2 * k] = taus[2 * k + 1] = tau
        # ---- compact WY T (zlarft, forward / columnwise)
        nb = 2 * npair
        T = np.zeros((nb, nb), dtype=complex)
        for i in range(nb):
            T[i, i] = taus[i]
            if i > 0:
                T[:i, i] = -taus[i] * (T[:i, :i] @ (V[:, :i].conj().T @ V[:, i]))
        # ---- trailing update of the other left-half columns and of the right-hand sides: C <- (1 - V T V^H)^H C
        for C in ([XL_int[r0:, j0 + npair:]] + ([rhs_int[r0:, :]] if rhs_int is not None else [])):
            if C.shape[1]:
                C -= V @ (T.conj().T @ (V.conj().T @ C))
        panels.append((j0, V, T))
    return panels


def explicit_q_left(panels, n, h):
    """Left half of Q (pair-interleaved rows): Q applied to the quaternion unit vectors e_{2i}."""
    QL = np.zeros((n, h), dtype=complex)
    QL[2 * np.arange(h), np.arange(h)] = 1.0
    for j0, V, T in reversed(panels):
        r0 = 2 * j0
        C = QL[r0:, :]
        C -= V @ (T @ (V.conj().T @ C))
    return QL


def r_left(XL_int):
    """R (left half, pair-interleaved rows) with exact zeros below the quaternion diagonal."""
    n, h = XL_int.shape
    R = np.zeros_like(XL_int)
    for j in range(h):
        R[:2 * j + 2, j] = XL_int[:2 * j + 2, j]
    return R


def main():
    rs = np.random.RandomState(4)
    for h, span in ((24, 0), (40, 30)):
        n = 2 * h
        A = rs.randn(h, h) + 1j * rs.randn(h, h)
        B = rs.randn(h, h) + 1j * rs.randn(h, h)
        Dh = np.sort(np.logspace(span, -span, h))[::-1]
        X = np.block([[A, B], [-B.conj(), A.conj()]]) * np.concatenate([Dh, Dh])[None, :]
        p = interleave_rows(n)
        XL = X[p][:, :h].copy()
        rhs = (rs.randn(n, 5) + 1j * rs.randn(n, 5))
        rhs_w = rhs.copy()
        work = XL.copy()
        panels = blocked_paired_qr(work, rhs_w, NP=16)
        QL = explicit_q_left(panels, n, h)
        R = r_left(work)
        # full matrices (natural row order) for the checks
        inv = np.argsort(p)
        Qn = full_from_left(QL[inv])                              # columns: quaternion index i, then partners
        # R as a full symmetric matrix: rows in quaternion order (i -> natural rows i, i+h), left-half columns R_L
        Rn = full_from_left(R[inv])
        qh_rhs = (Qn.conj().T @ rhs[inv])[p]                     # rows (i, i+h) of Q^H rhs -> pair-interleaved (2i, 2i+1)
        print(f"n={n}, grading 1e+-{span}: |Q^H Q - 1| = {np.abs(Qn.conj().T @ Qn - np.eye(n)).max():.1e}; "
              f"|Q R - X| / |col| = {(np.abs(Qn @ Rn - X) / np.linalg.norm(X, axis=0)[None, :]).max():.1e}; "
              f"|Q^H rhs through the panels - explicit| = {np.abs(rhs_w - qh_rhs).max():.1e}")


if __name__ == "__main__":
    main()


# ---------------------------------------------------------------------------------------------------------------------
# Dataflow of the device panel kernel, emulated: like today's qr_panel_kernel (one exchange per column, everything derived
# from RAW dot products taken before the reflector is known), a pair-step needs per remaining column c only
#     D1_c = x^H c           = sum_q conj(x_e[q]) c_e[q] + conj(x_o[q]) c_o[q]          (all active quaternion rows q >= j)
#     D2_c = phi(x)^H c      = sum_q x_o[q] c_e[q] - x_e[q] c_o[q]
# plus the quaternion row j of the panel (x_e[j], x_o[j], c_e[j], c_o[j]); |x|^2 = D1_j.  With s = |x| / |q_j|,
# (a, b) = (1 + s) (x_e[j], x_o[j]), N = |a|^2 + |b|^2:
#     u^H c      = D1_c + s (conj(x_e[j]) c_e[j] + conj(x_o[j]) c_o[j])
#     phi(u)^H c = D2_c + s (x_o[j] c_e[j] - x_e[j] c_o[j])
#     d1 = w^H c      = (a u^H c + conj(b) phi(u)^H c) / N
#     d2 = phi(w)^H c = (conj(a) phi(u)^H c - b u^H c) / N
#     w_e[q] = (conj(a) x_e[q] + b conj(x_o[q])) / N,  w_o[q] = (conj(a) x_o[q] - b conj(x_e[q])) / N      (q > j; w[j] = (1, 0))
#     tau = 2 / |w|^2 = 2 N / |u|^2,   |u|^2 = 2 |x| (|x| + |q_j|)
#     c_e[q] -= tau (w_e[q] d1 + conj(w_o[q]) d2),   c_o[q] -= tau (w_o[q] d1 - conj(w_e[q]) d2)
# The same raw dots taken against the FINISHED columns c < j (which hold w_c below their diagonal) give the Gram entries the
# compact-WY recurrence needs, as in today's kernel:  w_c^H x = conj(D1_c),  w_c^H phi(x) = conj(D2_c)  (rows > j; add the row-j
# term), hence  w_c^H w_j = conj(w_c,e[j]) + (conj(a) conj(D1_c') + b conj(D2_c')) / N  and the phi-partners by
# phi(w_c)^H phi(w_j) = conj(w_c^H w_j),  phi(w_c)^H w_j = -conj(w_c^H phi(w_j)).
# ---------------------------------------------------------------------------------------------------------------------
def panel_rawdots(P):
    """P: m x nc panel (pair-interleaved rows, m >= 2 nc), factorized in place with the formulas above.  Returns taus."""
    m, nc = P.shape
    taus = np.zeros(nc)
    for j in range(nc):
        t = 2 * j
        xe, xo = P[t::2, j].copy(), P[t + 1::2, j].copy()           # active rows of the current column
        Ce, Co = P[t::2, j + 1:], P[t + 1::2, j + 1:]                # views of the remaining columns
        D1 = xe.conj() @ Ce + xo.conj() @ Co
        D2 = xo @ Ce - xe @ Co
        nx = np.sqrt((np.abs(xe) ** 2 + np.abs(xo) ** 2).sum())
        q = np.hypot(abs(xe[0]), abs(xo[0]))
        if nx == 0.0 or q == 0.0:
            raise NotImplementedError("degenerate column: handled by the general routine above")
        s = nx / q
        a, b = (1 + s) * xe[0], (1 + s) * xo[0]
        N = abs(a) ** 2 + abs(b) ** 2
        uc = D1 + s * (np.conj(xe[0]) * Ce[0] + np.conj(xo[0]) * Co[0])
        pc = D2 + s * (xo[0] * Ce[0] - xe[0] * Co[0])
        d1 = (a * uc + np.conj(b) * pc) / N
        d2 = (np.conj(a) * pc - b * uc) / N
        we = (np.conj(a) * xe + b * np.conj(xo)) / N
        wo = (np.conj(a) * xo - b * np.conj(xe)) / N
        we[0], wo[0] = 1.0, 0.0
        tau = 2.0 * N / (2.0 * nx * (nx + q))
        Ce -= tau * (np.outer(we, d1) + np.outer(np.conj(wo), d2))
        Co -= tau * (np.outer(wo, d1) - np.outer(np.conj(we), d2))
        P[t, j], P[t + 1, j] = -s * xe[0], -s * xo[0]
        P[t + 2::2, j], P[t + 3::2, j] = we[1:], wo[1:]
        taus[j] = tau
    return taus


def check_rawdots():
    rs = np.random.RandomState(7)
    m, nc = 96, 16
    P = (rs.randn(m, nc) + 1j * rs.randn(m, nc)) * np.logspace(8, -8, nc)[None, :]
    ref = P.copy()
    blocked_paired_qr(ref, NP=nc)                                   # only the first panel matters: h = nc columns
    got = P.copy()
    panel_rawdots(got)
    print(f"panel from raw dot products vs reference panel: max rel diff {np.abs(got - ref).max() / np.abs(ref).max():.1e} "
          f"(column-wise {np.max(np.abs(got - ref).max(axis=0) / np.abs(ref).max(axis=0)):.1e})")


if __name__ == "__main__":
    check_rawdots()
