"""Test infrastructure (oracle side): executable specification of the DEVICE paired-Householder panel kernel
(`qr_panel_paired_kernel` in `dqmc_b200/csrc/qr.cu`), formula by formula, in the conventions the CUDA code uses.  Not used by the product.

Conventions (they differ from quaternion_qr_blocked.py only by signs / normalisation):

* a symmetric matrix  S = [[A, B], [-conj(B), conj(A)]]  is held as its left half with PAIR-INTERLEAVED rows
  (rows 2i, 2i+1 = natural rows i, i+h); the partner of a column x is  psi(x)[2i] = -conj(x[2i+1]),  psi(x)[2i+1] = conj(x[2i])
  (= column c+h of S);  psi(x)^H y = sum_i x_e[i] y_o[i] - x_o[i] y_e[i]  is bilinear and antisymmetric;
* pair-step j:  u = x with the quaternion entry j replaced by (1 + |x|/|x_j|) x_j;  v = u / |x|  (real scale, v = O(1));
  H_j = 1 - tau_j (v v^H + psi(v) psi(v)^H),  tau_j = |x| / (|x| + |x_j|);  H_j x = x_j - u_j  on the row pair j, zero below;
* everything a step needs comes from RAW dot products over the rows strictly below row pair j
      D1_c = sum_{q>j} conj(x_e) c_e + conj(x_o) c_o,      D2_c = sum_{q>j} x_e c_o - x_o c_e
  and from row pair j of the panel (what the cluster exchanges once per step);
* a panel of 16 pair-steps = 32 reflectors  V = [v_0, psi(v_0), v_1, psi(v_1), ...]  (explicit, block lower trapezoidal with
  2x2 diagonal blocks), compact-WY  H_0 ... H_15 = 1 - V T V^H,  T from the Gram entries the raw dots of the FINISHED columns
  give for free:   v_c^H v_j = r_c r_j conj(uc_c),  psi(v_c)^H v_j = -r_c r_j pc_c,  v_c^H psi(v_j) = r_c r_j conj(pc_c),
  psi(v_c)^H psi(v_j) = r_c r_j uc_c   (uc_c = u_j^H [column c],  pc_c = psi(u_j)^H [column c],  r = 1/|x|).

The kernel evaluates |x|, 1/|x|, |x_j| and 1/(|x| + |x_j|) with reciprocal square roots / reciprocals refined by two Newton steps
(1-2 ulp) instead of sqrt and divisions; the algebra below is the same, the results agree to rounding (tests/test_gpu_paired.py).

    python -m oracle.experiments.paired_panel_spec
"""
import numpy as np


def psi(x):
    w = np.empty_like(x)
    w[0::2] = -np.conj(x[1::2])
    w[1::2] = np.conj(x[0::2])
    return w


def full_from_left_int(XL):
    """n x h (pair-interleaved rows) -> n x n with pair-interleaved rows AND columns (column 2c = left column c, 2c+1 = partner)."""
    n, h = XL.shape
    F = np.zeros((n, n), dtype=complex)
    F[:, 0::2] = XL
    for c in range(h):
        F[:, 2 * c + 1] = psi(XL[:, c])
    return F


def panel_device_dataflow(P):
    """P: m x np panel (pair-interleaved rows, m >= 2 np).  Returns (R_panel, V (m x 2np), T (2np x 2np), dabs)."""
    P = P.copy()
    m, npair = P.shape
    V = np.zeros((m, 2 * npair), dtype=complex)
    T = np.zeros((2 * npair, 2 * npair), dtype=complex)
    dabs = np.zeros(npair)
    rn = np.zeros(npair)                      # 1/|x| of the finished columns (the deferred real scale `sc`)
    E, O = P[0::2, :], P[1::2, :]             # views: even / odd components
    for j in range(npair):
        xe, xo = E[:, j].copy(), O[:, j].copy()
        xe[:j + 1] = 0.0; xo[:j + 1] = 0.0    # the column buffer is zero on rows <= j
        # raw dots of the current column with ALL panel columns (finished ones hold u below their diagonal)
        D1 = xe.conj() @ E + xo.conj() @ O
        D2 = xe @ O - xo @ E
        xe0, xo0 = E[j, j], O[j, j]
        q2 = abs(xe0) ** 2 + abs(xo0) ** 2
        nrm2 = q2 + D1[j].real
        if nrm2 == 0.0:
            dabs[j] = 0.0; rn[j] = 0.0
            continue                          # tau = 0, nothing to do (T row/column stay zero)
        nx, q = np.sqrt(nrm2), np.sqrt(q2)
        f = 1.0 / (nx * (nx + q))
        if q > 0.0:
            s1 = (q + nx) / q
            ue0, uo0 = s1 * xe0, s1 * xo0
        else:
            ue0, uo0 = nx, 0.0
        uc = D1 + np.conj(ue0) * E[j, :] + np.conj(uo0) * O[j, :]      # u^H c      (for every panel column)
        pc = D2 + ue0 * O[j, :] - uo0 * E[j, :]                        # psi(u)^H c
        rnj = (nx + q) * f                                            # 1 / nx
        tau = nx * nx * f                                             # nx / (nx + q)
        # ---- Gram entries against the finished columns -> two columns of T
        g = np.zeros(2 * npair, dtype=complex); gp = np.zeros(2 * npair, dtype=complex)
        for c in range(j):
            rr = rn[c] * rnj
            g[2 * c], g[2 * c + 1] = rr * np.conj(uc[c]), -rr * pc[c]
            gp[2 * c], gp[2 * c + 1] = rr * np.conj(pc[c]), rr * uc[c]
        i0 = 2 * j
        T[:i0, i0] = -tau * (T[:i0, :i0] @ g[:i0]); T[i0, i0] = tau
        T[:i0 + 1, i0 + 1] = -tau * (T[:i0 + 1, :i0 + 1] @ gp[:i0 + 1]); T[i0 + 1, i0 + 1] = tau
        # ---- update of the remaining columns c > j:  c += A alpha + conj(B) beta  (rows > j), row j with u_j
        alpha = -f * uc
        for c in range(j + 1, npair):
            be = f * pc[c]                    # e-lanes: +f pc, o-lanes: -f pc
            E[j + 1:, c] += xe[j + 1:] * alpha[c] + np.conj(xo[j + 1:]) * be
            O[j + 1:, c] += xo[j + 1:] * alpha[c] + np.conj(xe[j + 1:]) * (-be)
            ej, oj = E[j, c], O[j, c]
            E[j, c] = ej + ue0 * alpha[c] + np.conj(uo0) * be
            O[j, c] = oj + uo0 * alpha[c] + np.conj(ue0) * (-be)
        # ---- column j: R entries on the row pair j; v below
        V[2 * j, 2 * j], V[2 * j + 1, 2 * j] = ue0 * rnj, uo0 * rnj
        V[2 * j + 2::2, 2 * j] = E[j + 1:, j] * rnj
        V[2 * j + 3::2, 2 * j] = O[j + 1:, j] * rnj
        V[:, 2 * j + 1] = psi(V[:, 2 * j])
        E[j, j], O[j, j] = xe0 - ue0, xo0 - uo0
        rn[j] = rnj
        dabs[j] = nx
    R = P.copy()
    for j in range(npair):
        R[2 * j + 2:, j] = 0.0
    return R, V, T, dabs


def paired_qr_device(XL, rhs=None, NP=16):
    """Blocked paired QR in the device's conventions.  XL: n x h (interleaved rows).  Returns (R_L, rhs <- Q^H rhs, dabs)."""
    XL = XL.copy()
    n, h = XL.shape
    rhs = None if rhs is None else rhs.copy()
    dabs = np.zeros(h)
    for j0 in range(0, h, NP):
        npair = min(NP, h - j0)
        r0 = 2 * j0
        R, V, T, d = panel_device_dataflow(XL[r0:, j0:j0 + npair])
        XL[r0:, j0:j0 + npair] = R
        dabs[j0:j0 + npair] = d
        for C in ([XL[r0:, j0 + npair:]] + ([rhs[r0:, :]] if rhs is not None else [])):
            if C.shape[1]:
                C -= V @ (T.conj().T @ (V.conj().T @ C))        # larfb with conjT = 1
    return XL, rhs, dabs


def main():
    rs = np.random.RandomState(11)
    for h, span in ((16, 0), (48, 0), (64, 40)):
        n = 2 * h
        XL = (rs.randn(n, h) + 1j * rs.randn(n, h)) * np.sort(np.logspace(span, -span, h))[::-1][None, :]
        # single panel: 1 - V T V^H is unitary, commutes with psi, and maps the panel onto its R
        R, V, T, d = panel_device_dataflow(XL[:, :16])
        Qp = np.eye(n) - V @ T @ V.conj().T
        print(f"n={n}: panel |Q^H Q - 1| = {np.abs(Qp.conj().T @ Qp - np.eye(n)).max():.1e}, "
              f"|Q^H P - R| / |col| = {(np.abs(Qp.conj().T @ XL[:, :16] - R) / np.linalg.norm(XL[:, :16], axis=0)).max():.1e}, "
              f"T quaternion structure {max(np.abs(T[1::2, 1::2] - T[0::2, 0::2].conj()).max(), np.abs(T[0::2, 1::2] + T[1::2, 0::2].conj()).max()):.1e}")
        rhs = np.zeros((n, h), dtype=complex); rhs[2 * np.arange(h), np.arange(h)] = 1.0
        RL, QHL, dabs = paired_qr_device(XL, rhs)
        QH = full_from_left_int(QHL)                      # Q^H (both interleaved)
        Rf = full_from_left_int(RL)
        Xf = full_from_left_int(XL)
        print(f"      blocked: |Q Q^H - 1| = {np.abs(QH @ QH.conj().T - np.eye(n)).max():.1e}, |Q R - X| / |col| = "
              f"{(np.abs(QH.conj().T @ Rf - Xf) / np.linalg.norm(Xf, axis=0)[None, :]).max():.1e}, "
              f"dabs vs |R_jj| {np.abs(dabs - np.sqrt(np.abs(RL[2 * np.arange(h), np.arange(h)]) ** 2 + np.abs(RL[2 * np.arange(h) + 1, np.arange(h)]) ** 2)).max() / dabs.max():.1e}")


if __name__ == "__main__":
    main()
