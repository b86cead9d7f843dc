"""The hot path of the reference, restated in NumPy/SciPy.  Test infrastructure.

One class, ``OracleDQMC``, holding the state the reference keeps in ``mc.p`` / ``mc.s``
(hsfield, greens, UDT stack, current_slice, direction) and one method per reference
function on the path.  Slices and sites are 0-based here: reference slice
``l`` is ``l-1``; the artificial slices 0 and M+1 of ``propagate`` are -1 and M.

Random numbers are *injected*: ``local_updates`` takes any object with a ``rand()`` method
(``JuliaMT`` for fixture parity, ``UniformStream`` for a shared pre-generated stream).
"""
import numpy as np
import scipy.linalg as sla
from scipy.linalg import blas
import scipy.sparse as sp

from .model import build_model


class UniformStream:
    """A pre-generated uniform [0,1) stream consumed in order (the C-ABI's ``u`` argument)."""

    def __init__(self, u):
        self.u = np.asarray(u, dtype=np.float64)
        self.pos = 0

    def rand(self):
        v = self.u[self.pos]
        self.pos += 1
        return v


def decompose_udt(A):
    """linalg.jl:20-39 ``decompose_udt!``: column-pivoted QR (zgeqp3), D=|Re R_ii|, T=D^-1 R P^T."""
    Q, R, piv = sla.qr(A, mode="full", pivoting=True, check_finite=False)
    D = np.abs(np.real(np.diag(R)))
    T = np.empty_like(R)
    T[:, piv] = R / D[:, None]
    return Q, D, T


def _apply_left(F, M):
    return F @ M


class OracleDQMC:
    def __init__(self, p, l=None, dense_hoppings=False):
        self.p = p
        self.l = l if l is not None else build_model(p, dense_hoppings=dense_hoppings)
        self.dense_B = False         # True: CBFalse slice matrices (slice_matrices.jl:25-94), needs dense_hoppings
        self.N = self.l.sites
        self.n = p.flv * self.N
        self.hsfield = None          # [opdim, N, M] (Julia index order)
        self.boson_action = 0.0
        self.greens = None
        self.log_det = 0.0
        self.current_slice = 0
        self.direction = -1
        self.n_uniforms = 0
        assert p.slices % p.safe_mult == 0
        self.n_elements = p.slices // p.safe_mult + 1
        self.ranges = [range(i * p.safe_mult, (i + 1) * p.safe_mult) for i in range(self.n_elements - 1)]

    # ---------------------------------------------------------------- action.jl
    def calc_boson_action(self, hsfield=None):
        """action.jl:1-52."""
        p, l = self.p, self.l
        hs = self.hsfield if hsfield is None else hsfield
        S = 0.0
        if not p.edrun:
            h = hs.reshape(p.opdim, l.L, l.L, p.slices, order="F")
            t = h - np.roll(h, -1, axis=3)
            S += 0.5 / p.delta_tau * 1 / p.c ** 2 * np.sum(t * t)
            for ax in (1, 2):
                t = h - np.roll(h, -1, axis=ax)
                S += p.delta_tau * 0.5 * np.sum(t * t)
            sq = np.sum(hs * hs, axis=0)
            S += np.sum(p.delta_tau * p.r / 2.0 * sq)
            S += np.sum(p.delta_tau * p.u / 4.0 * sq * sq)
        else:
            sq = np.sum(hs * hs, axis=0)
            S += np.sum(p.delta_tau * p.r / 2.0 * sq)
        return float(S)

    def calc_boson_action_diff(self, site, slc, new_op):
        """action.jl:57-101."""
        p, l = self.p, self.l
        hs = self.hsfield
        old_op = hs[:, site, slc]
        diff = new_op - old_op
        old_sq = float(np.dot(old_op, old_op))
        new_sq = float(np.dot(new_op, new_op))
        sq_diff = new_sq - old_sq
        pow4_diff = new_sq * new_sq - old_sq * old_sq
        op_earlier = hs[:, site, l.time_neighbors[1, slc]]
        op_later = hs[:, site, l.time_neighbors[0, slc]]
        op_time = op_later + op_earlier
        op_space = np.zeros(p.opdim)
        for nb in range(4):
            op_space = op_space + hs[:, l.neighbors[nb, site], slc]
        dS = 0.0
        if not p.edrun:
            dS += 1.0 / (p.delta_tau * p.c ** 2) * (sq_diff - float(np.dot(op_time, diff)))
            dS += 0.5 * p.delta_tau * (4 * sq_diff - 2.0 * float(np.dot(op_space, diff)))
            dS += p.delta_tau * (0.5 * p.r * sq_diff + 0.25 * p.u * pow4_diff)
        else:
            dS += p.delta_tau * (0.5 * p.r * sq_diff)
        return dS

    # ---------------------------------------------------------------- interactions.jl
    def interaction_matrix_exp(self, slc, power=1.0):
        """interactions.jl:35-88: sparse n x n e^{-power dtau V(slice)} (10 diagonal N-blocks, 3 nnz per row)."""
        p, N = self.p, self.N
        assert p.opdim == 3
        hs = self.hsfield[:, :, slc]
        nrm = np.sqrt(np.sum(hs * hs, axis=0))
        sh = np.sinh(p.lam * p.delta_tau * nrm) / nrm
        C = np.cosh(p.lam * p.delta_tau * nrm).astype(complex)
        S = (1j * hs[1] - hs[0]) * power * sh
        R = (-hs[2]) * power * sh + 0j
        idx = np.arange(N)
        blocks = [(0, 0, C), (0, 1, S), (1, 0, np.conj(S)), (1, 1, C), (0, 3, R), (1, 2, -R), (2, 1, -R), (2, 2, C),
                  (2, 3, np.conj(S)), (3, 0, R), (3, 2, S), (3, 3, C)]
        rows = np.concatenate([r * N + idx for r, _, _ in blocks])
        cols = np.concatenate([c * N + idx for _, c, _ in blocks])
        vals = np.concatenate([v for _, _, v in blocks])
        return sp.csc_matrix((vals, (rows, cols)), shape=(self.n, self.n))

    def interaction_matrix_exp_op(self, op, power=1.0):
        """interactions.jl:102-141 (4x4, O(3))."""
        p = self.p
        nrm = np.sqrt(op[0] * op[0] + op[1] * op[1] + op[2] * op[2])
        sh = power * np.sinh(p.lam * p.delta_tau * nrm) / nrm
        C = np.cosh(p.lam * p.delta_tau * nrm)
        S = (1j * op[1] - op[0]) * sh
        R = (-op[2]) * sh
        cS = np.conj(S)
        return np.array([[C, S, 0, R],
                         [cS, C, -R, 0],
                         [0, -R, C, cS],
                         [R, 0, S, C]], dtype=complex)

    # ---------------------------------------------------------------- slice_matrices.jl (no checkerboard)
    def _dense_slice_matrix(self, slc, power):
        """slice_matrices.jl:25-43 (CBFalse): e^{-dtau T/2} e^{-dtau T/2} e^{-dtau V} or its inverse."""
        l = self.l
        eV = self.interaction_matrix_exp(slc, power).toarray()
        if power > 0:
            return l.hopping_matrix_exp @ (l.hopping_matrix_exp @ eV)
        return eV @ (l.hopping_matrix_exp_inv @ l.hopping_matrix_exp_inv)

    # ---------------------------------------------------------------- slice_matrices.jl (CBAssaad)
    def multiply_B_left(self, slc, M):
        """slice_matrices.jl:101-129: M <- hopB½ hopA hopB½ e^{dtau mu} e^{-dtau V} M."""
        l = self.l
        if self.dense_B:
            return self._dense_slice_matrix(slc, 1.0) @ M
        M = self.interaction_matrix_exp(slc, 1.0) @ M
        M = l.chkr_mu @ M
        M = l.chkr_hop_half[1] @ M
        M = l.chkr_hop[0] @ M
        M = l.chkr_hop_half[1] @ M
        return M

    def multiply_B_right(self, slc, M):
        """slice_matrices.jl:131-153: M <- M hopB½ hopA hopB½ e^{dtau mu} e^{-dtau V}."""
        l = self.l
        if self.dense_B:
            return M @ self._dense_slice_matrix(slc, 1.0)
        eV = self.interaction_matrix_exp(slc, 1.0)
        M = (l.chkr_hop_half[1].T @ M.T).T
        M = (l.chkr_hop[0].T @ M.T).T
        M = (l.chkr_hop_half[1].T @ M.T).T
        M = (l.chkr_mu.T @ M.T).T
        M = (eV.T @ M.T).T
        return M

    def multiply_B_inv_left(self, slc, M):
        """slice_matrices.jl:155-177: M <- e^{+dtau V} mu^-1 hopB½^-1 hopA^-1 hopB½^-1 M."""
        l = self.l
        if self.dense_B:
            return self._dense_slice_matrix(slc, -1.0) @ M
        eV = self.interaction_matrix_exp(slc, -1.0)
        M = l.chkr_hop_half_inv[1] @ M
        M = l.chkr_hop_inv[0] @ M
        M = l.chkr_hop_half_inv[1] @ M
        M = l.chkr_mu_inv @ M
        M = eV @ M
        return M

    def multiply_B_inv_right(self, slc, M):
        """slice_matrices.jl:179-201: M <- M e^{+dtau V} mu^-1 hopB½^-1 hopA^-1 hopB½^-1."""
        l = self.l
        if self.dense_B:
            return M @ self._dense_slice_matrix(slc, -1.0)
        eV = self.interaction_matrix_exp(slc, -1.0)
        M = (eV.T @ M.T).T
        M = (l.chkr_mu_inv.T @ M.T).T
        M = (l.chkr_hop_half_inv[1].T @ M.T).T
        M = (l.chkr_hop_inv[0].T @ M.T).T
        M = (l.chkr_hop_half_inv[1].T @ M.T).T
        return M

    def multiply_daggered_B_left(self, slc, M):
        """slice_matrices.jl:203-226: M <- B(slice)^dagger M."""
        l = self.l
        if self.dense_B:
            return self._dense_slice_matrix(slc, 1.0).conj().T @ M
        eV = self.interaction_matrix_exp(slc, 1.0)
        M = l.chkr_hop_half_dagger[1] @ M
        M = l.chkr_hop_dagger[0] @ M
        M = l.chkr_hop_half_dagger[1] @ M
        M = l.chkr_mu @ M
        M = eV @ M
        return M

    def slice_matrix(self, slc, power=1.0):
        """slice_matrices.jl:4-17."""
        I = np.eye(self.n, dtype=complex)
        return self.multiply_B_left(slc, I) if power > 0 else self.multiply_B_inv_left(slc, I)

    # ---------------------------------------------------------------- stack.jl
    def initialize_stack(self):
        """stack.jl:183-242 (only the buffers that carry state)."""
        n, ne = self.n, self.n_elements
        self.u_stack = np.zeros((ne, n, n), dtype=complex)
        self.d_stack = np.zeros((ne, n))
        self.t_stack = np.zeros((ne, n, n), dtype=complex)
        self.greens = np.zeros((n, n), dtype=complex)
        eye = np.eye(n, dtype=complex)
        self.Ul, self.Ur, self.Tl, self.Tr = eye.copy(), eye.copy(), eye.copy(), eye.copy()
        self.Dl, self.Dr = np.ones(n), np.ones(n)

    def build_stack(self):
        """stack.jl:251-272."""
        n = self.n
        self.u_stack[0] = np.eye(n)
        self.d_stack[0] = 1.0
        self.t_stack[0] = np.eye(n)
        for i in range(len(self.ranges)):
            self.add_slice_sequence_left(i)
        self.current_slice = self.p.slices   # == reference p.slices + 1
        self.direction = -1

    def add_slice_sequence_left(self, idx):
        """stack.jl:278-292: slab idx+1 <- UDT( B(hi)...B(lo) U_idx D_idx ), T_{idx+1} = T T_idx."""
        U = self.u_stack[idx].copy()
        for slc in self.ranges[idx]:
            U = self.multiply_B_left(slc, U)
        U = U * self.d_stack[idx][None, :]
        Q, D, T = decompose_udt(U)
        self.u_stack[idx + 1], self.d_stack[idx + 1] = Q, D
        self.t_stack[idx + 1] = T @ self.t_stack[idx]

    def add_slice_sequence_right(self, idx):
        """stack.jl:298-313: slab idx <- UDT( B(lo)^†...B(hi)^† U_{idx+1} D_{idx+1} ), T_idx = T T_{idx+1}."""
        U = self.u_stack[idx + 1].copy()
        for slc in reversed(self.ranges[idx]):
            U = self.multiply_daggered_B_left(slc, U)
        U = U * self.d_stack[idx + 1][None, :]
        Q, D, T = decompose_udt(U)
        self.u_stack[idx], self.d_stack[idx] = Q, D
        self.t_stack[idx] = T @ self.t_stack[idx + 1]

    def wrap_greens(self, gf, curr_slice, direction):
        """stack.jl:316-325.  ``curr_slice`` 0-based."""
        if direction == -1:
            gf = self.multiply_B_inv_left(curr_slice - 1, gf)
            gf = self.multiply_B_right(curr_slice - 1, gf)
        else:
            gf = self.multiply_B_left(curr_slice, gf)
            gf = self.multiply_B_inv_right(curr_slice, gf)
        return gf

    def calculate_greens(self):
        """stack.jl:338-369: G = [1 + Ul Dl Tl (Ur Dr Tr)^†]^-1 through two UDTs; keeps U,d,T for logdet."""
        tmp = self.Tl @ self.Tr.conj().T
        tmp = tmp * self.Dr[None, :]
        tmp = self.Dl[:, None] * tmp
        U, D, T = decompose_udt(tmp)
        U = self.Ul @ U
        tmp2 = T @ self.Ur.conj().T
        # myrdiv!(tmp, U', tmp2): tmp = U' / tmp2  (linalg.jl:61)
        tmp = sla.solve(tmp2.conj().T, U, check_finite=False).conj().T
        tmp[np.diag_indices_from(tmp)] += D
        u, d, t = decompose_udt(tmp)
        tmp = t @ tmp2
        Tn = sla.inv(tmp, check_finite=False)
        Un = (U @ u).conj().T
        d = 1.0 / d
        self.U, self.d, self.T = Un, d, Tn
        self.greens = Tn @ (d[:, None] * Un)
        return self.greens

    def calculate_logdet(self):
        """stack.jl:377-385."""
        ld = np.linalg.slogdet(self.U)
        lt = np.linalg.slogdet(self.T)
        self.log_det = float(ld[1] + np.sum(np.log(self.d)) + lt[1])
        return self.log_det

    def propagate(self):
        """stack.jl:391-499, 0-based: reference current_slice c is c-1 here (0 -> -1, M+1 -> M)."""
        p = self.p
        M, sm = p.slices, p.safe_mult
        n = self.n
        eye = np.eye(n, dtype=complex)
        cs1 = self.current_slice + 1  # reference (1-based) value
        if self.direction == 1:
            if cs1 % sm == 0:
                cs1 += 1
                self.current_slice = cs1 - 1
                if cs1 == 1:
                    self.Ur, self.Dr, self.Tr = self.u_stack[0].copy(), self.d_stack[0].copy(), self.t_stack[0].copy()
                    self.u_stack[0] = eye; self.d_stack[0] = 1.0; self.t_stack[0] = eye
                    self.Ul, self.Dl, self.Tl = eye.copy(), np.ones(n), eye.copy()
                    self.calculate_greens()
                    self.calculate_logdet()
                elif 1 < cs1 <= M:
                    idx = (cs1 - 1) // sm - 1   # 0-based index of reference idx
                    self.Ur, self.Dr, self.Tr = (self.u_stack[idx + 1].copy(), self.d_stack[idx + 1].copy(),
                                                 self.t_stack[idx + 1].copy())
                    self.add_slice_sequence_left(idx)
                    self.Ul, self.Dl, self.Tl = (self.u_stack[idx + 1].copy(), self.d_stack[idx + 1].copy(),
                                                 self.t_stack[idx + 1].copy())
                    if p.all_checks:
                        gt = self.wrap_greens(self.greens.copy(), self.current_slice - 1, 1)
                    self.calculate_greens()
                    if p.all_checks:
                        self.last_check = float(np.max(np.abs(gt - self.greens)))
                else:
                    idx = self.n_elements - 2
                    self.add_slice_sequence_left(idx)
                    self.direction = -1
                    self.current_slice = M
                    self.propagate()
            else:
                self.greens = self.wrap_greens(self.greens, self.current_slice, 1)
                self.current_slice += 1
        else:
            if (cs1 - 1) % sm == 0:
                cs1 -= 1
                self.current_slice = cs1 - 1
                if cs1 == M:
                    self.Ul, self.Dl, self.Tl = self.u_stack[-1].copy(), self.d_stack[-1].copy(), self.t_stack[-1].copy()
                    self.u_stack[-1] = eye; self.d_stack[-1] = 1.0; self.t_stack[-1] = eye
                    self.Ur, self.Dr, self.Tr = eye.copy(), np.ones(n), eye.copy()
                    self.calculate_greens()
                    self.calculate_logdet()
                    self.greens = self.wrap_greens(self.greens, self.current_slice + 1, -1)
                elif 0 < cs1 < M:
                    idx = cs1 // sm      # 0-based index of reference idx = cs1/sm + 1
                    self.Ul, self.Dl, self.Tl = (self.u_stack[idx].copy(), self.d_stack[idx].copy(),
                                                 self.t_stack[idx].copy())
                    self.add_slice_sequence_right(idx)
                    self.Ur, self.Dr, self.Tr = (self.u_stack[idx].copy(), self.d_stack[idx].copy(),
                                                 self.t_stack[idx].copy())
                    gt = self.greens.copy() if p.all_checks else None
                    self.calculate_greens()
                    if p.all_checks:
                        self.last_check = float(np.max(np.abs(gt - self.greens)))
                    self.greens = self.wrap_greens(self.greens, self.current_slice + 1, -1)
                else:
                    self.add_slice_sequence_right(0)
                    self.direction = 1
                    self.current_slice = -1
                    self.propagate()
            else:
                self.greens = self.wrap_greens(self.greens, self.current_slice, -1)
                self.current_slice -= 1

    # ---------------------------------------------------------------- dqmc_framework.jl
    def init(self, start_conf):
        """dqmc_framework.jl:157-177 ``init!(mc, start_conf)``."""
        self.hsfield = np.array(start_conf, dtype=np.float64, copy=True)
        self.boson_action = self.calc_boson_action()
        self.initialize_stack()
        self.build_stack()
        self.propagate()

    # ---------------------------------------------------------------- local_updates.jl
    def calc_detratio(self, i, new_op):
        """local_updates.jl:42-59; leaves delta_i and M for update_greens."""
        N = self.N
        slc = self.current_slice
        eV1 = self.interaction_matrix_exp_op(self.hsfield[:, i, slc], -1.0)
        eV2 = self.interaction_matrix_exp_op(new_op, 1.0)
        self.delta_i = eV1 @ eV2 - np.eye(4)
        Mtmp = np.eye(4) - self.greens[i::N, i::N]
        self.Mmat = np.eye(4) + self.delta_i @ Mtmp
        return complex(np.linalg.det(self.Mmat))

    def update_greens(self, i):
        """local_updates.jl:61-95: G += (G[:,i::N]-E_i) M^-1 . delta_i G[i::N,:]  (in-place zgemm like mul!/axpy)."""
        N = self.N
        if not self.greens.flags.f_contiguous:
            self.greens = np.asfortranarray(self.greens)
        g = self.greens
        A = g[:, i::N].copy()
        for k in range(4):
            A[i + k * N, k] -= 1.0
        A = A @ np.linalg.inv(self.Mmat)
        B = self.delta_i @ g[i::N, :]
        out = blas.zgemm(1.0, A, B, beta=1.0, c=g, overwrite_c=1)
        assert out is g or np.shares_memory(out, g)

    def local_updates(self, rng):
        """local_updates.jl:1-39.  Returns the acceptance fraction."""
        p = self.p
        slc = self.current_slice
        acc = 0
        for i in range(self.N):
            # randuniform(box, opdim): dqmc_framework.jl:628-635
            new_op = self.hsfield[:, i, slc] + np.array([-p.box + 2 * p.box * rng.rand() for _ in range(p.opdim)])
            e_dS = np.exp(-self.calc_boson_action_diff(i, slc, new_op))
            detratio = self.calc_detratio(i, new_op)
            p_acc = e_dS * detratio.real
            if p_acc > 1.0 or rng.rand() < p_acc:
                acc += 1
                self.hsfield[:, i, slc] = new_op
                self.boson_action += -np.log(e_dS)
                self.update_greens(i)
        return acc / self.N

    # ---------------------------------------------------------------- boson_measurements.jl
    def measure_chi_dynamic(self, conf=None):
        """boson_measurements.jl:6-10, 48-56: dtau/(N M) * sum_k |rfft(phi_k)|^2 restricted to [:, 1:n, 1:nt]."""
        p = self.p
        conf = self.hsfield if conf is None else conf
        opdim, sites, slices = conf.shape
        L = int(round(np.sqrt(sites)))
        h = conf.reshape(opdim, L, L, slices, order="F")
        ft = np.fft.fftn(h, axes=(1, 2, 3))
        n, nt = L // 2 + 1, slices // 2 + 1
        C = np.sum(np.abs(ft) ** 2, axis=0)[:n, :n, :nt]
        return p.delta_tau / (sites * slices) * C

    # ---------------------------------------------------------------- global_updates.jl
    def global_update(self, rng):
        """global_updates.jl:18-59 (backup by copy instead of pointer swap).  Returns 0/1."""
        p = self.p
        assert self.current_slice == p.slices - 1 and self.direction == -1
        S_old = self.boson_action
        bk = (self.u_stack.copy(), self.d_stack.copy(), self.t_stack.copy(), self.greens.copy(), self.log_det, self.hsfield.copy())
        for k in range(p.opdim):   # global_update_perform_shift! (:10-16)
            self.hsfield[k] += -p.box_global + 2 * p.box_global * rng.rand()
        self.build_stack()
        self.propagate()
        self.boson_action = self.calc_boson_action()
        p_acc = np.exp(-(self.boson_action - S_old)) * np.exp(bk[4] - self.log_det)
        self.last_global_p_acc = float(p_acc)
        if p_acc > 1.0 or rng.rand() < p_acc:
            return 1
        self.boson_action = S_old
        self.u_stack, self.d_stack, self.t_stack, self.greens, self.log_det, self.hsfield = bk
        return 0

    # ---------------------------------------------------------------- time-displaced Green's functions
    def effective_greens2greens(self, G):
        """fermion_measurements.jl:1125-1142 (CBTrue): G <- hop½_inv[1] hop½_inv[2] G hop½[2] hop½[1]
        (the loops run over reverse(1:n_groups)); :1178-1185 for the dense (CBFalse) hoppings."""
        l = self.l
        if self.dense_B:
            return l.hopping_matrix_exp_inv @ (G @ l.hopping_matrix_exp)
        for i in reversed(range(len(l.chkr_hop_half))):
            G = (l.chkr_hop_half[i].T @ G.T).T
        for i in reversed(range(len(l.chkr_hop_half_inv))):
            G = l.chkr_hop_half_inv[i] @ G
        return G

    def calc_Bchain_udts(self, invert=False, left=True):
        """fermion_measurements.jl:1434-1503 ``calc_Bchain_udts!``: UDTs at the safe_mult slices of
        left,  invert=False:  B(tau,1)      = B(tau) ... B(1)            (udt[i]: slices 1..ranges[i][end])
        left,  invert=True :  B(tau,1)^-1   = B(1)^-1 ... B(tau)^-1
        right, invert=False:  B(beta,tau)   = B(beta) ... B(tau)         (udt[i]: slices ranges[i][1]..M)
        right, invert=True :  B(beta,tau)^-1 = B(tau)^-1 ... B(beta)^-1
        Returns (u_stack, d_stack, t_stack) with K = M/safe_mult entries, already reversed for dir = RIGHT."""
        n, K = self.n, len(self.ranges)
        us, ds, ts = [None] * K, [None] * K, [None] * K
        rightmult = (not left and not invert) or (left and invert)
        rng = range(K) if left else reversed(range(K))
        for i, ridx in enumerate(rng):
            if i == 0:
                cur = np.eye(n, dtype=complex)
            else:
                cur = (ts[i - 1] if rightmult else us[i - 1]).copy()
            srange = self.ranges[ridx] if left else reversed(self.ranges[ridx])
            for slc in srange:
                if not invert:
                    cur = self.multiply_B_left(slc, cur) if left else self.multiply_B_right(slc, cur)
                else:
                    cur = self.multiply_B_inv_right(slc, cur) if left else self.multiply_B_inv_left(slc, cur)
            if i != 0:
                cur = ds[i - 1][:, None] * cur if rightmult else cur * ds[i - 1][None, :]
            U, D, T = decompose_udt(cur)
            ds[i] = D
            if not rightmult:
                us[i] = U
                ts[i] = T if i == 0 else T @ ts[i - 1]
            else:
                ts[i] = T
                us[i] = U if i == 0 else us[i - 1] @ U
        if not left:
            us.reverse(); ds.reverse(); ts.reverse()
        return us, ds, ts

    @staticmethod
    def inv_one_plus_udt_scalettar(U, D, T):
        """linalg.jl:302-331: [1 + U D T]^-1 with scales above and below one separated and two intermediate UDTs."""
        Dpinv = 1.0 / np.maximum(D, 1.0)
        Dm = np.minimum(D, 1.0)
        l = sla.solve(T, np.diag(Dpinv).astype(complex), check_finite=False)
        r = U * Dm[None, :] + l
        u, d, t = decompose_udt(r)
        r = sla.solve(t, np.diag(1.0 / d).astype(complex), check_finite=False)
        l = Dpinv[:, None] * (r @ u.conj().T)
        u, d, t = decompose_udt(l)
        l = sla.solve(T, u, check_finite=False)
        return (l * d[None, :]) @ t

    @staticmethod
    def inv_sum_udts_scalettar(Ua, Da, Ta, Ub, Db, Tb):
        """linalg.jl:512-567: [Ua Da Ta + Ub Db Tb]^-1, same scale separation."""
        Dap, Dam = np.maximum(Da, 1.0), np.minimum(Da, 1.0)
        Dbp, Dbm = np.maximum(Db, 1.0), np.minimum(Db, 1.0)
        mat1 = sla.solve(Tb.T, Ta.T, check_finite=False).T           # Ta / Tb
        mat1 = mat1 * (Dam[:, None] / Dbp[None, :])
        mat2 = (Ua.conj().T @ Ub) * (Dbm[None, :] / Dap[:, None])
        U, D, T = decompose_udt(mat1 + mat2)
        mat1 = sla.solve(D[:, None] * T, U.conj().T, check_finite=False)
        mat1 = mat1 / Dbp[:, None] / Dap[None, :]
        U, D, T = decompose_udt(mat1)
        U = sla.solve(Tb, U, check_finite=False)
        T = T @ Ua.conj().T
        return (U * D[None, :]) @ T

    def measure_tdgfs(self):
        """fermion_measurements.jl:1343-1407 ``measure_tdgfs!``: Gt0[tau] = G(tau,0), G0t[tau] = G(0,tau) for all M slices
        (0-based tau here): stabilised at the safe_mult slices, B-propagated in between (``fill_tdgf!`` :1509-1541)."""
        M, sm = self.p.slices, self.p.safe_mult
        n = self.n
        Gt0 = np.zeros((M, n, n), dtype=complex)
        G0t = np.zeros((M, n, n), dtype=complex)
        BT0Inv = self.calc_Bchain_udts(invert=True, left=True)
        BBetaT = self.calc_Bchain_udts(invert=False, left=False)
        BT0 = self.calc_Bchain_udts(invert=False, left=True)
        BBetaTInv = self.calc_Bchain_udts(invert=True, left=False)
        self.tdgf_stacks = dict(BT0Inv=BT0Inv, BBetaT=BBetaT, BT0=BT0, BBetaTInv=BBetaTInv)
        for i, tau in enumerate(range(0, M, sm)):
            if i != 0:
                g = self.inv_sum_udts_scalettar(BT0Inv[0][i - 1], BT0Inv[1][i - 1], BT0Inv[2][i - 1],
                                                BBetaT[0][i], BBetaT[1][i], BBetaT[2][i])
                Gt0[tau] = self.effective_greens2greens(g)
                g = self.inv_sum_udts_scalettar(BT0[0][i - 1], BT0[1][i - 1], BT0[2][i - 1],
                                                BBetaTInv[0][i], BBetaTInv[1][i], BBetaTInv[2][i])
                G0t[tau] = self.effective_greens2greens(g)
            else:
                Gt0[tau] = self.effective_greens2greens(self.inv_one_plus_udt_scalettar(BBetaT[0][0], BBetaT[1][0], BBetaT[2][0]))
                G0t[tau] = self.effective_greens2greens(
                    self.inv_one_plus_udt_scalettar(BBetaTInv[0][0], BBetaTInv[1][0], BBetaTInv[2][0]))
        # fill_tdgf!: reference (1-based) Mhalf = M/2+1; tau in Mhalf:M forward, (Mhalf-1):-1:1 backward
        safe = set(range(0, M, sm))
        mhalf = M // 2              # 0-based index of reference slice Mhalf
        for tau in range(mhalf, M):
            if tau in safe:
                continue
            Gt0[tau] = self.multiply_B_left(tau, Gt0[tau - 1].copy())
            G0t[tau] = self.multiply_B_inv_right(tau, G0t[tau - 1].copy())
        for tau in range(mhalf - 1, -1, -1):
            if tau in safe:
                continue
            Gt0[tau] = self.multiply_B_inv_left(tau + 1, Gt0[tau + 1].copy())
            G0t[tau] = self.multiply_B_right(tau + 1, G0t[tau + 1].copy())
        G0t *= -1.0
        self.Gt0, self.G0t = Gt0, G0t
        return Gt0, G0t

    # ---------------------------------------------------------------- helpers used by the reference's tests
    def calc_greens_fresh(self, slc):
        """fermion_measurements.jl:1061-1101 spirit: G(slice) = [1 + B(slice-1)..B(0)B(M-1)..B(slice)]^-1
        by a plain UDT-stabilised product (fresh, independent of the stack)."""
        n, M = self.n, self.p.slices
        order = list(range(slc, M)) + list(range(0, slc))
        U = np.eye(n, dtype=complex); D = np.ones(n); T = np.eye(n, dtype=complex)
        for k, s in enumerate(order):
            U = self.multiply_B_left(s, U)
            if (k + 1) % self.p.safe_mult == 0 or k == len(order) - 1:
                U = U * D[None, :]
                U, D, Tn = decompose_udt(U)
                T = Tn @ T
        # [1 + U D T]^-1 = [U (U^† T^-1 + D) T]^-1
        X = sla.solve(T.conj().T, U, check_finite=False).conj().T   # U^† T^-1
        X[np.diag_indices_from(X)] += D
        u, d, t = decompose_udt(X)
        return sla.inv(t @ T, check_finite=False) @ ((1.0 / d)[:, None] * (U @ u).conj().T)
