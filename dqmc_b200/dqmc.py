"""Host-side mirror of the reference's driver-facing interface for the hot path.

`DQMC` keeps the names, argument meaning and 1-based slice/site conventions of the reference
(src/dqmc_framework.jl:98-177, 500-517; src/stack.jl; src/local_updates.jl; src/slice_matrices.jl)
so that tests read like the reference's own; every method forwards to the C ABI of
libdqmc_b200.so (include/dqmc_b200.h).  No arithmetic of the path happens in Python and there
is no CPU fallback.
"""
import ctypes as C

import numpy as np

from . import lib as _l
from .model import Lattice, Params


class UniformStream:
    """Buffered uniform [0,1) source standing in for the reference's global MersenneTwister.

    The library consumes a caller-supplied stream in the reference's order and reports how many draws it
    used; `take`/`advance` make the stream position independent of how calls are chunked (the Julia side
    would generate from `copy(rng)` and then discard `consumed` draws from the real RNG).
    """

    def __init__(self, source):
        if isinstance(source, (int, np.integer)):
            source = np.random.Generator(np.random.Philox(int(source)))
        self._gen = source if hasattr(source, "random") else None
        self._buf = np.zeros(0) if self._gen is not None else np.ascontiguousarray(source, dtype=np.float64)
        self.consumed = 0

    def take(self, n):
        if len(self._buf) < n:
            if self._gen is None:
                raise _l.DqmcError("uniform stream exhausted")
            self._buf = np.concatenate([self._buf, self._gen.random(n - len(self._buf))])
        return np.ascontiguousarray(self._buf[:n])

    def advance(self, k):
        self._buf = self._buf[k:]
        self.consumed += k


class DQMC:
    """`DQMC{CBAssaad,ComplexF64,H}` of the reference (dqmc_framework.jl:98-135) with its stack on one B200."""

    def __init__(self, p: Params, device=0, delay=0):
        self.p = p
        self.l = Lattice(p)                       # load_lattice (lattice.jl:57-67)
        self.lib = _l.load()
        self.n = p.flv * self.l.sites
        self._ctx = C.c_void_p()
        self.boson_action = 0.0
        self.acc_rate = 0.0
        self.acc_rate_global = 0.0
        self.acc_global = 0
        self.prop_global = 0
        cp = _l.DqmcParams(L=p.L, flv=p.flv, opdim=p.opdim, slices=p.slices, safe_mult=p.safe_mult,
                           edrun=int(p.edrun), all_checks=int(p.all_checks), device=device, delay=delay, reserved=0,
                           delta_tau=p.delta_tau, lambda_=p.lambda_, r=p.r, c=p.c, u=p.u)
        rc = self.lib.dqmc_create(C.byref(self._ctx), C.byref(cp))
        if rc != 0:
            raise _l.DqmcError(self.lib.dqmc_last_error(None).decode())
        l = self.l
        for which, m in ((_l.OP_HOP_HALF_B, l.chkr_hop_half[1]), (_l.OP_HOP_A, l.chkr_hop[0]),
                         (_l.OP_HOP_HALF_INV_B, l.chkr_hop_half_inv[1]), (_l.OP_HOP_INV_A, l.chkr_hop_inv[0]),
                         (_l.OP_MU, l.chkr_mu), (_l.OP_MU_INV, l.chkr_mu_inv),
                         (_l.OP_HOP_HALF_A, l.chkr_hop_half[0]), (_l.OP_HOP_HALF_INV_A, l.chkr_hop_half_inv[0])):
            self._set_operator(which, m)
        nb = np.asfortranarray(l.neighbors.astype(np.int64))
        self._chk(self.lib.dqmc_set_neighbors(self._ctx, nb.ctypes.data_as(_l._I64)))

    # -- plumbing ----------------------------------------------------------------------------------------
    def _chk(self, rc):
        if rc != 0:
            raise _l.DqmcError(self.lib.dqmc_last_error(self._ctx).decode())

    def _set_operator(self, which, m):
        m = m.tocsc()
        m.sort_indices()
        colptr = (m.indptr.astype(np.int64) + 1)
        rowval = (m.indices.astype(np.int64) + 1)
        is_c = np.iscomplexobj(m.data)
        nz = np.ascontiguousarray(m.data, dtype=np.complex128 if is_c else np.float64)
        self._chk(self.lib.dqmc_set_operator(self._ctx, which, m.shape[0], m.shape[1], colptr.ctypes.data_as(_l._I64),
                                             rowval.ctypes.data_as(_l._I64), nz.ctypes.data_as(C.c_void_p), int(is_c)))

    def close(self):
        if self._ctx:
            self.lib.dqmc_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- state (mc.s.current_slice, mc.s.direction, mc.s.greens, mc.p.hsfield, mc.s.log_det) --------------
    @property
    def current_slice(self):
        s, d = C.c_int32(), C.c_int32()
        self.lib.dqmc_get_state(self._ctx, C.byref(s), C.byref(d))
        return s.value

    @property
    def direction(self):
        s, d = C.c_int32(), C.c_int32()
        self.lib.dqmc_get_state(self._ctx, C.byref(s), C.byref(d))
        return d.value

    def set_state(self, slice_, direction):
        self._chk(self.lib.dqmc_set_state(self._ctx, slice_, direction))

    @property
    def greens(self):
        g = _l.cplx_buf((self.n, self.n))
        self._chk(self.lib.dqmc_get_greens(self._ctx, _l.dptr(g)))
        return g

    @greens.setter
    def greens(self, g):
        g = _l.cplx_in(g, (self.n, self.n))
        self._chk(self.lib.dqmc_set_greens(self._ctx, _l.dptr(g)))

    @property
    def hsfield(self):
        h = np.zeros((self.p.opdim, self.l.sites, self.p.slices), order="F")
        self._chk(self.lib.dqmc_get_hsfield(self._ctx, _l.dptr(h)))
        return h

    @hsfield.setter
    def hsfield(self, h):
        h = np.asfortranarray(np.asarray(h, dtype=np.float64))
        assert h.shape == (self.p.opdim, self.l.sites, self.p.slices)
        self._chk(self.lib.dqmc_set_hsfield(self._ctx, _l.dptr(h)))

    @property
    def log_det(self):
        v = C.c_double()
        self._chk(self.lib.dqmc_logdet(self._ctx, C.byref(v)))
        return v.value

    # -- action.jl:1-52 (host, O(N*M); evaluated once at init like the reference) ------------------------
    def calc_boson_action(self, hsfield=None):
        p, L = self.p, self.l.L
        hs = self.hsfield if hsfield is None else np.asarray(hsfield)
        sq = np.einsum("kis,kis->is", hs, hs)
        S = p.delta_tau * p.r / 2.0 * sq.sum()
        if not p.edrun:
            h = hs.reshape(p.opdim, L, L, p.slices, order="F")
            S += 0.5 / p.delta_tau / p.c ** 2 * ((h - np.roll(h, -1, axis=3)) ** 2).sum()
            S += 0.5 * p.delta_tau * (((h - np.roll(h, -1, axis=1)) ** 2).sum() + ((h - np.roll(h, -1, axis=2)) ** 2).sum())
            S += p.delta_tau * p.u / 4.0 * (sq * sq).sum()
        return float(S)

    # -- dqmc_framework.jl:152-177 -----------------------------------------------------------------------
    def init(self, start_conf=None, rng=None):
        """`init!(mc[, start_conf])`: field, boson action, stack, first propagate."""
        p = self.p
        if start_conf is None:
            rng = rng if rng is not None else np.random.Generator(np.random.Philox(p.seed % (2 ** 63)))
            start_conf = rng.random((p.slices, self.l.sites, p.opdim)).T   # column-major fill order like rand(opdim,N,M)
        self.hsfield = start_conf
        self.boson_action = self.calc_boson_action(np.asarray(start_conf))
        self.build_stack()
        self.propagate()

    # -- stack.jl ----------------------------------------------------------------------------------------
    def build_stack(self):
        self._chk(self.lib.dqmc_build_stack(self._ctx))

    def propagate(self):
        s, d = C.c_int32(), C.c_int32()
        self._chk(self.lib.dqmc_propagate(self._ctx, C.byref(s), C.byref(d)))
        return s.value, d.value

    def wrap_greens(self, gf=None, slice_=None, direction=None):
        """`wrap_greens!(mc, gf, slice, direction)`; gf=None wraps mc.s.greens on the device."""
        slice_ = self.current_slice if slice_ is None else slice_
        direction = self.direction if direction is None else direction
        if gf is None:
            self._chk(self.lib.dqmc_wrap_greens(self._ctx, None, slice_, direction))
            return None
        g = _l.cplx_in(gf, (self.n, self.n)).copy(order="F")
        self._chk(self.lib.dqmc_wrap_greens(self._ctx, _l.dptr(g), slice_, direction))
        return g

    def calculate_greens(self, Ul, Dl, Tl, Ur, Dr, Tr):
        """`calculate_greens(mc)` with explicit UDT inputs (the reference reads them from mc.s)."""
        n = self.n
        mats = [_l.cplx_in(a, (n, n)) for a in (Ul, Tl, Ur, Tr)]
        ds = [np.ascontiguousarray(d, dtype=np.float64) for d in (Dl, Dr)]
        g = _l.cplx_buf((n, n))
        self._chk(self.lib.dqmc_calculate_greens_from(self._ctx, _l.dptr(mats[0]), _l.dptr(ds[0]), _l.dptr(mats[1]),
                                                      _l.dptr(mats[2]), _l.dptr(ds[1]), _l.dptr(mats[3]), _l.dptr(g)))
        return g

    def decompose_udt(self, X):
        """`decompose_udt!(A, D)` (linalg.jl:20-39) -> U, D, T."""
        n = self.n
        x = _l.cplx_in(X, (n, n))
        U, T, D = _l.cplx_buf((n, n)), _l.cplx_buf((n, n)), np.zeros(n)
        self._chk(self.lib.dqmc_decompose_udt(self._ctx, _l.dptr(x), _l.dptr(U), _l.dptr(D), _l.dptr(T)))
        return U, D, T

    # -- slice_matrices.jl:101-226 -----------------------------------------------------------------------
    def _mulB(self, op, slice_, M):
        m = _l.cplx_in(M, (self.n, self.n)).copy(order="F")
        self._chk(self.lib.dqmc_multiply_B(self._ctx, op, slice_, _l.dptr(m)))
        return m

    def multiply_B_left(self, slice_, M):
        return self._mulB(_l.B_LEFT, slice_, M)

    def multiply_B_right(self, slice_, M):
        return self._mulB(_l.B_RIGHT, slice_, M)

    def multiply_B_inv_left(self, slice_, M):
        return self._mulB(_l.B_INV_LEFT, slice_, M)

    def multiply_B_inv_right(self, slice_, M):
        return self._mulB(_l.B_INV_RIGHT, slice_, M)

    def multiply_daggered_B_left(self, slice_, M):
        return self._mulB(_l.B_DAGGER_LEFT, slice_, M)

    def slice_matrix(self, slice_, power=1.0):
        I = np.eye(self.n, dtype=np.complex128)
        return self.multiply_B_left(slice_, I) if power > 0 else self.multiply_B_inv_left(slice_, I)

    # -- local_updates.jl:1-39 ---------------------------------------------------------------------------
    def local_updates(self, stream: UniformStream):
        """One pass over the sites of mc.s.current_slice; returns the acceptance fraction."""
        N = self.l.sites
        u = stream.take(4 * N)
        consumed, accepted, dS = C.c_int64(), C.c_int64(), C.c_double()
        self._chk(self.lib.dqmc_local_updates(self._ctx, self.p.box, _l.dptr(u), len(u), C.byref(consumed),
                                              C.byref(accepted), C.byref(dS)))
        stream.advance(consumed.value)
        self.boson_action += dS.value
        return accepted.value / N

    # -- global_updates.jl:18-59 -------------------------------------------------------------------------
    def global_update(self, stream: UniformStream):
        """Uniform shift of the whole field, full stack rebuild, Metropolis test on exp(-dS) det ratio.  Returns 0/1."""
        u = stream.take(4)
        S_new, acc, used = C.c_double(), C.c_int32(), C.c_int32()
        self._chk(self.lib.dqmc_global_update(self._ctx, self.p.box_global, _l.dptr(u), self.boson_action, C.byref(S_new),
                                              C.byref(acc), C.byref(used)))
        stream.advance(used.value)
        self.boson_action = S_new.value
        return acc.value

    # -- boson_measurements.jl:6-39 ----------------------------------------------------------------------
    def measure_chi_dynamic(self):
        """chi(qy,qx,iw) of the current configuration, shape (L/2+1, L/2+1, M/2+1)."""
        nq, nt = self.l.L // 2 + 1, self.p.slices // 2 + 1
        chi = np.zeros((nq, nq, nt), order="F")
        self._chk(self.lib.dqmc_measure_chi_dynamic(self._ctx, _l.dptr(chi)))
        return chi

    # -- fermion_measurements.jl:1343-1541 --------------------------------------------------------------
    def measure_tdgfs(self):
        """`measure_tdgfs!`: G(tau,0) and G(0,tau) for all slices, kept on the device (read them with Gt0/G0t)."""
        self._chk(self.lib.dqmc_measure_tdgfs(self._ctx))

    def _tdgf(self, which, slc):
        g = _l.cplx_buf((self.n, self.n))
        self._chk(self.lib.dqmc_get_tdgf(self._ctx, which, int(slc), _l.dptr(g)))
        return g

    def Gt0(self, slc):
        """`mc.s.meas.Gt0[slc]` (1-based slice)."""
        return self._tdgf(0, slc)

    def G0t(self, slc):
        """`mc.s.meas.G0t[slc]` (1-based slice)."""
        return self._tdgf(1, slc)

    def deallocate_tdgfs_stacks(self):
        self._chk(self.lib.dqmc_free_tdgfs(self._ctx))

    def inv_sum_udts_scalettar(self, Ua, Da, Ta, Ub, Db, Tb):
        """`inv_sum_udts_scalettar!` (linalg.jl:512-567) on host operands."""
        n = self.n
        res = _l.cplx_buf((n, n))
        mats = [_l.cplx_in(x, (n, n)) for x in (Ua, Ta, Ub, Tb)]
        da, db = np.ascontiguousarray(Da, dtype=np.float64), np.ascontiguousarray(Db, dtype=np.float64)
        self._chk(self.lib.dqmc_inv_sum_udts(self._ctx, _l.dptr(mats[0]), _l.dptr(da), _l.dptr(mats[1]),
                                             _l.dptr(mats[2]), _l.dptr(db), _l.dptr(mats[3]),
                                             _l.dptr(res)))
        return res

    def measure_chi_static(self):
        return float(self.measure_chi_dynamic()[0, 0, 0])

    def device_boson_action(self):
        v = C.c_double()
        self._chk(self.lib.dqmc_calc_boson_action(self._ctx, C.byref(v)))
        return v.value

    # -- dqmc_framework.jl:500-517 -----------------------------------------------------------------------
    def update(self, stream: UniformStream, i=0):
        """`update(mc, i)`: propagate, a global update every `global_rate`-th sweep at (slices, -1), local updates."""
        p = self.p
        s, d = self.propagate()
        if p.global_updates and s == p.slices and d == -1 and i % p.global_rate == 0:
            self.prop_global += 1
            b = self.global_update(stream)
            self.acc_rate_global += b
            self.acc_global += b
        self.acc_rate += self.local_updates(stream)

    def sweep(self, stream=None, nupdates=None):
        """`nupdates` x update(mc) in one library call (default: one sweep = M updates).

        stream=None uses the device-resident uniforms uploaded with `set_uniforms`.  Returns
        (accepted proposals, consumed uniforms).
        """
        N, M = self.l.sites, self.p.slices
        nupdates = M if nupdates is None else nupdates
        consumed, accepted, dS = C.c_int64(), C.c_int64(), C.c_double()
        if stream is None:
            self._chk(self.lib.dqmc_sweep(self._ctx, nupdates, self.p.box, None, 0, C.byref(consumed),
                                          C.byref(accepted), C.byref(dS)))
        else:
            u = stream.take(4 * N * nupdates)
            self._chk(self.lib.dqmc_sweep(self._ctx, nupdates, self.p.box, _l.dptr(u), len(u), C.byref(consumed),
                                          C.byref(accepted), C.byref(dS)))
            stream.advance(consumed.value)
        self.boson_action += dS.value
        self.acc_rate += accepted.value / N
        return accepted.value, consumed.value

    def set_uniforms(self, u):
        u = np.ascontiguousarray(u, dtype=np.float64)
        self._chk(self.lib.dqmc_set_uniforms(self._ctx, _l.dptr(u), len(u)))

    # -- telemetry ---------------------------------------------------------------------------------------
    def set_timing(self, level=1):
        """0 off, 1 sweep total only (CUDA graphs stay on), 2 all phase timers (graphs bypassed)."""
        self._chk(self.lib.dqmc_set_timing(self._ctx, int(level)))

    def timers(self):
        ms = np.zeros(5)
        self._chk(self.lib.dqmc_timers(self._ctx, _l.dptr(ms), 5))
        return dict(zip(("wrap", "local_updates", "stack_udt", "calculate_greens", "sweep"), ms.tolist()))

    def checks(self):
        e, k = C.c_double(), C.c_int64()
        self._chk(self.lib.dqmc_checks(self._ctx, C.byref(e), C.byref(k)))
        return e.value, k.value

    def sync(self):
        self._chk(self.lib.dqmc_sync(self._ctx))

    def bench_kernel(self, which, reps):
        ms = C.c_double()
        self._chk(self.lib.dqmc_bench_kernel(self._ctx, which, reps, C.byref(ms)))
        return ms.value

    def lu_profile(self, enable=True, read=False):
        out = np.zeros(32, dtype=np.int64)
        self._chk(self.lib.dqmc_lu_profile(self._ctx, int(enable), out.ctypes.data_as(_l._I64) if read else None))
        return out

    def kernel_launches(self):
        return int(self.lib.dqmc_kernel_launches(self._ctx))

    def test_qr_paired(self, XL, rhs, lookahead=True):
        """Paired Householder QR (half-matrix path) on host data: XL, rhs n x n/2 with pair-interleaved rows.
        Returns R_L, Q^H rhs, V (n x n explicit reflector blocks), T factors (n/32, 32, 32), dabs (n)."""
        n, h = self.n, self.n // 2
        x = _l.cplx_in(XL, (n, h)).copy(order="F")
        r = _l.cplx_in(rhs, (n, h)).copy(order="F")
        V = _l.cplx_buf((n, n))
        Tf = np.zeros((n // 32, 32, 32), dtype=np.complex128)
        dabs = np.zeros(n)
        self._chk(self.lib.dqmc_test_qr_paired(self._ctx, _l.dptr(x), _l.dptr(r), _l.dptr(V), _l.dptr(Tf), _l.dptr(dabs), int(lookahead)))
        return x, r, V, Tf.transpose(0, 2, 1).copy(), dabs      # T factors are column-major on the device

    def test_udt(self, X, paired):
        """The sweep's decompose_udt! (sort-once QR, or the paired half-matrix one for a symmetric X) -> U, D, T."""
        n = self.n
        x = _l.cplx_in(X, (n, n))
        U, T, D = _l.cplx_buf((n, n)), _l.cplx_buf((n, n)), np.zeros(n)
        self._chk(self.lib.dqmc_test_udt(self._ctx, _l.dptr(x), _l.dptr(U), _l.dptr(D), _l.dptr(T), int(paired)))
        return U, D, T

    def test_zgemm(self, opA, opB, A, B, Cmat=None, alpha=1.0, beta=0.0):
        A = np.asfortranarray(A, dtype=np.complex128)
        B = np.asfortranarray(B, dtype=np.complex128)
        M = A.shape[0] if opA == 0 else A.shape[1]
        K = A.shape[1] if opA == 0 else A.shape[0]
        N = B.shape[1] if opB == 0 else B.shape[0]
        Cm = np.zeros((M, N), dtype=np.complex128, order="F") if Cmat is None else np.asfortranarray(Cmat, dtype=np.complex128).copy(order="F")
        al = np.array([np.real(alpha), np.imag(alpha)], dtype=np.float64)
        be = np.array([np.real(beta), np.imag(beta)], dtype=np.float64)
        self._chk(self.lib.dqmc_test_zgemm(self._ctx, opA, opB, M, N, K, _l.dptr(al), _l.dptr(A), A.shape[0], _l.dptr(B),
                                           B.shape[0], _l.dptr(be), _l.dptr(Cm), M))
        return Cm
