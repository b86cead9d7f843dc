"""Pin the CPU oracle against the reference's own fixtures and known answers.

Mirrors test/tests_O3.jl and test/tests_linalg.jl of the reference (line numbers cited per test).
CPU only.
"""
import numpy as np
import pytest

from oracle import Params, OracleDQMC, JuliaMT, build_model
from oracle.dqmc import decompose_udt
from tests.helpers import csc_dense, maxabs

SEED = 4729339882041979125  # parameters.jl:67


def make(bfield, **kw):
    p = Params(L=4, slices=10, delta_tau=0.1, safe_mult=10, Bfield=bfield, **kw)
    return OracleDQMC(p, dense_hoppings=True)


@pytest.fixture(scope="module")
def mc_b():
    return make(True)


@pytest.fixture(scope="module")
def mc_nob():
    return make(False)


def seeded_field():
    return JuliaMT(SEED).rand_array(3, 16, 10)  # dqmc_framework.jl:152-155


def test_julia_mt_reproduces_initial_field(golden_o3):
    # local_updates only touches slice M, so slices 1..9 of afterlocal_hsfield are the seeded start
    h = seeded_field()
    assert np.array_equal(h[:, :, :9], golden_o3["afterlocal_hsfield"][:, :, :9])


def test_lattice_tables(mc_b):
    # tests_O3.jl:267-272
    nb = np.array([[2, 3, 4, 1, 6, 7, 8, 5, 10, 11, 12, 9, 14, 15, 16, 13],
                   [5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 1, 2, 3, 4],
                   [4, 1, 2, 3, 8, 5, 6, 7, 12, 9, 10, 11, 16, 13, 14, 15],
                   [13, 14, 15, 16, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12]])
    assert np.array_equal(mc_b.l.neighbors + 1, nb)
    tn = np.array([[2, 3, 4, 5, 6, 7, 8, 9, 10, 1], [10, 1, 2, 3, 4, 5, 6, 7, 8, 9]])
    assert np.array_equal(mc_b.l.time_neighbors + 1, tn)


def test_boson_action(mc_b, golden_o3):
    # tests_O3.jl:50-59
    mc = mc_b
    mc.hsfield = golden_o3["randconf"].copy()
    assert np.isclose(mc.calc_boson_action(), 75.57712964980982, rtol=1e-13)
    new = np.array([0.370422, 0.797014, 0.956094])
    assert np.isclose(mc.calc_boson_action_diff(2, 9, new), 0.015453643701304112, rtol=1e-12)
    mc.p.edrun = True
    assert np.isclose(mc.calc_boson_action(), 16.348917129437076, rtol=1e-13)
    assert np.isclose(mc.calc_boson_action_diff(2, 9, new), 0.014491910524600016, rtol=1e-12)
    mc.p.edrun = False


def test_hopping_matrices(mc_b, golden_o3):
    # tests_O3.jl:62-66
    assert maxabs(mc_b.l.hopping_matrix_exp, golden_o3["hopping_matrix_exp"]) < 1e-14
    assert maxabs(mc_b.l.hopping_matrix_exp_inv, golden_o3["hopping_matrix_exp_inv"]) < 1e-14


def test_peirls(mc_b, golden_o3):
    # tests_O3_peirls.jl (fixtures peirls{s}{f})
    for s in range(2):
        for f in range(2):
            ref = golden_o3[f"peirls{s+1}{f+1}"]
            got = mc_b.l.peirls[s][f]
            assert np.array_equal(np.isnan(ref), np.isnan(got))
            assert np.allclose(np.nan_to_num(ref), np.nan_to_num(got), atol=1e-15)


def test_checkerboard_nob(mc_nob, golden_o3):
    # tests_O3.jl:92-112
    A, B = mc_nob.l.corners
    assert (A + 1).tolist() == [1, 3, 9, 11] and (B + 1).tolist() == [6, 8, 14, 16]
    l = mc_nob.l
    for k in range(2):
        assert maxabs(l.chkr_hop[k].toarray(), csc_dense(golden_o3, f"nob_chkr_hop{k+1}")) < 1e-15
        assert maxabs(l.chkr_hop_inv[k].toarray(), csc_dense(golden_o3, f"nob_chkr_hop_inv{k+1}")) < 1e-15
        assert maxabs(l.chkr_hop_half[k].toarray(), csc_dense(golden_o3, f"nob_chkr_hop_half{k+1}")) < 1e-15
        assert maxabs(l.chkr_hop_half_inv[k].toarray(), csc_dense(golden_o3, f"nob_chkr_hop_half_inv{k+1}")) < 1e-15
    for nm in ("chkr_mu", "chkr_mu_half", "chkr_mu_inv", "chkr_mu_half_inv"):
        assert maxabs(getattr(l, nm).toarray(), csc_dense(golden_o3, "nob_" + nm)) < 1e-15


def test_checkerboard_bfield(mc_b, golden_o3):
    # tests_O3.jl:115-135
    from oracle.model import build_four_site_hopping_matrix_Bfield
    T = build_four_site_hopping_matrix_Bfield(mc_b.l, 2, 0, 1)
    assert maxabs(T, csc_dense(golden_o3, "build_four_site_hopping_matrix_Bfield")) < 1e-15
    l = mc_b.l
    for k in range(2):
        assert maxabs(l.chkr_hop[k].toarray(), csc_dense(golden_o3, f"chkr_hop{k+1}")) < 1e-14
        assert maxabs(l.chkr_hop_inv[k].toarray(), csc_dense(golden_o3, f"chkr_hop_inv{k+1}")) < 1e-14
        assert maxabs(l.chkr_hop_half[k].toarray(), csc_dense(golden_o3, f"chkr_hop_half{k+1}")) < 1e-14
        assert maxabs(l.chkr_hop_half_inv[k].toarray(), csc_dense(golden_o3, f"chkr_hop_half_inv{k+1}")) < 1e-14
    for nm in ("chkr_mu", "chkr_mu_half", "chkr_mu_inv", "chkr_mu_half_inv"):
        assert maxabs(getattr(l, nm).toarray(), csc_dense(golden_o3, nm)) < 1e-15


def test_checkerboard_error_quadratic():
    # tests_O3.jl:73-89: |e^{-dtau T/2} - chkr| <= dtau^2
    for dt in (0.1, 0.01, 0.001):
        for bf in (False, True):
            p = Params(L=4, slices=10, delta_tau=dt, Bfield=bf)
            l = build_model(p)
            hop_chkr = (l.chkr_hop_half[0] @ l.chkr_hop_half[1] @ l.chkr_mu.sqrt()).toarray()
            assert maxabs(l.hopping_matrix_exp, hop_chkr) <= dt ** 2


def test_interactions(mc_b, mc_nob, golden_o3):
    # tests_O3.jl:155-175.  The nob_ fixtures were dumped with hsfield == randconf, the B-field ones
    # with the seeded start field of init!(mc) (found by trying both; slice 3 either way).
    for mc, pre, h in ((mc_nob, "nob_", golden_o3["randconf"]), (mc_b, "", seeded_field())):
        mc.hsfield = h.copy()
        assert maxabs(mc.interaction_matrix_exp(2, 1.0).toarray(), csc_dense(golden_o3, pre + "eVplus")) < 1e-15
        assert maxabs(mc.interaction_matrix_exp(2, -1.0).toarray(), csc_dense(golden_o3, pre + "eVminus")) < 1e-15
        ev = mc.interaction_matrix_exp_op(np.array([0.130018, 0.792039, 0.683411]), 1.0)
        assert maxabs(ev, golden_o3[pre + "eVexpop"]) < 1e-15


def test_local_updates(golden_o3):
    # tests_O3.jl:179-194
    mc = make(True)
    mc.init(seeded_field())
    assert mc.current_slice == 9 and mc.direction == -1
    dr = mc.calc_detratio(6, np.array([0.488033, 0.0196912, 0.438309]))
    assert abs(dr - (1.000380293015979 + 1.0842021724855044e-18j)) < 1e-14
    assert maxabs(mc.delta_i, golden_o3["delta_i"]) < 1e-15
    assert maxabs(mc.Mmat, golden_o3["M"]) < 1e-15
    mc.update_greens(6)
    assert maxabs(mc.greens, golden_o3["afterupdate_greens"]) < 1e-13
    rng = JuliaMT(123456789)
    acc = mc.local_updates(rng)
    assert acc == 0.6875
    assert np.array_equal(mc.hsfield, golden_o3["afterlocal_hsfield"])   # bit-exact accept/reject sequence
    assert maxabs(mc.greens, golden_o3["afterlocal_greens"]) < 1e-13
    assert np.isclose(mc.boson_action, 75.18407422927604, rtol=1e-13)
    assert rng.idx == 58  # 16*3 proposal draws + 10 accept draws (6 proposals had p_acc > 1)


def test_calculate_greens_from_dumped_udts(golden_o3):
    # tests_O3.jl:215-230
    mc = make(True)
    mc.init(seeded_field())
    for k in ("Ur", "Dr", "Tr", "Ul", "Dl", "Tl"):
        setattr(mc, k, golden_o3[k].copy())
    g = mc.calculate_greens()
    ld = mc.calculate_logdet()
    assert maxabs(g, golden_o3["greens"]) < 1e-13
    gfresh = mc.calc_greens_fresh(0)
    assert maxabs(g, gfresh) < 1e-12
    assert np.isclose(ld, np.linalg.slogdet(gfresh)[1], rtol=1e-10)


def test_wrapping(golden_o3):
    # tests_O3.jl:232-238
    mc = make(True)
    mc.init(seeded_field())
    g = mc.calc_greens_fresh(1)
    assert maxabs(mc.wrap_greens(g.copy(), 1, 1), mc.calc_greens_fresh(2)) < 1e-12
    assert maxabs(mc.wrap_greens(g.copy(), 1, -1), mc.calc_greens_fresh(0)) < 1e-12


def test_propagation_to_slice_one(golden_o3):
    # tests_O3.jl:240-249
    mc = make(True)
    mc.init(seeded_field())
    while mc.current_slice != 0:
        mc.propagate()
    mc.propagate()
    assert mc.direction == 1 and mc.current_slice == 0
    assert maxabs(mc.greens, golden_o3["greens"]) < 1e-12


def test_propagation_always_within_bounds():
    # tests_O3.jl:251-263, with a deeper stack (M=50, safe_mult=10, K=5) and the visit order of SURVEY §3.2
    p = Params(L=4, slices=50, delta_tau=0.1, safe_mult=10, Bfield=True)
    mc = OracleDQMC(p)
    mc.init(JuliaMT(7).rand_array(3, 16, 50))
    order = []
    mc.propagate()
    while not (mc.current_slice == p.slices - 1 and mc.direction == -1):
        order.append((mc.current_slice + 1, mc.direction))
        assert maxabs(mc.calc_greens_fresh(mc.current_slice), mc.greens) < 1e-12
        mc.propagate()
    expect = [(c, -1) for c in range(49, 0, -1)] + [(c, 1) for c in range(1, 51)]
    assert order == expect


def test_slice_matrices(golden_o3):
    # tests_O3.jl:283-332 (CBAssaad variant; field = seeded start, slice 3)
    mc = make(True)
    mc.init(seeded_field())
    B = mc.slice_matrix(2, 1.0)
    Binv = mc.slice_matrix(2, -1.0)
    assert maxabs(B, golden_o3["Bplus"]) < 1e-14
    assert maxabs(Binv, golden_o3["Bminus"]) < 1e-14
    rs = np.random.RandomState(0)
    A = rs.rand(64, 64) + 1j * rs.rand(64, 64)
    assert maxabs(mc.multiply_daggered_B_left(2, A.copy()), B.conj().T @ A) < 1e-13
    assert maxabs(mc.multiply_B_inv_left(2, mc.multiply_B_left(2, A.copy())), A) < 1e-13
    assert maxabs(mc.multiply_B_inv_right(2, mc.multiply_B_right(2, A.copy())), A) < 1e-13
    I = np.eye(64, dtype=complex)
    assert maxabs(mc.multiply_B_inv_right(2, mc.multiply_B_left(2, I.copy())), I) < 1e-13
    assert maxabs(mc.multiply_B_inv_left(2, mc.multiply_B_right(2, I.copy())), I) < 1e-13


def test_decompose_udt_properties(golden_linalg):
    # tests_linalg.jl:37-60 (properties) and :105-113 (multiply_safely extremes on the B_QR fixture)
    rs = np.random.RandomState(1234)
    X = rs.rand(10, 10)
    U, D, T = decompose_udt(X)
    assert np.allclose(U @ U.conj().T, np.eye(10))
    assert np.allclose(U @ np.diag(D) @ T, X)
    assert np.all(D > 0)
    U, D, T = golden_linalg["B_QR_U"], golden_linalg["B_QR_D"], golden_linalg["B_QR_T"]
    # multiply_safely (linalg.jl:80-100 region): (U1 D1 T1)(U2 D2 T2) = U1 [D1 (T1 U2) D2] T2 -> UDT
    mat = (D[:, None] * (T @ U)) * D[None, :]
    u, d, t = decompose_udt(mat)
    assert np.isclose(d.max(), 5.8846316709257896e16, rtol=1e-9)
    assert np.isclose(d.min(), 1.9123571535539083e-24, rtol=1e-6)


def test_chi_dynamic(mc_b, golden_o3):
    # tests_O3_measurements.jl:1-6
    chi = mc_b.measure_chi_dynamic(golden_o3["randconf"])
    assert maxabs(chi, golden_o3["chi_dyn"]) < 1e-12
    assert np.isclose(chi[0, 0, 0], 12.420575691388407, rtol=1e-13)


# ------------------------------------------------------------------------------------------ time-displaced G
def _free_tdgf_analytic(p, L, M, tau_i, g0t=False):
    """test/deps/ed.jl:664-707: eps(k) = -2 th cos kx - 2 tv cos ky - mu, G(tau,0)(k) = e^{-tau eps}/(1+e^{-beta eps}),
    G(0,tau)(k) = -e^{tau eps}/(1+e^{beta eps}); transformed to real space (site = y + L x, flavours x,y,x,y)."""
    N = L * L
    beta, tau = M * p.delta_tau, tau_i * p.delta_tau
    ks = 2 * np.pi * np.fft.fftfreq(L)
    ys, xs = np.meshgrid(np.arange(L), np.arange(L), indexing="ij")
    ys, xs = ys.ravel(order="F"), xs.ravel(order="F")
    out = np.zeros((4 * N, 4 * N), dtype=complex)
    for f, (th, tv) in enumerate([(1.0, 0.5), (-0.5, -1.0), (1.0, 0.5), (-0.5, -1.0)]):
        blk = np.zeros((N, N), dtype=complex)
        for ky in ks:
            for kx in ks:
                e = -2 * th * np.cos(kx) - 2 * tv * np.cos(ky) - p.mu1
                g = -np.exp(tau * e) / (1 + np.exp(beta * e)) if g0t else np.exp(-tau * e) / (1 + np.exp(-beta * e))
                ph = np.exp(1j * (ky * ys + kx * xs))
                blk += g * np.outer(ph, ph.conj()) / N
        out[f * N:(f + 1) * N, f * N:(f + 1) * N] = blk
    return out


def test_tdgf_free_fermions_analytic():
    # tests_freefermions.jl:174-237 (free fermions, no checkerboard): measure_tdgfs! against the k-space formulas,
    # Gt0[1] == greens, Gt0[1] - G0t[1] == 1, G(tau,0) == -G(0,beta-tau)
    L, M = 4, 40
    p = Params(L=L, slices=M, safe_mult=10, Bfield=False, lam=0.0)
    om = OracleDQMC(p, dense_hoppings=True)
    om.dense_B = True
    om.init(np.random.RandomState(3).rand(3, L * L, M) + 0.1)
    Gt0, G0t = om.measure_tdgfs()
    for tau_i in (0, 1, 7, 10, 19, 20, 25, 39):
        assert maxabs(Gt0[tau_i], _free_tdgf_analytic(p, L, M, tau_i)) < 1e-12
        assert maxabs(G0t[tau_i], _free_tdgf_analytic(p, L, M, tau_i, g0t=True)) < 1e-12
    assert maxabs(Gt0[0], om.effective_greens2greens(om.greens)) < 1e-12
    assert maxabs(Gt0[0] - G0t[0], np.eye(4 * L * L)) < 1e-12
    assert max(maxabs(Gt0[t], -G0t[M - t]) for t in range(1, M)) < 1e-10


def test_tdgf_interacting_consistency():
    # tests_O3_measurements.jl:195-212 spirit (CBAssaad, B-field, lambda > 0): Gt0[1] is the equal-time G of slice 1
    # (effective -> actual), Gt0[1] - G0t[1] = 1, the four UDT chains have unitary U, and the inverse sums agree with a
    # plain dense inverse of the same chains
    L, M = 4, 20
    p = Params(L=L, slices=M, safe_mult=10, Bfield=True)
    om = OracleDQMC(p)
    om.init(np.random.RandomState(5).rand(3, L * L, M))
    Gt0, G0t = om.measure_tdgfs()
    n = om.n
    g1 = om.effective_greens2greens(om.calc_greens_fresh(0))
    assert maxabs(Gt0[0], g1) < 1e-11
    assert maxabs(Gt0[0] - G0t[0], np.eye(n)) < 1e-11
    for us, ds, ts in om.tdgf_stacks.values():
        for U in us:
            assert maxabs(U @ U.conj().T, np.eye(n)) < 1e-12
    # dense check of slice 11 (i = 2): G(tau,0) = [B(tau,1)^-1 + B(beta,tau)]^-1 with tau = ranges[1][0]
    Binv = np.eye(n, dtype=complex)
    for s in range(0, 10):
        Binv = om.multiply_B_inv_right(s, Binv)
    Bbt = np.eye(n, dtype=complex)
    for s in range(M - 1, 9, -1):
        Bbt = om.multiply_B_right(s, Bbt)
    dense = om.effective_greens2greens(np.linalg.inv(Binv + Bbt))
    assert maxabs(Gt0[10], dense) < 1e-9


# ---------------------------------------------------------------------------------- antiunitary flavour symmetry
def _sym_residual(G):
    """max |G - [[A, B], [-conj(B), conj(A)]]| over the lower half, flavour blocks (1,2 | 3,4)."""
    h = G.shape[0] // 2
    A, B = G[:h, :h], G[:h, h:]
    return max(np.abs(G[h:, :h] + B.conj()).max(), np.abs(G[h:, h:] - A.conj()).max())


def test_antiunitary_symmetry_of_reference_greens(golden_o3):
    # The device path computes the upper half of G in the local-update flush and in calculate_greens and mirrors the lower
    # half.  Pinned here on the REFERENCE's own dumps (test/data/O3.jld, B-field on): the stabilized G, the G after one
    # Woodbury update and the G after a whole slice of local updates all have the structure to rounding.
    for key in ("greens", "afterupdate_greens", "afterlocal_greens"):
        assert _sym_residual(golden_o3[key]) < 1e-13, key


def test_antiunitary_symmetry_along_oracle_sweep():
    import oracle
    from oracle.dqmc import UniformStream
    for bfield in (False, True):
        om = oracle.OracleDQMC(oracle.Params(L=4, slices=20, safe_mult=10, Bfield=bfield))
        rs = np.random.RandomState(5)
        om.init(rs.rand(3, 16, 20))
        st = UniformStream(rs.rand(4 * 16 * 40))
        worst = _sym_residual(om.greens)
        for _ in range(40):
            om.propagate()
            om.local_updates(st)
            worst = max(worst, _sym_residual(om.greens))
        assert worst < 1e-12


def test_paired_householder_udt_prototype():
    # oracle/experiments/quaternion_qr.py (preparation of the half-matrix device path): a UDT that eliminates a column and
    # its antiunitary partner per step, on a matrix graded over 80 orders of magnitude
    from oracle.experiments.quaternion_qr import paired_udt, full_from_left
    rs = np.random.RandomState(2)
    h = 16
    A = rs.randn(h, h) + 1j * rs.randn(h, h)
    B = rs.randn(h, h) + 1j * rs.randn(h, h)
    Dh = np.logspace(40, -40, h)[rs.permutation(h)]
    X = np.block([[A, B], [-B.conj(), A.conj()]]) * np.concatenate([Dh, Dh])[None, :]
    QL, D, TL = paired_udt(X)
    Q, T, Df = full_from_left(QL), full_from_left(TL), np.concatenate([D, D])
    assert maxabs(Q.conj().T @ Q, np.eye(2 * h)) < 1e-13
    assert np.max(np.abs((Q * Df[None, :]) @ T - X) / np.linalg.norm(X, axis=0)[None, :]) < 1e-13
    assert np.all(np.diff(D) <= 1e-12 * D[:-1]) and np.linalg.cond(T) < 1e3


def test_half_matrix_stabilization_prototype():
    # oracle/experiments/half_matrix_greens.py: paired-UDT stacks, blocked paired QR with the right-hand side carried along,
    # staircase back substitution - G from half matrices only, against the reference algorithm (beta = 10, D over 12 orders)
    from oracle.experiments.half_matrix_greens import main
    assert main(L=4, M=100) < 1e-12
