"""Model setup of the reference, restated (host-side, one-off).  Test infrastructure.

Follows parameters.jl:4-80,88-190 (Params and defaults), lattice.jl:105-153 (neighbour tables),
hoppings.jl:16-95 (dense hopping exponentials), hoppings.jl:97-172 (Peierls phases),
hoppings_checkerboard.jl:4-136 (Assaad four-site checkerboard, analytic) and :141-270 (B-field,
numeric exponentials).  All indices here are 0-based; the reference's are 1-based.
"""
from dataclasses import dataclass, field

import numpy as np
import scipy.linalg as sla
import scipy.sparse as sp


@dataclass
class Params:
    """parameters.jl:4-80 (fields) and :53-79 (defaults)."""
    L: int = 4
    slices: int = 10
    delta_tau: float = 0.1
    safe_mult: int = 10
    opdim: int = 3
    flv: int = 4
    hoppings: tuple = (1.0, 0.5, -0.5, -1.0)   # "HOPPINGS" reshaped column-major to t[hor/ver, flavour]
    mu1: float = -0.5
    mu2: float = -0.5
    lam: float = 0.5
    r: float = 2.0
    c: float = 3.0
    u: float = 1.0
    box: float = 0.5
    box_global: float = 0.5
    global_updates: bool = False
    global_rate: int = 5
    chkr: bool = True
    Bfield: bool = False
    edrun: bool = False
    all_checks: bool = True
    seed: int = 4729339882041979125

    @property
    def beta(self):
        return self.slices * self.delta_tau


@dataclass
class Lattice:
    L: int = 0
    sites: int = 0
    t: np.ndarray = None                  # t[hor/ver, flavour]  (lattice.jl:59)
    neighbors: np.ndarray = None          # [4, sites]: up, right, down, left (lattice.jl:112-117)
    time_neighbors: np.ndarray = None     # [2, slices]: later, earlier (lattice.jl:144-153)
    peirls: list = None                   # peirls[s][f][trg, src]
    hopping_matrix_exp: np.ndarray = None
    hopping_matrix_exp_inv: np.ndarray = None
    chkr_hop_half: list = field(default_factory=list)
    chkr_hop_half_inv: list = field(default_factory=list)
    chkr_hop_half_dagger: list = field(default_factory=list)
    chkr_hop: list = field(default_factory=list)
    chkr_hop_inv: list = field(default_factory=list)
    chkr_hop_dagger: list = field(default_factory=list)
    chkr_mu_half: sp.csc_matrix = None
    chkr_mu_half_inv: sp.csc_matrix = None
    chkr_mu: sp.csc_matrix = None
    chkr_mu_inv: sp.csc_matrix = None
    corners: tuple = None


def init_neighbors_table(l):
    """lattice.jl:105-138.  sql[y, x] = y + L*x (column-major linear index)."""
    L = l.L
    sql = np.arange(L * L).reshape(L, L, order="F")
    up = np.roll(sql, (-1, 0), axis=(0, 1))
    right = np.roll(sql, (0, -1), axis=(0, 1))
    down = np.roll(sql, (1, 0), axis=(0, 1))
    left = np.roll(sql, (0, 1), axis=(0, 1))
    l.neighbors = np.vstack([a.reshape(-1, order="F") for a in (up, right, down, left)])


def init_time_neighbors_table(l, slices):
    """lattice.jl:144-153 (periodic in imaginary time)."""
    tn = np.zeros((2, slices), dtype=np.int64)
    for s in range(slices):
        tn[0, s] = 0 if s == slices - 1 else s + 1
        tn[1, s] = slices - 1 if s == 0 else s - 1
    l.time_neighbors = tn


def init_hopping_matrix_exp(p, l):
    """hoppings.jl:16-95 (no B-field; nearest-neighbour hoppings only)."""
    N = l.sites
    Tx = np.diag(np.full(N, -p.mu1))
    Ty = np.diag(np.full(N, -p.mu2))
    for src in range(N):
        for nb in (1, 3):  # horizontal: right, left
            trg = l.neighbors[nb, src]
            Tx[trg, src] += -l.t[0, 0]
            Ty[trg, src] += -l.t[0, 1]
        for nb in (0, 2):  # vertical: up, down
            trg = l.neighbors[nb, src]
            Tx[trg, src] += -l.t[1, 0]
            Ty[trg, src] += -l.t[1, 1]
    em = [sla.expm(-0.5 * p.delta_tau * T) for T in (Tx, Ty)]
    ep = [sla.expm(0.5 * p.delta_tau * T) for T in (Tx, Ty)]
    reps = 2 if p.opdim == 3 else 1
    l.hopping_matrix_exp = sla.block_diag(*(em * reps))
    l.hopping_matrix_exp_inv = sla.block_diag(*(ep * reps))


def init_peirls_phases(p, l):
    """hoppings.jl:97-172 (nearest-neighbour part).  phis[x,y,x',y'] then permuted to [trg,src]."""
    L = l.L
    B = np.zeros((2, 2))
    if p.Bfield:
        B[0, 0] = B[1, 1] = 2 * np.pi / l.sites
        B[0, 1] = B[1, 0] = -2 * np.pi / l.sites
    l.peirls = [[None, None], [None, None]]
    for f in range(2):
        for s in range(2):
            phis = np.full((L, L, L, L), np.nan)
            for x in range(1, L + 1):
                for y in range(1, L + 1):
                    xp = x % L + 1
                    yp = y % L + 1
                    phis[x - 1, y - 1, x - 1, yp - 1] = 0
                    phis[x - 1, yp - 1, x - 1, y - 1] = 0
                    phis[x - 1, y - 1, xp - 1, y - 1] = -B[s, f] * (y - 1)
                    phis[xp - 1, y - 1, x - 1, y - 1] = -phis[x - 1, y - 1, xp - 1, y - 1]
                    if y == L:
                        phis[x - 1, y - 1, x - 1, yp - 1] = B[s, f] * L * (x - 1)
                        phis[x - 1, yp - 1, x - 1, y - 1] = -phis[x - 1, y - 1, x - 1, yp - 1]
            # reshape(permutedims(phis,[2,1,4,3]), (sites,sites)) in column-major order
            q = np.transpose(phis, (1, 0, 3, 2))
            l.peirls[s][f] = q.reshape(l.sites, l.sites, order="F")


def init_hopping_matrix_exp_Bfield(p, l):
    """hoppings.jl:174-254 (nearest-neighbour part): dense exponentials with Peierls phases."""
    N = l.sites
    T = [[None, None], [None, None]]
    for f in range(2):
        for s in range(2):
            mu = p.mu1 if f == 0 else p.mu2
            M = np.diag(np.full(N, -mu)).astype(complex)
            for src in range(N):
                for nb in (1, 3):
                    trg = l.neighbors[nb, src]
                    M[trg, src] += -np.exp(1j * l.peirls[s][f][trg, src]) * l.t[0, f]
                for nb in (0, 2):
                    trg = l.neighbors[nb, src]
                    M[trg, src] += -np.exp(1j * l.peirls[s][f][trg, src]) * l.t[1, f]
            T[s][f] = M
    order = [(0, 0), (1, 1), (1, 0), (0, 1)] if p.opdim == 3 else [(0, 0), (1, 1)]  # (spin, flavour)
    l.hopping_matrix_exp = sla.block_diag(*[sla.expm(-0.5 * p.delta_tau * T[s][f]) for s, f in order])
    l.hopping_matrix_exp_inv = sla.block_diag(*[sla.expm(0.5 * p.delta_tau * T[s][f]) for s, f in order])


def find_four_site_hopping_corners(l):
    """hoppings_checkerboard.jl:4-18."""
    L = l.L
    tolin = np.arange(l.sites).reshape(L, L, order="F")
    A = np.array([tolin[y, x] for x in range(0, L, 2) for y in range(0, L, 2)], dtype=np.int64)
    B = l.neighbors[0, l.neighbors[1, A]]
    return A, B


def _four_site_exp_analytic(p, l, corner, tflv, prefac, sign):
    """hoppings_checkerboard.jl:39-57: analytic exp(fac*T_plaquette) on an N x N identity."""
    N = l.sites
    M = np.eye(N)
    i = corner
    j = l.neighbors[0, i]
    m = l.neighbors[1, i]
    n = l.neighbors[1, j]
    th, tv = l.t[0, tflv], l.t[1, tflv]
    fac = sign * (-prefac * p.delta_tau)
    cc = np.cosh(fac * -th) * np.cosh(fac * -tv)
    ss = np.sinh(fac * -th) * np.sinh(fac * -tv)
    cs = np.cosh(fac * -th) * np.sinh(fac * -tv)
    sc = np.sinh(fac * -th) * np.cosh(fac * -tv)
    for a in (i, j, m, n):
        M[a, a] = cc
    M[i, n] = M[j, m] = M[m, j] = M[n, i] = ss
    M[i, j] = M[j, i] = M[m, n] = M[n, m] = cs
    M[i, m] = M[j, n] = M[m, i] = M[n, j] = sc
    return M


def build_four_site_hopping_matrix_Bfield(l, corner, f, s, dtype=complex):
    """hoppings_checkerboard.jl:141-163: plaquette hopping matrix with Peierls phases."""
    N = l.sites
    cw = [corner, l.neighbors[0, corner], l.neighbors[0, l.neighbors[1, corner]], l.neighbors[1, corner]]
    sh = cw[1:] + cw[:1]
    h, v = l.t[0, f], l.t[1, f]
    hop = -1 * np.array([v, h, v, h] * 2, dtype=dtype)
    for k in range(4):
        i, j = cw[k], sh[k]
        hop[k] *= np.exp(1j * l.peirls[s][f][i, j])
        hop[k + 4] *= np.exp(1j * l.peirls[s][f][j, i])
    T = np.zeros((N, N), dtype=dtype)
    for k in range(4):
        T[cw[k], sh[k]] += hop[k]
        T[sh[k], cw[k]] += hop[k + 4]
    return T


def _rem_eff_zeros(X):
    """hoppings_checkerboard_generic.jl:35."""
    X = X.copy()
    X[np.abs(X) < 1e-15] = 0
    return X


def _mu_factors(p, l):
    """hoppings_checkerboard.jl:122-127."""
    muv = np.concatenate([np.full(l.sites, p.mu1), np.full(l.sites, p.mu2)])
    muv = np.tile(muv, p.flv // 2)
    l.chkr_mu_half = sp.diags(np.exp(-0.5 * p.delta_tau * -muv)).tocsc()
    l.chkr_mu_half_inv = sp.diags(np.exp(0.5 * p.delta_tau * -muv)).tocsc()
    l.chkr_mu = sp.diags(np.exp(-p.delta_tau * -muv)).tocsc()
    l.chkr_mu_inv = sp.diags(np.exp(p.delta_tau * -muv)).tocsc()


def _fold(mats):
    out = mats[0]
    for m in mats[1:]:
        out = out @ m
    return out


def init_checkerboard_matrices(p, l):
    """hoppings_checkerboard.jl:65-136 (no B-field, analytic)."""
    A, B = find_four_site_hopping_corners(l)
    l.corners = (A, B)
    corners = (A, B)

    def group(g, tflv, prefac, sign):
        return _fold([_four_site_exp_analytic(p, l, c, tflv, prefac, sign) for c in corners[g]])

    flavours = [0, 1, 0, 1] if p.opdim == 3 else [0, 1]
    for g in range(2):
        l.chkr_hop_half.append(sp.csc_matrix(sla.block_diag(*[group(g, f, 0.5, +1) for f in flavours])))
        l.chkr_hop_half_inv.append(sp.csc_matrix(sla.block_diag(*[group(g, f, 0.5, -1) for f in flavours])))
        l.chkr_hop.append(sp.csc_matrix(sla.block_diag(*[group(g, f, 1.0, +1) for f in flavours])))
        l.chkr_hop_inv.append(sp.csc_matrix(sla.block_diag(*[group(g, f, 1.0, -1) for f in flavours])))
    l.chkr_hop_half_dagger = [m.conj().T.tocsc() for m in l.chkr_hop_half]
    l.chkr_hop_dagger = [m.conj().T.tocsc() for m in l.chkr_hop]
    _mu_factors(p, l)


def init_checkerboard_matrices_Bfield(p, l):
    """hoppings_checkerboard.jl:165-270 (numeric plaquette exponentials, spin/flavour blocks)."""
    A, B = find_four_site_hopping_corners(l)
    l.corners = (A, B)
    corners = (A, B)

    def group(g, s, f, fac):
        return _fold([_rem_eff_zeros(sla.expm(fac * build_four_site_hopping_matrix_Bfield(l, c, f, s)))
                      for c in corners[g]])

    order = [(0, 0), (1, 1), (1, 0), (0, 1)] if p.opdim == 3 else [(0, 0), (1, 1)]  # (spin, flavour)
    dt = p.delta_tau
    for g in range(2):
        l.chkr_hop_half.append(sp.csc_matrix(sla.block_diag(*[group(g, s, f, -0.5 * dt) for s, f in order])))
        l.chkr_hop_half_inv.append(sp.csc_matrix(sla.block_diag(*[group(g, s, f, 0.5 * dt) for s, f in order])))
        l.chkr_hop.append(sp.csc_matrix(sla.block_diag(*[group(g, s, f, -dt) for s, f in order])))
        l.chkr_hop_inv.append(sp.csc_matrix(sla.block_diag(*[group(g, s, f, dt) for s, f in order])))
    l.chkr_hop_half_dagger = [m.conj().T.tocsc() for m in l.chkr_hop_half]
    l.chkr_hop_dagger = [m.conj().T.tocsc() for m in l.chkr_hop]
    _mu_factors(p, l)


def build_model(p, dense_hoppings=True):
    """load_lattice (lattice.jl:57-67) + init_hopping_matrices (hoppings.jl:1-14)."""
    l = Lattice()
    l.L = p.L
    l.sites = p.L * p.L
    l.t = np.array(p.hoppings, dtype=float).reshape(2, 2, order="F")
    init_neighbors_table(l)
    init_time_neighbors_table(l, p.slices)
    if p.Bfield:
        init_peirls_phases(p, l)
        if dense_hoppings:
            init_hopping_matrix_exp_Bfield(p, l)
        if p.chkr:
            init_checkerboard_matrices_Bfield(p, l)
    else:
        if dense_hoppings:
            init_hopping_matrix_exp(p, l)
        if p.chkr:
            init_checkerboard_matrices(p, l)
    return l
