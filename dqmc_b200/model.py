"""Host-side model setup for the B200 DQMC path: parameters, lattice tables and the sparse
checkerboard factors the device consumes as data.

Mirrors, with the same field names, the parts of the reference's `Params` (src/parameters.jl:4-80),
`Lattice` (src/lattice.jl:1-55), `load_lattice` (:57-67) and `init_checkerboard_matrices[_Bfield]`
(src/hoppings_checkerboard.jl:65-136, 165-270) that the hot path reads.  In a Julia deployment this
file is not needed: the unchanged reference driver builds `mc.l.chkr_*` itself and hands the CSC
arrays to `dqmc_set_operator`.  One-off O(N) host work; not on the GPU path.
"""
import math
import re
from dataclasses import dataclass

import numpy as np
import scipy.sparse as sp


@dataclass
class Params:
    """src/parameters.jl:4-80; XML names in `from_xml` follow set_parameters (:88-190)."""
    L: int = 4
    slices: int = 10
    delta_tau: float = 0.1
    safe_mult: int = 10
    opdim: int = 3
    flv: int = 4
    hoppings: str = "1.0,0.5,-0.5,-1.0"
    mu1: float = -0.5
    mu2: float = -0.5
    lambda_: float = 0.5
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
    thermalization: int = 0
    measurements: int = 0
    write_every_nth: int = 1

    @property
    def beta(self):
        return self.slices * self.delta_tau

    @classmethod
    def from_xml(cls, path):
        """ALPS-style `<PARAMETER name="...">value</PARAMETER>` (src/tools/xml_parameters.jl:24-43)."""
        txt = open(path).read()
        d = {m.group(1).upper(): m.group(2).strip()
             for m in re.finditer(r'<PARAMETER\s+name="([^"]+)"\s*>([^<]*)</PARAMETER>', txt)}
        return cls.from_dict(d)

    @classmethod
    def from_dict(cls, d):
        p = cls()
        tobool = lambda s: s.strip().lower() == "true"
        p.thermalization = int(d["WARMUP"]) // 2 if "WARMUP" in d else int(d["THERMALIZATION"])
        p.measurements = int(d["SWEEPS"]) // 2 if "SWEEPS" in d else int(d["MEASUREMENTS"])
        p.delta_tau = float(d["DELTA_TAU"])
        if "BETA" in d:
            p.slices = int(round(float(d["BETA"]) / p.delta_tau))
        elif "T" in d:
            p.slices = int(round(1.0 / float(d["T"]) / p.delta_tau))
        else:
            p.slices = int(d["SLICES"])
        p.safe_mult = int(d["SAFE_MULT"])
        p.L = int(d["L"])
        if "HOPPINGS" in d:
            p.hoppings = d["HOPPINGS"]
        if "MU" in d:
            p.mu1 = p.mu2 = float(d["MU"])
        if "MU1" in d:
            p.mu1 = float(d["MU1"])
        if "MU2" in d:
            p.mu2 = float(d["MU2"])
        p.lambda_, p.r, p.c, p.u = float(d["LAMBDA"]), float(d["R"]), float(d["C"]), float(d["U"])
        if "OPDIM" in d:
            p.opdim = int(d["OPDIM"])
            p.flv = 4 if p.opdim == 3 else 2
        if "SEED" in d:
            p.seed = int(d["SEED"])
        for key, attr in (("GLOBAL_UPDATES", "global_updates"), ("CHECKERBOARD", "chkr"), ("BFIELD", "Bfield"),
                          ("EDRUN", "edrun")):
            if key in d:
                setattr(p, attr, tobool(d[key]))
        if "BOX_HALF_LENGTH" in d:
            p.box = p.box_global = float(d["BOX_HALF_LENGTH"])
        if "BOX_GLOBAL_HALF_LENGTH" in d:
            p.box_global = float(d["BOX_GLOBAL_HALF_LENGTH"])
        if "GLOBAL_RATE" in d:
            p.global_rate = int(d["GLOBAL_RATE"])
        if "WRITE_EVERY_NTH" in d:
            p.write_every_nth = int(d["WRITE_EVERY_NTH"])
        return p


class Lattice:
    """Square lattice with periodic boundaries; fields named as in src/lattice.jl:1-55 (1-based tables)."""

    def __init__(self, p):
        if p.opdim != 3 or p.flv != 4:
            raise NotImplementedError("only the O(3) model (opdim=3, flv=4) is implemented")
        if not p.chkr or p.L % 2:
            raise NotImplementedError("only the Assaad checkerboard on even L is implemented "
                                      "(CBFalse / CBGeneric are out of scope)")
        L = self.L = p.L
        self.sites = L * L
        self.t = np.array([float(f) for f in p.hoppings.split(",")]).reshape(2, 2, order="F")  # t[hor/ver, flavour]
        y, x = np.meshgrid(np.arange(L), np.arange(L), indexing="ij")
        lin = lambda yy, xx: (yy % L) + L * (xx % L)            # site index, y fastest (lattice.jl:108)
        up, right, down, left = lin(y + 1, x), lin(y, x + 1), lin(y - 1, x), lin(y, x - 1)
        self.neighbors = np.vstack([a.reshape(-1, order="F") for a in (up, right, down, left)]).astype(np.int64) + 1
        s = np.arange(1, p.slices + 1)
        self.time_neighbors = np.vstack([np.where(s == p.slices, 1, s + 1), np.where(s == 1, p.slices, s - 1)])
        self._build_checkerboard(p)

    # -- Assaad four-site checkerboard -------------------------------------------------------------------
    def _plaquettes(self):
        """Corner sites (0-based) of group A and B plaquettes (hoppings_checkerboard.jl:4-18)."""
        L = self.L
        A = np.array([y + L * x for x in range(0, L, 2) for y in range(0, L, 2)])
        nb = self.neighbors - 1
        B = nb[0, nb[1, A]]
        return A, B

    def _phase(self, p, s, f, trg, src):
        """Peierls phase for the hop src -> trg (hoppings.jl:97-172, nearest neighbours), spin s, flavour f."""
        if not p.Bfield:
            return 0.0
        L = self.L
        Bsf = (1 if s == f else -1) * 2 * math.pi / self.sites
        ys, xs, yt, xt = src % L, src // L, trg % L, trg // L   # 0-based; reference x,y are 1-based
        # the reference stores phis[x, y, x', y'] with its first coordinate = column index x, second = row y;
        # after permutedims the matrix index is y + L*x, so "y" below is the fast (vertical) coordinate.
        if xs == xt:
            # vertical hop: zero phase except across the periodic boundary, where the hop L-1 -> 0 (+y) carries
            # -B*L*x and its reverse +B*L*x   (phis[x,y,x,yp] for y == L, hoppings.jl:127-130)
            if ys == L - 1 and yt == 0:
                return -Bsf * L * xs
            if ys == 0 and yt == L - 1:
                return Bsf * L * xs
            return 0.0
        # horizontal hop at row y: +B*y in the +x direction, -B*y in the -x direction (hoppings.jl:124-125)
        return Bsf * ys if (xs + 1) % L == xt else -Bsf * ys

    def _group_factor(self, p, corners, s, f, fac):
        """exp(fac * T_plaquettes) for one (spin, flavour) sector as an N x N sparse matrix (disjoint 4x4 blocks)."""
        nb = self.neighbors - 1
        N = self.sites
        rows, cols, vals = [], [], []
        for c in corners:
            cw = [c, nb[0, c], nb[0, nb[1, c]], nb[1, c]]          # clockwise: corner, up, up-right, right
            T = np.zeros((4, 4), dtype=complex)
            hop = [self.t[1, f], self.t[0, f], self.t[1, f], self.t[0, f]]   # v, h, v, h
            for k in range(4):
                i, j = k, (k + 1) % 4
                T[i, j] += -hop[k] * np.exp(1j * self._phase(p, s, f, cw[i], cw[j]))
                T[j, i] += -hop[k] * np.exp(1j * self._phase(p, s, f, cw[j], cw[i]))
            w, V = np.linalg.eigh(T)
            E = (V * np.exp(fac * w)) @ V.conj().T
            for a in range(4):
                for b in range(4):
                    rows.append(cw[a]); cols.append(cw[b]); vals.append(E[a, b])
        vals = np.array(vals)
        vals[np.abs(vals) < 1e-15] = 0           # rem_eff_zeros! (hoppings_checkerboard_generic.jl:35)
        return sp.csc_matrix((vals, (rows, cols)), shape=(N, N))

    def _build_checkerboard(self, p):
        A, B = self._plaquettes()
        order = [(0, 0), (1, 1), (1, 0), (0, 1)]                 # (spin, flavour) blocks (hoppings_checkerboard.jl:219-227)
        dt = p.delta_tau

        def full(corners, fac):
            m = sp.block_diag([self._group_factor(p, corners, s, f, fac) for s, f in order], format="csc")
            return m if p.Bfield else m.real.tocsc()

        self.chkr_hop_half = [full(A, -0.5 * dt), full(B, -0.5 * dt)]
        self.chkr_hop_half_inv = [full(A, 0.5 * dt), full(B, 0.5 * dt)]
        self.chkr_hop = [full(A, -dt), full(B, -dt)]
        self.chkr_hop_inv = [full(A, dt), full(B, dt)]
        self.chkr_hop_half_dagger = [m.conj().T.tocsc() for m in self.chkr_hop_half]
        self.chkr_hop_dagger = [m.conj().T.tocsc() for m in self.chkr_hop]
        muv = np.tile(np.concatenate([np.full(self.sites, p.mu1), np.full(self.sites, p.mu2)]), p.flv // 2)
        self.chkr_mu = sp.diags(np.exp(dt * muv)).tocsc()          # exp(-dtau * -mu)  (hoppings_checkerboard.jl:126)
        self.chkr_mu_inv = sp.diags(np.exp(-dt * muv)).tocsc()
        self.n_groups = 2
