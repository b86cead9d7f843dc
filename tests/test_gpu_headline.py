"""Parity at the BASELINE headline depth: O(3), beta = 40 (M = 400, safe_mult = 10), L = 12 (configs[2]) and L = 16 (configs[3]).

The oracle side was generated once on the CPU by tests/golden/make_headline_golden.py (the committed script; ~10 min at
L=16) into tests/golden/headline_L{L}_beta40.npz; here the CUDA path replays the same seeded inputs through the C ABI.
Tolerance (north star): G within 1e-10 relative to max|G|; field / accept sequence / consumed uniforms bit-exact.

What is pinned (reference: src/stack.jl:338-369, 391-499; src/local_updates.jl:1-95; test/tests_O3.jl:240-263):
  * G and log_det after `init!` (build_stack + first propagate) at D spanning ~100 decades,
  * G right after EVERY stabilization of a full up-down sweep (both directions, both turn-arounds) and right before the
    next one (deepest wrap chain), log_det at the turn-arounds,
  * two safe_mult blocks of {propagate; local_updates} on a shared uniform stream.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle
from tests.golden.make_headline_golden import M, SM, field_for, probes, stream_for
from tests.helpers import maxabs

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL = 1e-10


def _golden(L):
    return dict(np.load(os.path.join(ROOT, "tests", "golden", f"headline_L{L}_beta40.npz")))


def _mk(L, all_checks=False):
    from dqmc_b200 import DQMC, Params
    return DQMC(Params(L=L, slices=M, safe_mult=SM, Bfield=False, lambda_=0.5, all_checks=all_checks), device=0)


def _errs(G, Gv, sample, gmax, v, rows, cols):
    """(direct error on the sampled entries, statistical error of all entries through the probe vector), both relative to
    max|G|.  (dG v)_i is a sum of n terms of random phase, |v_j|^2 = 2 on average, hence the sqrt(2n)."""
    n = G.shape[0]
    e_s = maxabs(G[np.ix_(rows, cols)], sample) / gmax
    e_v = maxabs(G @ v, Gv) / (gmax * np.sqrt(2.0 * n))
    return e_s, e_v


@pytest.mark.parametrize("L", [12, 16])
def test_headline_propagation_vs_oracle(L):
    g = _golden(L)
    mc = _mk(L, all_checks=True)
    v, rows, cols = probes(mc.n)
    mc.init(field_for(L))
    assert (mc.current_slice, mc.direction) == tuple(g["init_state"])
    e_s, e_v = _errs(mc.greens, g["init_Gv"], g["init_sample"], float(g["init_gmax"]), v, rows, cols)
    e_ld = abs(mc.log_det - float(g["init_logdet"])) / abs(float(g["init_logdet"]))
    print(f"\nL={L} beta=40 init: |dG|/max|G| sample {e_s:.2e}, probe {e_v:.2e}; log_det rel {e_ld:.2e}")
    assert e_s < TOL and e_v < TOL and e_ld < TOL
    ck = {int(k): i for i, k in enumerate(g["ck_step"])}
    ld = {int(k): float(x) for k, x in zip(g["ld_step"], g["ld_val"])}
    worst_s = worst_v = worst_ld = 0.0
    worst_at = None
    for k in range(2 * M):
        s, d = mc.propagate()
        if k in ck:
            i = ck[k]
            assert (s, d) == tuple(g["ck_state"][i])
            e_s, e_v = _errs(mc.greens, g["ck_Gv"][i], g["ck_sample"][i], float(g["ck_gmax"][i]), v, rows, cols)
            if max(e_s, e_v) > max(worst_s, worst_v):
                worst_at = (s, d)
            worst_s, worst_v = max(worst_s, e_s), max(worst_v, e_v)
        if k in ld:
            worst_ld = max(worst_ld, abs(mc.log_det - ld[k]) / abs(ld[k]))
    err, _ = mc.checks()
    print(f"L={L} beta=40 up-down sweep of propagate, {len(ck)} checkpoints: worst |dG|/max|G| sample {worst_s:.2e}, "
          f"probe {worst_v:.2e} (at slice,dir {worst_at}); log_det rel {worst_ld:.2e}; wrapped-vs-fresh {err:.2e}")
    assert worst_s < TOL and worst_v < TOL and worst_ld < TOL
    assert err < 1e-9
    mc.close()


@pytest.mark.parametrize("L", [12, 16])
def test_headline_local_updates_shared_stream(L):
    from dqmc_b200 import UniformStream
    g = _golden(L)
    mc = _mk(L)
    N = L * L
    v, rows, cols = probes(mc.n)
    mc.init(field_for(L))
    assert np.isclose(mc.boson_action, float(g["lu_boson_action0"]), rtol=1e-13)
    nupd = len(g["lu_accepted"])
    st = UniformStream(stream_for(L, nupd))
    for k in range(nupd):
        s, _ = mc.propagate()
        acc = mc.local_updates(st)
        assert s == g["lu_slices"][k]
        assert round(acc * N) == g["lu_accepted"][k], f"update {k}: accepted {round(acc * N)} vs oracle {g['lu_accepted'][k]}"
        assert st.consumed == g["lu_pos"][k], f"update {k}"
        if k + 1 in (SM, nupd):
            e_s, e_v = _errs(mc.greens, g[f"lu{k + 1}_Gv"], g[f"lu{k + 1}_sample"], float(g[f"lu{k + 1}_gmax"]), v, rows, cols)
            print(f"\nL={L} beta=40 after {k + 1} x (propagate; local_updates): |dG|/max|G| sample {e_s:.2e}, probe {e_v:.2e}")
            assert e_s < TOL and e_v < TOL
    lo, hi = int(g["lu_slices"].min()), int(g["lu_slices"].max())
    assert np.array_equal(mc.hsfield[:, :, lo - 1:hi], g["lu_field"])      # bit-identical field
    assert np.isclose(mc.boson_action, float(g["lu_boson_action"]), rtol=1e-12)
    mc.close()


def test_tdgfs_beta40_vs_oracle():
    # measure_tdgfs! (fermion_measurements.jl:1343-1407) at beta = 40 (L = 8, M = 400): every slice of G(tau,0), G(0,tau)
    # against the oracle run live (n = 256: seconds); D of the B chains spans > 60 decades here
    from dqmc_b200 import DQMC, Params
    L = 8
    mc = DQMC(Params(L=L, slices=M, safe_mult=SM, Bfield=True, lambda_=0.5, all_checks=False), device=0)
    om = oracle.OracleDQMC(oracle.Params(L=L, slices=M, safe_mult=SM, Bfield=True, lam=0.5))
    field = np.random.RandomState(23).rand(3, L * L, M)
    mc.init(field)
    om.init(field)
    mc.measure_tdgfs()
    Gt0, G0t = om.measure_tdgfs()
    worst = 0.0
    for tau in range(M):
        for got, ref in ((mc.Gt0(tau + 1), Gt0[tau]), (mc.G0t(tau + 1), G0t[tau])):
            worst = max(worst, maxabs(got, ref) / max(1.0, np.abs(ref).max()))
    print(f"\nTDGF L=8 beta=40: worst |dG(tau)|/max(1,|G|) over {M} slices {worst:.2e}")
    assert worst < TOL
    mc.deallocate_tdgfs_stacks()
    mc.close()


_AB_SCRIPT = r"""
import sys, numpy as np
sys.path.insert(0, {root!r})
from dqmc_b200 import DQMC, Params, UniformStream
L, M = 12, 40
rs = np.random.RandomState(31)
field = rs.rand(3, L * L, M)
u = rs.rand(4 * L * L * 2 * M)
mc = DQMC(Params(L=L, slices=M, safe_mult=10, Bfield=False, all_checks=False), device=0)
mc.init(field)
st = UniformStream(u)
nacc, consumed = mc.sweep(st, nupdates=2 * M)
np.savez({out!r}, G=mc.greens, h=mc.hsfield, nacc=nacc, consumed=consumed, ld=mc.log_det)
"""


def test_symmetry_and_3m_switches_ab(tmp_path):
    # the default path (block-lookahead local updates, antiunitary-symmetric flush, half product in calculate_greens, 3M complex
    # products, half-matrix stabilization with the paired Householder QR) against (a) the plain one (DQMC_LU_SYM=0 DQMC_GREENS_SYM=0
    # DQMC_ZGEMM_3M=0 DQMC_PAIRED=0: full matrices, sort-once unpaired QR), (b) the per-site-lookahead local-update kernel
    # (DQMC_LU_KERNEL=site) and (c) full-matrix stabilization alone (DQMC_PAIRED=0) on a full up-down sweep at n = 576: same
    # decisions, G to 1e-12
    switches = ("DQMC_LU_SYM", "DQMC_GREENS_SYM", "DQMC_ZGEMM_3M", "DQMC_LU_KERNEL", "DQMC_PAIRED", "DQMC_LARFB_NARROW")
    res = {}
    for tag, env in (("default", {}), ("plain", {"DQMC_LU_SYM": "0", "DQMC_GREENS_SYM": "0", "DQMC_ZGEMM_3M": "0", "DQMC_PAIRED": "0"}),
                     ("site_kernel", {"DQMC_LU_KERNEL": "site"}), ("unpaired", {"DQMC_PAIRED": "0"})):
        out = str(tmp_path / f"{tag}.npz")
        e = dict(os.environ)
        for k in switches:
            e.pop(k, None)
        e.update(env)
        subprocess.run([sys.executable, "-c", _AB_SCRIPT.format(root=ROOT, out=out)], check=True, env=e, timeout=900)
        res[tag] = dict(np.load(out))
    a = res["default"]
    for tag in ("plain", "site_kernel", "unpaired"):
        b = res[tag]
        assert int(a["nacc"]) == int(b["nacc"]) and int(a["consumed"]) == int(b["consumed"]), tag
        assert np.array_equal(a["h"], b["h"]), tag
        d = maxabs(a["G"], b["G"]) / np.abs(b["G"]).max()
        print(f"\nA/B default vs {tag}: |dG|/max|G| = {d:.2e}")
        assert d < 1e-12, tag
        assert np.isclose(float(a["ld"]), float(b["ld"]), rtol=1e-12), tag
