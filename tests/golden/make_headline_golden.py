"""Generate tests/golden/headline_L{L}_beta40.npz: the CPU oracle's results at the BASELINE headline depth
(O(3), beta = 40, M = 400 slices, safe_mult = 10; L = 16 is configs[3], L = 12 is configs[2]).

Run here (CPU container, ~10 min at L=16):   python tests/golden/make_headline_golden.py 16
The GPU parity tests (tests/test_gpu_headline.py) replay the same seeded inputs through the C ABI and compare.

The oracle's G (16 MiB at L=16) is not stored; each checkpoint keeps
  * ``Gv``     = G @ v          for a fixed seeded complex probe vector v        (n complex)
  * ``sample`` = G[rows][:, cols] for fixed seeded index sets                    (24 x 24 complex)
  * ``gmax``   = max |G|
which pins every element of G statistically (Gv) and a few hundred of them directly (sample).

Phase B (shared uniform stream): from the init state, 20 x {propagate; local_updates} = two safe_mult blocks incl. two
stabilizations: per-update accepted counts and stream positions, the bit pattern of the updated field slices, G
checkpoints after 10 and 20 updates, the boson action.
Phase A (propagation only): from the init state a full up-down sweep of `propagate` (2M calls: 2K stabilizations,
both turn-arounds, 2(M-K)+K wraps); a checkpoint at every slice = 0 or 1 (mod safe_mult), i.e. right after every
stabilization and right before the next one (deepest wrap chain), plus log_det at the turn-arounds.

Reference: src/stack.jl:251-499 (build_stack / propagate / calculate_greens), src/local_updates.jl:1-95,
test/tests_O3.jl:240-263 (the reference's own propagated-vs-fresh check, there at L=4, M=10).
"""
import copy
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from oracle.dqmc import UniformStream  # noqa: E402

M, SM, NS = 400, 10, 24


def probes(n):
    rs = np.random.RandomState(777)
    v = rs.randn(n) + 1j * rs.randn(n)
    rows = np.sort(rs.choice(n, NS, replace=False))
    cols = np.sort(rs.choice(n, NS, replace=False))
    return v, rows, cols


def field_for(L):
    return np.random.RandomState(1600 + L).rand(3, L * L, M)


def stream_for(L, nupd):
    return np.random.RandomState(900 + L).rand(4 * L * L * nupd)


def fingerprint(G, v, rows, cols):
    return G @ v, G[np.ix_(rows, cols)].copy(), float(np.max(np.abs(G)))


def main(L):
    t0 = time.time()
    N = L * L
    n = 4 * N
    v, rows, cols = probes(n)
    field = field_for(L)
    om = oracle.OracleDQMC(oracle.Params(L=L, slices=M, safe_mult=SM, Bfield=False, lam=0.5))
    om.p.all_checks = False
    om.init(field)
    print(f"init done {time.time() - t0:.0f}s", flush=True)
    out = {"L": L, "M": M, "safe_mult": SM}
    gv, smp, gmax = fingerprint(om.greens, v, rows, cols)
    out.update(init_Gv=gv, init_sample=smp, init_gmax=gmax, init_logdet=om.log_det,
               init_state=np.array([om.current_slice + 1, om.direction]))

    # ---- phase B: two safe_mult blocks of {propagate; local_updates} on a shared stream
    ob = copy.deepcopy(om)
    nupd = 2 * SM
    u = stream_for(L, nupd)
    st = UniformStream(u)
    acc, pos, slices = [], [], []
    for k in range(nupd):
        ob.propagate()
        a = ob.local_updates(st)
        acc.append(int(round(a * N)))
        pos.append(st.pos)
        slices.append(ob.current_slice + 1)
        if k + 1 in (SM, nupd):
            gv, smp, gmax = fingerprint(ob.greens, v, rows, cols)
            out[f"lu{k + 1}_Gv"], out[f"lu{k + 1}_sample"], out[f"lu{k + 1}_gmax"] = gv, smp, gmax
    out.update(lu_accepted=np.array(acc), lu_pos=np.array(pos), lu_slices=np.array(slices),
               lu_field=ob.hsfield[:, :, min(slices) - 1:max(slices)].copy(), lu_boson_action=ob.boson_action,
               lu_boson_action0=om.boson_action)
    del ob
    print(f"phase B done {time.time() - t0:.0f}s: accepted {sum(acc)} of {nupd * N}, consumed {pos[-1]}", flush=True)

    # ---- phase A: full up-down sweep of propagate
    ck_step, ck_state, ck_Gv, ck_sample, ck_gmax, ld_step, ld_val = [], [], [], [], [], [], []
    last_ld = om.log_det
    for k in range(2 * M):
        om.propagate()
        s1 = om.current_slice + 1
        if om.log_det != last_ld:
            last_ld = om.log_det
            ld_step.append(k)
            ld_val.append(om.log_det)
        if s1 % SM in (0, 1):
            gv, smp, gmax = fingerprint(om.greens, v, rows, cols)
            ck_step.append(k); ck_state.append((s1, om.direction)); ck_Gv.append(gv); ck_sample.append(smp); ck_gmax.append(gmax)
        if k % 50 == 49:
            print(f"  propagate {k + 1}/{2 * M} {time.time() - t0:.0f}s", flush=True)
    out.update(ck_step=np.array(ck_step), ck_state=np.array(ck_state), ck_Gv=np.array(ck_Gv), ck_sample=np.array(ck_sample),
               ck_gmax=np.array(ck_gmax), ld_step=np.array(ld_step), ld_val=np.array(ld_val))
    path = os.path.join(ROOT, "tests", "golden", f"headline_L{L}_beta40.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path} ({os.path.getsize(path) / 1e6:.2f} MB) in {time.time() - t0:.0f}s; {len(ck_step)} checkpoints")


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 16)
