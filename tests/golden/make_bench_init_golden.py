"""Generate tests/golden/bench_init_<config>.npz: the oracle's G after `init!` for bench.py's own chain 0 (Philox field seed
1234, see bench.synthetic_inputs), so that the bench line can carry `checks.g_vs_oracle_rel` at the headline config without
running the CPU oracle inside the timed job.   python tests/golden/make_bench_init_golden.py [L16_beta40]   (~1 min)
Stored: probe vector v, sampled index sets, G @ v, G[rows, cols], max|G|  (reference: src/stack.jl:251-272, 338-369)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
import oracle  # noqa: E402


def main(config):
    cfg = bench.CONFIGS[config]
    L, M, sm = cfg["L"], cfg["slices"], cfg["safe_mult"]
    field, _ = bench.synthetic_inputs(cfg, 0, 0)
    om = oracle.OracleDQMC(oracle.Params(L=L, slices=M, safe_mult=sm, lam=bench.MODEL["lam"], all_checks=False))
    om.init(field)
    n = om.n
    rs = np.random.RandomState(4242)
    v = rs.randn(n) + 1j * rs.randn(n)
    rows = np.sort(rs.choice(n, 32, replace=False))
    cols = np.sort(rs.choice(n, 32, replace=False))
    G = om.greens
    path = os.path.join(ROOT, "tests", "golden", f"bench_init_{config}.npz")
    np.savez_compressed(path, v=v, rows=rows, cols=cols, Gv=G @ v, sample=G[np.ix_(rows, cols)], gmax=float(np.max(np.abs(G))),
                        logdet=om.log_det)
    print("wrote", path)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "L16_beta40")
