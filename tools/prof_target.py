"""Small driver for ncu: sets up the headline-size problem on a short chain and launches each kernel group once
(python tools/prof_target.py [L] [which ...]).  Not a benchmark: numbers printed under a profiler are meaningless."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dqmc_b200 import DQMC, Params  # noqa: E402

L = int(sys.argv[1]) if len(sys.argv) > 1 else 16
which = [int(x) for x in sys.argv[2:]] or [0, 1, 3, 4, 5, 7]
M = 40
mc = DQMC(Params(L=L, slices=M, safe_mult=10, Bfield=False), device=0)
rs = np.random.RandomState(0)
mc.init(rs.rand(3, L * L, M))
mc.set_uniforms(rs.rand(4 * L * L * M))
for _ in range(5):
    mc.propagate()
for w in which:
    if w == 99:      # the column-pivoted QR of dqmc_decompose_udt (TDGF chains)
        n = mc.n
        X = (rs.rand(n, n) + 1j * rs.rand(n, n) - (0.5 + 0.5j)) * np.logspace(10, -30, n)[None, :]
        mc.decompose_udt(X)
        print(w, "qrcp")
        continue
    print(w, mc.bench_kernel(w, 1))
mc.close()
