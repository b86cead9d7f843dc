"""A/B timing helper (run under different DQMC_* environment switches):  python tools/ab_time.py [L]
prints ms per call of the UDT, calculate_greens, the triangular solve, ZGEMM and one local-update slice."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dqmc_b200 import DQMC, Params
L = int(sys.argv[1]) if len(sys.argv) > 1 else 16
M = 40
mc = DQMC(Params(L=L, slices=M, safe_mult=10, Bfield=False), device=0)
rs = np.random.RandomState(0)
mc.init(rs.rand(3, L * L, M)); mc.set_uniforms(rs.rand(4 * L * L * M))
for _ in range(5): mc.propagate()
print({k: round(mc.bench_kernel(w, 5), 4) for k, w in (("udt", 3), ("greens", 4), ("trsm", 11), ("zgemm", 1), ("lu_slice", 5))})
mc.close()
