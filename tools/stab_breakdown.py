"""Time the pieces of one stabilization (python tools/stab_breakdown.py [L]): not a benchmark line, a tuning aid."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dqmc_b200 import DQMC, Params
L = int(sys.argv[1]) if len(sys.argv) > 1 else 16
mc = DQMC(Params(L=L, slices=40, safe_mult=10, Bfield=False), device=0)
rs = np.random.RandomState(0)
mc.init(rs.rand(3, L * L, 40))
names = {1: "zgemm", 7: "B chain (10 slices)", 3: "UDT (sort+QR+Q+T)", 8: "  QR factor", 10: "  panel chain only", 9: "  form Q",
         12: "QR + Q^H on n rhs", 11: "trsm (n rhs)", 4: "calculate_greens", 15: "paired QR + n/2 rhs", 16: "  paired panel chain only",
         17: "paired QR, no rhs"}
for w in (1, 7, 3, 8, 10, 9, 12, 11, 4, 17, 16, 15):
    print(f"{names[w]:24s} {mc.bench_kernel(w, 5):8.3f} ms")
mc.close()
