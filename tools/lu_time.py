"""us per slice of the production (un-profiled) local-update kernel:  python tools/lu_time.py [L] [delay]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dqmc_b200 import DQMC, Params
L = int(sys.argv[1]) if len(sys.argv) > 1 else 16
delay = int(sys.argv[2]) if len(sys.argv) > 2 else 0
M = 40
mc = DQMC(Params(L=L, slices=M, safe_mult=10, Bfield=False), device=0, delay=delay)
rs = np.random.RandomState(0)
mc.init(rs.rand(3, L * L, M)); mc.set_uniforms(rs.rand(4 * L * L * M))
for _ in range(5): mc.propagate()
ms = mc.bench_kernel(5, 10)
print(f"L={L} delay={delay}: {ms*1e3:.1f} us/slice  {ms*1e3/(L*L):.3f} us/proposal")
mc.close()
