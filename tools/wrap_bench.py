"""wrap / B-chain timing (python tools/wrap_bench.py [L])"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dqmc_b200 import DQMC, Params
L = int(sys.argv[1]) if len(sys.argv) > 1 else 16
mc = DQMC(Params(L=L, slices=40, safe_mult=10, Bfield=False), device=0)
mc.init(np.random.RandomState(0).rand(3, L * L, 40))
print(f"L={L}: wrap {mc.bench_kernel(0, 50)*1e3:.1f} us, B chain (10 slices) {mc.bench_kernel(7, 20)*1e3:.1f} us, copy G {mc.bench_kernel(6, 50)*1e3:.1f} us")
mc.close()
