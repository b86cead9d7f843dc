"""hand-written ZGEMM vs the cuBLAS probe (python tools/zgemm_bench.py [L])"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dqmc_b200 import DQMC, Params
L = int(sys.argv[1]) if len(sys.argv) > 1 else 16
mc = DQMC(Params(L=L, slices=40, safe_mult=10, Bfield=False), device=0)
mc.init(np.random.RandomState(0).rand(3, L * L, 40))
n = mc.n
t1, t2 = mc.bench_kernel(1, 20), mc.bench_kernel(2, 20)
print(f"n={n}: zgemm {t1*1e3:.1f} us = {8*n**3/t1/1e9:.2f} TFLOP/s; cuBLAS {t2*1e3:.1f} us = {8*n**3/t2/1e9:.2f} TFLOP/s; ratio {t2/t1:.3f}")
mc.close()
