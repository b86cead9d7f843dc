"""Time measure_tdgfs! on the device (python tools/tdgf_time.py [L] [M])."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dqmc_b200 import DQMC, Params
L = int(sys.argv[1]) if len(sys.argv) > 1 else 16
M = int(sys.argv[2]) if len(sys.argv) > 2 else 400
mc = DQMC(Params(L=L, slices=M, safe_mult=10, Bfield=False), device=0)
mc.init(np.random.RandomState(0).rand(3, L * L, M))
mc.measure_tdgfs()          # allocates on first use
mc.sync(); t0 = time.perf_counter()
mc.measure_tdgfs()
mc.sync(); dt = time.perf_counter() - t0
n = mc.n
g1, g2 = mc.Gt0(1), mc.G0t(1)
print(f"L={L} M={M}: measure_tdgfs {dt*1e3:.0f} ms; |Gt0[1]-G0t[1]-1| = {np.abs(g1 - g2 - np.eye(n)).max():.2e}; "
      f"device memory for Gt0+G0t+stacks {(2*M + 8*(M//10)) * n*n*16/2**30:.1f} GiB")
mc.close()
