"""Per-phase cycles of the QR panel kernel (cluster rank 0, thread 0):  python tools/qr_profile.py [L]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dqmc_b200 import DQMC, Params
from dqmc_b200 import lib as _l
L = int(sys.argv[1]) if len(sys.argv) > 1 else 16
mc = DQMC(Params(L=L, slices=40, safe_mult=10, Bfield=False), device=0)
rs = np.random.RandomState(0)
mc.init(rs.rand(3, L * L, 40))
out = np.zeros(16, dtype=np.int64)
mc.lib.dqmc_qr_profile(mc._ctx, 1, None)
ms = mc.bench_kernel(3, 1)
mc.lib.dqmc_qr_profile(mc._ctx, 0, out.ctypes.data_as(_l._I64))
ncol = max(out[5], 1)
print(f"UDT {ms:.2f} ms; panel columns profiled: {ncol}")
for name, v in zip(("A: dots", "reduce+push", "cluster barrier", "C: parameters", "D: update"), out[:5]):
    print(f"  {name:16s} {v/ncol:8.0f} cycles/column")
print(f"  total            {out[:5].sum()/ncol:8.0f} cycles/column")
mc.close()
