"""Per-phase cycles of the PAIRED QR panel kernel (thread 0 of a chosen cluster rank):  python tools/qr_profile_paired.py [L]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dqmc_b200 import DQMC, Params
from dqmc_b200 import lib as _l
L = int(sys.argv[1]) if len(sys.argv) > 1 else 16
mc = DQMC(Params(L=L, slices=40, safe_mult=10, Bfield=False), device=0)
rs = np.random.RandomState(0)
mc.init(rs.rand(3, L * L, 40))
names = ("A: dots", "reduce+push", "T columns", "wait exchange", "C: parameters", "D: update", "epilogue")
for rank in (0, 1, 4, 7):
    out = np.zeros(16, dtype=np.int64)
    mc.lib.dqmc_qr_profile(mc._ctx, 1 + rank, None)
    ms = mc.bench_kernel(16, 1)
    mc.lib.dqmc_qr_profile(mc._ctx, 0, out.ctypes.data_as(_l._I64))
    nst = max(out[7], 1)
    print(f"paired panel chain {ms:.3f} ms (profile instantiation); cluster rank {rank}; pair-steps profiled: {nst}")
    for name, v in zip(names, out[:7]):
        print(f"  {name:16s} {v/nst:8.0f} cycles/step")
    print(f"  total            {out[:7].sum()/nst:8.0f} cycles/step")
    if rank == 1:
        for name, v in zip(("C1: sums of the exchange", "C2: shuffles + row pair", "C3: rsqrt / rcp chain", "C4: uc, pc, stores", "C5: barrier + reload"), out[8:13]):
            print(f"    {name:28s} {v/nst:8.0f} cycles/step")
mc.close()
