"""Cycle breakdown of the persistent local-update kernel (CTA 0), for tuning:  python tools/lu_profile.py [L] [delay]"""
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
mc.lu_profile(True)
ms = mc.bench_kernel(5, 1)
p = mc.lu_profile(True, read=True)
N = L * L
print(f"L={L} delay={delay}: {ms*1e3:.0f} us/slice  {ms*1e3/N:.2f} us/proposal")
import os as _os
if _os.environ.get("DQMC_LU_KERNEL") == "site":
    tot = p[0]
    print(f" [site kernel] cycles total {tot}  stage1 {p[1]} ({p[1]/tot:.0%})  stage2(acc) {p[2]} ({p[2]/tot:.0%})  stage2(rej) {p[6]} ({p[6]/tot:.0%})  flush iters {p[3]} ({p[3]/tot:.0%})  flushes {p[4]} accepts {p[5]}")
else:
    tot = p[0]; nblk = (N + 7) // 8
    names = ("gather", "form", "site loops", "-", "flush")
    print(" [block kernel] cycles total %d: " % tot + "  ".join(f"{nm} {p[1+k]} ({p[1+k]/tot:.0%})" for k, nm in enumerate(names)) + f"  flushes {p[6]} accepts {p[7]}")
    print(" per block: gather %.0f, form %.0f; per site: %.0f; per flush: %.0f cycles" %
          (p[1]/nblk, p[2]/nblk, p[3]/N, p[5]/max(p[6],1)))
    nf = max(p[6], 1)
    print(" per flush: barrier-1 %.0f, staging+DMMA %.0f, G read-modify-write %.0f, barrier-2 %.0f, re-arm %.0f" % tuple(p[8+k]/nf for k in range(5)))
mc.close()
