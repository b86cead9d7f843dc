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
tot = p[0]
print(f" cycles total {tot}  stage1 {p[1]} ({p[1]/tot:.0%})  stage2(acc) {p[2]} ({p[2]/tot:.0%})  stage2(rej) {p[6]} ({p[6]/tot:.0%})  flush iters {p[3]} ({p[3]/tot:.0%})  flushes {p[4]} accepts {p[5]}")
print(" per-site stage1 %.0f cyc; per-accept stage2 %.0f cyc; per-reject tail %.0f; per-flush-iteration %.0f cyc" % (p[1]/N, p[2]/max(p[5]-p[4],1), p[6]/max(N-p[5],1), p[3]/max(p[4],1)))
print(" role cycles/site: decision %.0f  prep %s  prefetch %s" % (p[8]/N, [int(p[9+k]/N) for k in range(3)], [int(p[12+k]/N) for k in range(4)]))
mc.close()
print(" role-only cycles/site (before the speculative dot products): %s" % [int(p[16+k]/N) for k in range(8)])
print(" per flush: barrier-1 %.0f, tiles %.0f, barrier-2 %.0f cycles" % tuple(p[24+k]/max(p[4],1) for k in range(3)))
