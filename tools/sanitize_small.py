"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / synccheck): L=4, M=20, one up-down sweep + TDGF + QRCP UDT.
    compute-sanitizer --tool racecheck python tools/sanitize_small.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dqmc_b200 import DQMC, Params, UniformStream
L, M = 4, 20
rs = np.random.RandomState(3)
mc = DQMC(Params(L=L, slices=M, safe_mult=10, Bfield=False), device=0)
mc.init(rs.rand(3, L * L, M))
st = UniformStream(rs.rand(4 * L * L * 2 * M))
nsw = int(os.environ.get("SAN_UPDATES", 2 * M))
for _ in range(nsw):
    mc.update(st)
print("accepted fraction", mc.acc_rate / nsw, "consumed", st.consumed)
if os.environ.get("SAN_TDGF", "1") == "1":
    mc.measure_tdgfs()
    g = mc.Gt0(1) - mc.G0t(1)
    print("tdgf identity err", float(np.max(np.abs(g - np.eye(mc.n)))))
    mc.deallocate_tdgfs_stacks()
mc.close()
print("done")
