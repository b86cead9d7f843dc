"""CPU oracle for the dqmc hot path.  TEST INFRASTRUCTURE ONLY.

A NumPy/SciPy restatement of the reference's (carstenbauer/dqmc, Julia) local-update sweep,
Green's-function wrap and UDT stabilization, function by function, each citing the reference
file:line it follows.  It exists to *check* the CUDA path; nothing in ``dqmc_b200`` may import
it.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may use it.

Parity status: PINNED.  ``tests/test_oracle_golden.py`` checks this restatement against the
reference's own fixtures (``test/data/O3.jld``, ``linalg.jld`` → ``tests/golden/*.npz``) and
the known-answer scalars in ``test/tests_O3.jl``; the RNG-dependent fixtures are reproduced
through a bit-exact emulation of Julia's MersenneTwister (``oracle/dsfmt.py``).

The dense algebra of the reference lives in LAPACK/BLAS inside the Julia distribution (Julia
1.3.1 bundles OpenBLAS 0.3.5; not under /root/reference): ``zgeqp3``/``zungqr`` via
``qr!(A, Val(true))`` (linalg.jl:22,38), ``zgetrf/zgetrs/zgetri`` via ``\\``, ``inv``, ``det``,
``logdet`` (linalg.jl:61, stack.jl:359,383, local_updates.jl:58,82), ``zgemm``.  Here they are
SciPy's LAPACK (OpenBLAS 0.3.30) — same published algorithms.
"""
from .dsfmt import JuliaMT  # noqa: F401
from .model import Params, Lattice, build_model  # noqa: F401
from .dqmc import OracleDQMC  # noqa: F401
