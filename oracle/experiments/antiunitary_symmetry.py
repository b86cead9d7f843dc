"""Test infrastructure (oracle side): search for an antiunitary flavour symmetry of the equal-time Green's function.

For every signed permutation J of the four flavours, test  (J x 1_N) conj(G) (J x 1_N)^T == G  on the oracle's G
(L=4, beta=2, random field), with the magnetic flux off and on.  Result (2026-10, this repo's oracle):
exactly J = +-[[0, 1_2], [-1_2, 0]] passes (residual 1e-15) in both cases, i.e. with the flavour blocks (1,2 | 3,4)

    G = [[A, B], [-conj(B), conj(A)]]            (A, B: 2N x 2N)

so half of every matrix of the path determines the other half (and every determinant ratio is real).  Not used by the
product yet; DESIGN.md section 6 lists it as the largest remaining lever.

    python -m oracle.experiments.antiunitary_symmetry
"""
import itertools

import numpy as np

import oracle


def main():
    L, M = 4, 20
    N = L * L
    for bfield in (False, True):
        om = oracle.OracleDQMC(oracle.Params(L=L, slices=M, safe_mult=10, Bfield=bfield))
        om.init(np.random.RandomState(1).rand(3, N, M))
        G = om.greens
        found = []
        for perm in itertools.permutations(range(4)):
            for signs in itertools.product([1, -1], repeat=4):
                J = np.zeros((4, 4))
                for i, (p, sg) in enumerate(zip(perm, signs)):
                    J[i, p] = sg
                U = np.kron(J, np.eye(N))
                err = np.abs(U @ G.conj() @ U.T - G).max()
                if err < 1e-10:
                    found.append((perm, signs, float(err)))
        print(f"Bfield={bfield}: {len(found)} signed flavour permutations J with J conj(G) J^T = G")
        for f in found:
            print("   ", f)
        A, B = G[:2 * N, :2 * N], G[:2 * N, 2 * N:]
        print("    |G21 + conj(B)| =", np.abs(G[2 * N:, :2 * N] + B.conj()).max(), " |G22 - conj(A)| =", np.abs(G[2 * N:, 2 * N:] - A.conj()).max())


if __name__ == "__main__":
    main()
