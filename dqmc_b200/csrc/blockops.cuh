// Quad-block sparse operators (checkerboard plaquette exponentials, interaction exponentials)
// applied to column panels (left multiplication) or row panels (right multiplication) of a dense
// complex-FP64 matrix staged in shared memory.
//
// Replaces the reference's per-factor `mul!(tmp, SparseFactor, M); M .= tmp` passes
// (slice_matrices.jl:101-226, linalg.jl:45-59): one kernel applies a whole chain of factors
// to a panel with one global read and one global write (32*n^2 algorithmic bytes per side).
#pragma once
#include "common.cuh"

#define DQMC_MAX_CHAIN 44

struct QuadOp {        // n x n operator = n/4 disjoint 4-index blocks with a dense 4x4 each
  int nblk;
  int* idx;            // [4*nblk] 0-based indices
  cplx* val;           // [16*nblk] row-major 4x4 per block
};

struct ChainStep {
  int kind;            // 0: stored QuadOp, 1: interaction exponential e^{-sign*dtau*V(slice)}
  int mode;            // OP_N / OP_T / OP_C / OP_J applied to the 4x4 blocks
  int nblk;
  int slice;           // 0-based time slice (kind 1)
  double sign;         // +1 / -1 (kind 1)
  const int* idx;
  const cplx* val;
};

struct Chain {
  int nsteps;
  ChainStep s[DQMC_MAX_CHAIN];
};

int launch_apply_chain(bool rows, cplx* mat, int n, int ld, const Chain& chain, const double* hsfield,
                       int nsites, double lam_dtau, const double* colscale, double* colnorm2,
                       int num_sms, cudaStream_t stream, int nvtot = 0, int mirror = 0);
// nvtot: vectors to process (0 = all n).  mirror: the matrix has the antiunitary flavour symmetry [[A, B], [-conj(B), conj(A)]]: only the
// left-half columns (upper-half rows) are processed and the other half is written as their mirror image
