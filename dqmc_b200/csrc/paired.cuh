// Layout kernels of the HALF-MATRIX stabilization path ("paired factorization").  Every n x n matrix on
// that path has the antiunitary flavour symmetry  S = [[A, B], [-conj(B), conj(A)]]  (h = n/2), i.e. column c+h is the partner
// (-conj(bottom); conj(top)) of column c, so a matrix is fully described by its LEFT HALF (n x h).  The paired QR works on
// left halves with PAIR-INTERLEAVED rows ("Lint": row 2i = natural row i, row 2i+1 = natural row i+h).  Nothing here has a
// counterpart in the reference (it factors full matrices with LAPACK, linalg.jl:20-39, stack.jl:338-369).
#pragma once
#include "common.cuh"

// out[:, c] (Lint, n x h) = in[:, perm[c]] (natural rows, left-half column), c < h
int gather_interleave_cols(cudaStream_t st, const cplx* in, int ldi, int n, const int* perm, cplx* out, int ldo, int num_sms);
// out (Lint, n x h) = left half of the identity
int set_identity_lint(cudaStream_t st, cplx* out, int ldo, int n, int num_sms);
// U (n x n, natural) = Q from QHL = left half of Q^H (Lint, as left by qr_factor_paired applied to the identity)
int build_U_paired(cudaStream_t st, const cplx* QHL, int ldq, int n, cplx* U, int ldu, int num_sms);
// T (n x n, natural) = D^-1 R P^T from R_L (Lint, sorted left-half columns), dabs (h moduli), perm (h)
int build_T_paired(cudaStream_t st, const cplx* RL, int ldr, int n, const double* dabs, const int* perm, cplx* T, int ldt, int num_sms);
// right half of M from its left half
int mirror_right_half(cudaStream_t st, cplx* M, int ldm, int n, int num_sms);
// inner_L (Lint) = M1_L / (Dlp Drp^T) + (Dlm Drm^T) .* M2_L ;  rhs_L (Lint) = left half of Ul^H / Dlp ;  drp_inv = 1 / max(Dr, 1)
int loh_assemble_paired(cudaStream_t st, int n, const cplx* M1L, const cplx* M2L, const double* Dl, const double* Dr, const cplx* Ul,
                        cplx* innerL, cplx* rhsL, double* drp_inv, int num_sms);
// Rfull (n x n, rows AND columns pair-interleaved: column 2c = R_L[:, c], column 2c+1 = its partner) from R_L (Lint)
int expand_R_paired(cudaStream_t st, const cplx* RL, int ldr, int n, cplx* Rfull, int ldf, int num_sms);
// out (natural rows, n x h) = rowscale .* in (Lint)
int uninterleave_rows(cudaStream_t st, const cplx* in, int ldi, int n, const double* rowscale, cplx* out, int ldo, int num_sms);
