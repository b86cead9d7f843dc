// Blocked Householder QR (compact WY), explicit-Q formation, and a fused multi-RHS triangular solve,
// all complex FP64.  Replaces the reference's LAPACK call sites on the stabilization path:
// zgeqp3 + zungqr inside decompose_udt! (linalg.jl:20-39) and the LU solve / inverse of
// calculate_greens (stack.jl:355-361, linalg.jl:61).  Pivoting is a one-off column sort by norm
// before the factorization ("stabilization algebra").
#pragma once
#include "common.cuh"

#define QR_NB 32      // panel width
#define QR_CL 8       // CTAs per thread-block cluster in the panel factorization

struct QrAsync {      // second stream + two events for the panel lookahead (nullptr = everything on one stream)
  cudaStream_t st2;
  cudaEvent_t eA, eB;
};

// Householder QR of the n x n matrix A (in place: R in the upper triangle, V below, tau, |R_ii| in dabs).
// tfac receives ceil(n/32) compact-WY T factors (32x32, column-major, ld 32).  If rhs != nullptr,
// Q^H is applied to the n x nrhs matrix rhs as the factorization proceeds.
int qr_factor(cudaStream_t st, cplx* A, int lda, int n, cplx* tau, double* dabs, cplx* tfac,
              cplx* rhs, int ldr, int nrhs, int num_sms, const QrAsync* as);
// Q (n x n, explicit) from the output of qr_factor.
int qr_form_q(cudaStream_t st, const cplx* A, int lda, int n, const cplx* tfac, cplx* Q, int ldq, int num_sms);
// X = R^{-1} Y in place in Y (n x nrhs); R = upper triangle of A.  work: ceil(n/32)*1024 cplx.
// quat = 1: R is a QUATERNION upper triangle (2x2 diagonal blocks), as produced by qr_factor_paired + expand_R_paired.
int trsm_upper(cudaStream_t st, const cplx* A, int lda, int n, cplx* Y, int ldy, int nrhs, cplx* work,
               const double* rowscale, int num_sms, int quat = 0);
// Householder QR of a matrix with the antiunitary flavour symmetry [[A, B], [-conj(B), conj(A)]], held as its left half with
// pair-interleaved rows (AL: n x n/2): n/2 paired steps (qr.cu, "PAIRED panel factorization").  In place: the quaternion upper
// triangle R_L with exact zeros below; V (n x n scratch) receives the explicit reflector blocks; dabs[0:n/2] and dabs[n/2:n] =
// moduli of the quaternion diagonal; rhs (n x nrhs, pair-interleaved rows) <- Q^H rhs.  n must be a multiple of 32.
int qr_factor_paired(cudaStream_t st, cplx* AL, int lda, int n, cplx* V, int ldv, double* dabs, cplx* tfac, cplx* rhs, int ldr,
                     int nrhs, const QrAsync* as);
// bench hook: the panel factorizations of qr_factor without any trailing update
int qr_panels_only(cudaStream_t st, cplx* A, int lda, int n, cplx* tau, double* dabs, cplx* tfac);

// Column-pivoted Householder QR in one cooperative kernel (qrcp.cu): the reference's decompose_udt! proper (zgeqp3).
// A (n x n, destroyed) -> QH = Q^H, dabs = |R_jj| (descending), T = D^-1 R P^T, perm / pos = pivot order and its inverse.
// vn, rdiag: n doubles of scratch each; bar: one unsigned int of scratch.
int qrcp_udt(cudaStream_t st, cplx* A, int lda, int n, cplx* QH, int ldq, cplx* T, int ldt, double* dabs, double* vn, double* rdiag,
             int* perm, int* pos, unsigned int* bar, int num_sms);
// bench hook: the paired panel factorizations alone
int qr_panels_only_paired(cudaStream_t st, cplx* AL, int lda, int n, cplx* V, int ldv, double* dabs, cplx* tfac);
