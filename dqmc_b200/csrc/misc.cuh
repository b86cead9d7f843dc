// Small elementwise / reduction kernels around the dense algebra of the stabilization path.
#pragma once
#include "common.cuh"

int colnorm2(cudaStream_t st, const cplx* A, int lda, int n, double* out);
int argsort_desc(cudaStream_t st, const double* key, int n, int* perm);           // n <= 2048
int gather_cols(cudaStream_t st, const cplx* in, int ldi, int n, const int* perm, cplx* out, int ldo, int num_sms);
// T[i, perm[j]] = (i <= j) ? R[i,j] / dabs[i] : 0          (linalg.jl:34-38: T = D^-1 R P^T)
int build_T(cudaStream_t st, const cplx* R, int ldr, int n, const double* dabs, const int* perm, cplx* T, int ldt, int num_sms);
// inner = M1 / (Dlp Drp^T) + (Dlm Drm^T) .* M2 ;  rhs = Ul^H / Dlp   (Dp = max(D,1), Dm = min(D,1))
int loh_assemble(cudaStream_t st, int n, const cplx* M1, const cplx* M2, const double* Dl, const double* Dr,
                 const cplx* Ul, cplx* inner, cplx* rhs, double* drp_inv, int num_sms);
// out[0] = max |A - B|
int max_abs_diff(cudaStream_t st, const cplx* A, const cplx* B, size_t count, double* out, int num_sms);
// out[0] = -(sum log max(Dl,1) + sum log max(Dr,1) + sum log dabs)
int logdet_from_factors(cudaStream_t st, int n, const double* Dl, const double* Dr, const double* dabs, double* out);
int set_identity(cudaStream_t st, cplx* Q, int ldq, int n, int num_sms);
int fill_ones(cudaStream_t st, double* d, int n);
// out = alpha * (f_a .* op_a(A) + f_b .* op_b(B)): see misc.cu (time-displaced Green's functions)
struct EwTerm {
  const cplx* M;       // nullptr: term absent
  int trans;           // 0: as is, 1: conjugate transpose
  const double* rd;    // row factor source (or nullptr)
  int rmode;           // 0: 1, 1: D, 2: 1/max(D,1), 3: min(D,1)
  const double* cd;    // column factor source
  int cmode;
};
int ew_combine(cudaStream_t st, int n, EwTerm a, EwTerm b, double alpha, cplx* out, int num_sms);
int mirror_lower_half(cudaStream_t st, cplx* G, int n, int num_sms);   // G = [[A, B], [-conj(B), conj(A)]] from its upper half
// out2[0] = max deviation of G from that symmetry, out2[1] = max |G|
int sym_violation(cudaStream_t st, const cplx* G, int n, double* out2, int num_sms);
