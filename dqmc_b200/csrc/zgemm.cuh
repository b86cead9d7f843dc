// Complex-FP64 GEMM on the FP64 tensor pipe (DMMA.8x8x4), C = alpha*op(A)*op(B) + beta*C, column-major.
// Replaces the reference's BLAS zgemm call sites (stack.jl:291,312,346-367; local_updates.jl:86).
#pragma once
#include "common.cuh"

int zgemm(cudaStream_t stream, int opA, int opB, int M, int N, int K, cplx alpha, const cplx* A, int lda,
          const cplx* B, int ldb, cplx beta, cplx* C, int ldc, int num_sms);
