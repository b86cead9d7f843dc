// Persistent local-update kernel: the serial Metropolis loop over the sites of one time slice
// (local_updates.jl:1-39) with delayed rank-4 Woodbury updates of the equal-time Green's function
// (calc_detratio :42-59, update_greens! :61-95).
#pragma once
#include "common.cuh"

struct LUArgs {
  int n, nsites, nslices, slice;   // slice 0-based
  int kmax;                        // accepted updates batched before G is flushed with one GEMM (block kernel: flush at the first
                                   // block end with more than kmax pending)
  int rpc;                         // rows of A / columns of B owned by each CTA
  int edrun;
  double box, dtau, lam_dtau, inv_dtau_c2, r, u;
  cplx* G;                         // n x n, column-major
  cplx* At;                        // [4*kmax x n]  At[p + ldk*row] = A[row,p]
  cplx* Bm;                        // [4*kmax x n]  Bm[p + ldk*col] = B[p,col]
  double* hs;                      // hsfield [3, nsites, nslices]
  const int* nbr;                  // [4, nsites] 0-based spatial neighbours
  const double* unif;              // uniform stream
  long long nunif;
  long long* pos;                  // in/out: next unread position in unif
  long long* accepted;             // in/out: accumulated accepted proposals
  double* dS;                      // in/out: accumulated -log(exp(-dS)) of accepted proposals
  int* flags;                      // [0] stream exhausted, [1] non-real determinant ratios seen
  unsigned int* bar;               // grid barrier state: [0..1] {count, generation} (bar_mode 0); [2..3] two monotonic counters (bar_mode 1)
  int sym;                         // 1: the flush exploits G = [[A, B], [-conj(B), conj(A)]] (upper half computed, lower half mirrored)
  int bar_mode, bar_parity;        // bar_mode 1: this launch counts on bar[2 + bar_parity] and clears bar[2 + (1 - bar_parity)]
  long long* prof;                 // optional [16] cycle counters of CTA 0 (nullptr = off)
};

int launch_local_updates(cudaStream_t st, const LUArgs& a, int grid);   // per-site lookahead kernel (local_updates.cu)
// block-lookahead kernel (local_updates_blk.cu): rpc <= 16, kmax <= 16; At / Bm must hold 2 x 4 (kmax + 8) x n elements
#define LU_BLOCK_SITES 8
int launch_lu_block(cudaStream_t st, const LUArgs& a, int grid);
size_t lu_block_smem(const LUArgs& a);
int local_updates_grid(int n, int num_sms, int* rpc);
size_t local_updates_smem(const LUArgs& a);
