/*
 * libdqmc_b200 — C ABI of the B200-native DQMC hot path (local-update sweep, Green's-function wrap,
 * UDT stabilization) for the O(3) spin-fermion model.
 *
 * The reference (carstenbauer/dqmc, pure Julia) has no FFI: its seam is multiple dispatch on
 * AbstractDQMC{C<:Checkerboard} (src/dqmc_framework.jl:4-13).  Every entry point below names the reference
 * function it replaces; INTEGRATION.md shows the Julia `ccall` methods a maintainer would add.
 *
 * Conventions: plain pointers and sizes only; complex arrays are interleaved (re,im) doubles, column-major
 * (bit-identical to Julia's Array{ComplexF64}); `site` and `slice` arguments are 1-based like the reference;
 * every call returns 0 on success and a negative code on error (message: dqmc_last_error).  A dqmc_ctx is
 * bound to one device and one stream and is not thread-safe (the reference driver is single-threaded,
 * app/dqmc.jl:29-37).  There is no CPU fallback: without a CUDA device dqmc_create fails.
 */
#ifndef DQMC_B200_H
#define DQMC_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct dqmc_ctx dqmc_ctx;

/* POD mirror of the fields of `Params` (src/parameters.jl:4-80) and `Lattice` (src/lattice.jl:1-55) the path reads. */
typedef struct {
  int32_t L;            /* linear size; sites = L*L (lattice.jl:3-4) */
  int32_t flv;          /* 4 (O(3)); flv = 2 models are not implemented */
  int32_t opdim;        /* 3 */
  int32_t slices;       /* M = beta / delta_tau (parameters.jl:102-110) */
  int32_t safe_mult;    /* slices between stabilizations (parameters.jl:111) */
  int32_t edrun;        /* only the mass term of the boson action (action.jl:94) */
  int32_t all_checks;   /* compare wrapped vs fresh G at each stabilization (stack.jl:416-429, 469-480) */
  int32_t device;       /* CUDA device ordinal */
  int32_t delay;        /* accepted local updates batched per G flush (0 = default 16) */
  int32_t reserved;
  double delta_tau, lambda, r, c, u;
} dqmc_params;

/* sparse factors taken from mc.l (lattice.jl:24-47), as Julia SparseMatrixCSC (Int64, 1-based) */
enum {
  DQMC_OP_HOP_HALF_B = 0,      /* l.chkr_hop_half[2]      */
  DQMC_OP_HOP_A = 1,           /* l.chkr_hop[1]           */
  DQMC_OP_HOP_HALF_INV_B = 2,  /* l.chkr_hop_half_inv[2]  */
  DQMC_OP_HOP_INV_A = 3,       /* l.chkr_hop_inv[1]       */
  DQMC_OP_MU = 4,              /* l.chkr_mu               */
  DQMC_OP_MU_INV = 5,          /* l.chkr_mu_inv           */
  DQMC_OP_HOP_HALF_A = 6,      /* l.chkr_hop_half[1]      (optional: only effective_greens2greens! needs it) */
  DQMC_OP_HOP_HALF_INV_A = 7,  /* l.chkr_hop_half_inv[1]  (optional, same) */
  DQMC_OP_COUNT = 8
};

/* dqmc_multiply_B op codes (slice_matrices.jl:101-226) */
enum { DQMC_B_LEFT = 0, DQMC_B_RIGHT = 1, DQMC_B_INV_LEFT = 2, DQMC_B_INV_RIGHT = 3, DQMC_B_DAGGER_LEFT = 4 };

/* replaces initialize_stack (stack.jl:224-242): allocates every device buffer of Stack{G} (stack.jl:45-118) */
int dqmc_create(dqmc_ctx** out, const dqmc_params* p);
int dqmc_destroy(dqmc_ctx* ctx);
const char* dqmc_last_error(dqmc_ctx* ctx);   /* ctx may be NULL: last global error */

/* data of init_checkerboard_matrices[_Bfield] (hoppings_checkerboard.jl:65-136,165-270); Julia SparseMatrixCSC arrays
 * (Int64, 1-based), copied.  Limit: every factor must decompose into disjoint groups of at most 4 coupled indices (a
 * plaquette of the Assaad checkerboard, or the 4 flavours of a site); the folded multi-bond groups of CBGeneric
 * (slice_matrices.jl:233-393) and the dense CBFalse factors are rejected with an error (out of BASELINE scope). */
int dqmc_set_operator(dqmc_ctx* ctx, int which, int64_t m, int64_t n, const int64_t* colptr,
                      const int64_t* rowval, const void* nzval, int nz_is_complex);
/* l.neighbors [4,N] (lattice.jl:112-117), 1-based; time neighbours are periodic (lattice.jl:144-153) */
int dqmc_set_neighbors(dqmc_ctx* ctx, const int64_t* neighbors);

/* mc.p.hsfield [opdim,N,M] Float64 (parameters.jl:19) */
int dqmc_set_hsfield(dqmc_ctx* ctx, const double* h);
int dqmc_get_hsfield(dqmc_ctx* ctx, double* h);
/* mc.s.greens [n,n] ComplexF64 (stack.jl:55).
 * Every Green's function of this model has the antiunitary flavour symmetry G = [[A, B], [-conj(B), conj(A)]] (flavour
 * blocks (1,2 | 3,4); it holds for the reference's own dumped G matrices, tests/test_oracle_golden.py), and the
 * local-update flush and the last product of calculate_greens use it (upper half computed, lower half mirrored) -- but
 * only while it is KNOWN to hold: the operators are checked for it in dqmc_set_operator, a G supplied here (or produced by
 * dqmc_calculate_greens_from) is measured, and if it deviates by more than 1e-10 max|G| the full-matrix paths are used
 * until the next internal calculate_greens.  Nothing is symmetrised silently. */
int dqmc_set_greens(dqmc_ctx* ctx, const double* g);
int dqmc_get_greens(dqmc_ctx* ctx, double* g);
/* mc.s.current_slice / mc.s.direction (stack.jl:391-499); slice in 0..M+1 */
int dqmc_get_state(dqmc_ctx* ctx, int32_t* slice, int32_t* direction);
int dqmc_set_state(dqmc_ctx* ctx, int32_t slice, int32_t direction);

/* build_stack (stack.jl:251-272) */
int dqmc_build_stack(dqmc_ctx* ctx);
/* propagate (stack.jl:391-499); returns the new (current_slice, direction) */
int dqmc_propagate(dqmc_ctx* ctx, int32_t* slice, int32_t* direction);
/* wrap_greens! (stack.jl:316-325) on mc.s.greens (g == NULL) or on a caller-owned host matrix.  direction = +1 uses
 * B(slice), -1 uses B(slice-1); that slice must lie in 1..M (the reference throws a BoundsError), otherwise -1 is returned.
 * Imaginary time is periodic (lattice.jl:144-153 time_neighbors): the library derives l+1 / l-1 itself and takes no
 * time_neighbors table. */
int dqmc_wrap_greens(dqmc_ctx* ctx, double* g_or_null, int32_t slice, int32_t direction);
/* multiply_B_left!/right!/inv_left!/inv_right!/daggered_B_left! (slice_matrices.jl:101-226) on a host n x n matrix */
int dqmc_multiply_B(dqmc_ctx* ctx, int op, int32_t slice, double* m);
/* calculate_greens (stack.jl:338-369) from caller-supplied UDTs (host) into mc.s.greens; also returns it if g != NULL.
 * Arbitrary inputs are allowed: this entry always forms the full matrix (no symmetry shortcut). */
int dqmc_calculate_greens_from(dqmc_ctx* ctx, const double* Ul, const double* Dl, const double* Tl,
                               const double* Ur, const double* Dr, const double* Tr, double* g_or_null);
/* calculate_logdet (stack.jl:377-385) of the last calculate_greens */
int dqmc_logdet(dqmc_ctx* ctx, double* logdet);
/* decompose_udt! (linalg.jl:20-39) of a host n x n matrix: U [n,n], D [n], T [n,n] */
int dqmc_decompose_udt(dqmc_ctx* ctx, const double* x, double* U, double* D, double* T);

/* local_updates (local_updates.jl:1-39) at mc.s.current_slice.  `u` is the caller's uniform [0,1) stream, consumed
 * in the reference's order (opdim proposal draws, then one accept draw only if p_acc <= 1); `consumed` tells the
 * caller how far to advance its own RNG.  dS_total accumulates the change of mc.p.boson_action. */
int dqmc_local_updates(dqmc_ctx* ctx, double box, const double* u, int64_t nu, int64_t* consumed,
                       int64_t* accepted, double* dS_total);
/* nupdates x { propagate; local_updates } = the body of `for u in 1:2M update(mc,i)` (dqmc_framework.jl:258-261,
 * 500-517) without global updates.  u == NULL uses the stream uploaded with dqmc_set_uniforms. */
int dqmc_sweep(dqmc_ctx* ctx, int32_t nupdates, double box, const double* u, int64_t nu, int64_t* consumed,
               int64_t* accepted, double* dS_total);
int dqmc_set_uniforms(dqmc_ctx* ctx, const double* u, int64_t nu);

/* calc_boson_action (action.jl:1-52) of the device-resident field */
int dqmc_calc_boson_action(dqmc_ctx* ctx, double* S);
/* measure_chi_dynamic (boson_measurements.jl:6-10,48-56) of the device-resident field: chi[qy,qx,w], column-major,
 * (L/2+1) x (L/2+1) x (M/2+1) Float64 */
int dqmc_measure_chi_dynamic(dqmc_ctx* ctx, double* chi);
/* global_update (global_updates.jl:18-59) at (current_slice, direction) == (slices, -1): shift the whole field by
 * randuniform(box_global) per component (u[0..2]), rebuild the stack, accept with exp(-dS) * exp(logdet_old - logdet_new)
 * (u[3], consumed only if p_acc <= 1); on rejection stack, G, logdet and field are restored from the backups. */
int dqmc_global_update(dqmc_ctx* ctx, double box_global, const double* u, double S_old, double* S_new, int32_t* accepted,
                       int32_t* consumed);

/* measure_tdgfs! (fermion_measurements.jl:1343-1407, with calc_Bchain_udts! :1434-1503, inv_sum_udts_scalettar! /
 * inv_one_plus_udt_scalettar! linalg.jl:302-331,512-567, effective_greens2greens! :1125-1142 and fill_tdgf! :1509-1541):
 * G(tau,0) and G(0,tau) for all M slices of the device-resident field, kept on the device (2 M n^2 ComplexF64 + four UDT
 * chains; allocated on first use, released by dqmc_free_tdgfs = deallocate_tdgfs_stacks!).  slices/2 must be a multiple of
 * safe_mult (fill_tdgf! propagates outwards from the stabilized slice at beta/2); other values return -1. */
int dqmc_measure_tdgfs(dqmc_ctx* ctx);
/* mc.s.meas.Gt0[slice] (which = 0) or G0t[slice] (which = 1), slice 1-based, into a host n x n matrix */
int dqmc_get_tdgf(dqmc_ctx* ctx, int which, int32_t slice, double* out);
int dqmc_free_tdgfs(dqmc_ctx* ctx);
/* inv_sum_udts_scalettar! (linalg.jl:512-567) of host operands: res = [Ua Da Ta + Ub Db Tb]^-1 */
int dqmc_inv_sum_udts(dqmc_ctx* ctx, const double* Ua, const double* Da, const double* Ta, const double* Ub, const double* Db,
                      const double* Tb, double* res);

/* telemetry: phase times in ms (CUDA events): [0] wrap, [1] local updates, [2] stack UDT (add_slice_sequence),
 * [3] calculate_greens, [4] total of dqmc_sweep; resets the accumulators.  Replaces the @mytimeit labels
 * (slice_matrices.jl:105-125, local_updates.jl:47,68, stack.jl:256,290,344). */
int dqmc_timers(dqmc_ctx* ctx, double* ms, int32_t n);
/* level 0: timers off (default); 1: only the total of dqmc_sweep ([4]) -- stabilization steps keep running as captured CUDA
 * graphs; 2: all phase timers (the graphs are bypassed so that events can be recorded between the kernels) */
int dqmc_set_timing(dqmc_ctx* ctx, int32_t level);
/* largest |G_wrapped - G_fresh| seen since the last call (the reference prints it when > 1e-7, stack.jl:426,477);
 * nonreal = number of proposals with |Im/Re| of the determinant ratio > 1e-4 (local_updates.jl:19-20) */
int dqmc_checks(dqmc_ctx* ctx, double* max_propagation_error, int64_t* nonreal);
int dqmc_sync(dqmc_ctx* ctx);

/* --- measurement hooks used by bench.py (not part of the reference's interface) --- */
/* time `reps` launches of one kernel group with CUDA events on the context's stream; returns ms per launch.
 * which: 0 wrap (+1, current slice), 1 ZGEMM n x n x n (hand-written DMMA), 2 cuBLAS ZGEMM n^3 (peak probe, dlopen),
 * 3 UDT (sort+QR+Q+T), 4 calculate_greens, 5 local_updates on slice 1 with the uploaded uniforms (state restored),
 * 6 device copy of G (HBM probe), 7 add_slice_sequence B-chain (safe_mult slices), 8-12 QR / trsm pieces,
 * 13 cuBLAS DGEMM 4096^3, 14 cuBLAS ZGEMM 4096^3 (FP64 ceilings on scratch buffers, dlopen, peak probes only) */
int dqmc_bench_kernel(dqmc_ctx* ctx, int which, int reps, double* ms_per_launch);
/* C = alpha*op(A)*op(B) + beta*C on host matrices through the hand-written kernel (test hook); op: 0 N, 1 T, 2 C */
int dqmc_test_zgemm(dqmc_ctx* ctx, int opA, int opB, int M, int N, int K, const double* alpha, const double* A, int lda,
                    const double* B, int ldb, const double* beta, double* C, int ldc);
/* test hooks of the half-matrix (antiunitary-symmetric) stabilization path: the paired Householder QR on host data (XL, rhs:
 * n x n/2 complex, pair-interleaved rows; V: n x n; Tfac: n/32 blocks of 32 x 32; dabs: n), and the sweep's decompose_udt! on a
 * host matrix (linalg.jl:20-39) through the sort-once QR (paired = 0) or the paired one (paired = 1, symmetric input) */
int dqmc_test_qr_paired(dqmc_ctx* ctx, double* XL, double* rhs, double* V, double* Tfac, double* dabs, int32_t lookahead);
int dqmc_test_udt(dqmc_ctx* ctx, const double* x, double* U, double* D, double* T, int32_t paired);
/* cycle counters of the last local_updates launch: [0] total, [1] stage 1, [2] stage 2, [3] flush, [4] #flushes,
 * [5] accepts, [8..15] per-warp stage-1 time, [16..23] per-warp role time before the speculative part,
 * [24..26] flush: first grid barrier, tiles, second grid barrier (debug; 32 values) */
int dqmc_lu_profile(dqmc_ctx* ctx, int32_t enable, int64_t* out32);
/* per-phase cycle counters of the QR panel kernels accumulated since the last call (debug; out16: 16 values; enable = 1 + the
 * cluster rank whose thread 0 is stamped in the paired kernel, 0 = off) */
int dqmc_qr_profile(dqmc_ctx* ctx, int32_t enable, int64_t* out16);
int64_t dqmc_kernel_launches(dqmc_ctx* ctx);   /* kernels launched by this context so far */

#ifdef __cplusplus
}
#endif
#endif
