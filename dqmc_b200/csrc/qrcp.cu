// Column-pivoted Householder QR (the reference's decompose_udt! = LAPACK zgeqp3, linalg.jl:20-39) as ONE persistent
// cooperative kernel.  Used where the conditioning of T = D^-1 R P^T matters: the UDT chains of the time-displaced Green's
// functions, whose inverse sums apply T^-1 (linalg.jl:512-567) -- with the sort-once-then-unpivoted QR of the sweep path
// cond(T) reaches 1e6..1e7 at beta = 40 and costs four digits there; with true pivoting it stays ~1e2.
//
// Level-2 by nature (every step touches the whole trailing matrix), so the design goal is one grid barrier per column:
//   * physical columns never move: `pos[c]` = position of column c in the pivot order (or -1 while it is active);
//   * columns are dealt cyclically to the CTAs; the E = identity block (n more columns, never pivot candidates) rides
//     along and ends up as Q^H, so no separate Q formation is needed;
//   * per step j every CTA redundantly (a) finds the active column of largest residual norm, (b) reads it from L2 and forms
//     the reflector (zlarfg), then (c) applies H_j^H to its own columns, one warp per column staged through shared memory,
//     and recomputes their residual norms exactly (no downdating, hence no cancellation safeguard);
//   * grid barrier; next column.
#include <cooperative_groups.h>

#include "misc.cuh"
#include "qr.cuh"

#define QRCP_WARPS 8

__device__ __forceinline__ cplx ldcg_c(const cplx* p) { return __ldcg(reinterpret_cast<const double2*>(p)); }

__global__ void __launch_bounds__(QRCP_WARPS * 32)
qrcp_kernel(cplx* A, int lda, cplx* E, int lde, int n, double* vn, int* __restrict__ perm,
            int* __restrict__ pos_out, double* __restrict__ dabs, double* __restrict__ rdiag, unsigned int* __restrict__ bar, int staged) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cplx* vbuf = reinterpret_cast<cplx*>(smem_raw);               // [n] current reflector (v[0] = 1)
  int* pos = reinterpret_cast<int*>(vbuf + n);                  // [n] position in the pivot order, -1 = active
  // [QRCP_WARPS][n] column staging, one slot per warp (large n: not staged, the second pass re-reads the column through L1)
  cplx* stage = reinterpret_cast<cplx*>(smem_raw + (((size_t)n * (sizeof(cplx) + sizeof(int))) + 15) / 16 * 16);
  __shared__ double red_v[QRCP_WARPS];
  __shared__ int red_i[QRCP_WARPS];
  __shared__ double s_xn2;
  __shared__ int s_pc;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int G = gridDim.x, b = blockIdx.x;
  unsigned int bar_target = 0;

  // ---- initial norms of my columns
  for (int c = b + G * warp; c < n; c += G * QRCP_WARPS) {
    double s = 0.0;
    for (int r = lane; r < n; r += 32) s += cabs2(A[(size_t)c * lda + r]);
    s = warp_sum(s);
    if (lane == 0) vn[c] = sqrt(s);
  }
  for (int c = tid; c < n; c += blockDim.x) pos[c] = -1;
  bar_target += G; grid_barrier_mono(bar, bar_target);

  for (int j = 0; j < n; ++j) {
    const int m = n - j;                                        // rows j..n-1 take part
    // ---- (a) pivot: active column with the largest residual norm (lowest index on ties)
    {
      double best = -1.0; int bi = n;
      for (int c = tid; c < n; c += blockDim.x) {
        if (pos[c] < 0) {
          double v = __ldcg(vn + c);
          if (!(v >= 0.0)) v = 0.0;                             // a NaN norm must not leave the step without a pivot
          if (v > best || (v == best && c < bi)) { best = v; bi = c; }
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
      }
      if (lane == 0) { red_v[warp] = best; red_i[warp] = bi; }
      __syncthreads();
      if (tid == 0) {
        double bv = red_v[0]; int bc = red_i[0];
        for (int w = 1; w < QRCP_WARPS; ++w)
          if (red_v[w] > bv || (red_v[w] == bv && red_i[w] < bc)) { bv = red_v[w]; bc = red_i[w]; }
        s_pc = bc;
      }
      __syncthreads();
    }
    const int pc = s_pc;
    // ---- (b) reflector of column pc on rows j.. (zlarfg): beta = -sign(Re alpha) sqrt(|alpha|^2 + |x|^2), tau = (beta - alpha) / beta,
    //      v = x / (alpha - beta)
    {
      const cplx* col = A + (size_t)pc * lda + j;
      double s = 0.0;
      for (int r = tid; r < m; r += blockDim.x) {
        const cplx x = ldcg_c(col + r);
        vbuf[r] = x;
        if (r > 0) s += cabs2(x);
      }
      s = warp_sum(s);
      if (lane == 0) red_v[warp] = s;
      __syncthreads();
      if (tid == 0) {
        double t = 0.0;
        for (int w = 0; w < QRCP_WARPS; ++w) t += red_v[w];
        s_xn2 = t;
      }
      __syncthreads();
    }
    const cplx alpha = vbuf[0];
    const double xn2 = s_xn2;
    cplx tau = cmake(0.0, 0.0);
    double beta = alpha.x;
    const bool trivial = (xn2 == 0.0 && alpha.y == 0.0);
    cplx inv = cmake(0.0, 0.0);
    if (!trivial) {
      beta = -copysign(sqrt(alpha.x * alpha.x + alpha.y * alpha.y + xn2), alpha.x);
      tau = cmake((beta - alpha.x) / beta, -alpha.y / beta);
      inv = cdiv(cmake(1.0, 0.0), cmake(alpha.x - beta, alpha.y));
    }
    __syncthreads();                                            // everybody has read vbuf[0] before it becomes 1
    for (int r = tid; r < m; r += blockDim.x) vbuf[r] = (r == 0) ? cmake(1.0, 0.0) : (trivial ? cmake(0.0, 0.0) : cmul(vbuf[r], inv));
    if (tid == 0) pos[pc] = j;
    __syncthreads();
    if (b == pc % G && tid == 0) {                              // the owner records the step
      // R_jj = beta goes to a side array, NOT into A[j, pc]: slower CTAs may still be reading the pivot column in this step
      rdiag[j] = beta;
      dabs[j] = fabs(beta);
      perm[j] = pc;
      pos_out[pc] = j;
    }
    // ---- (c) H_j^H = I - conj(tau) v v^H on my columns: active columns of A, and every column of E
    const cplx ctau = cmake(tau.x, -tau.y);
    cplx* st = stage + (size_t)warp * n;
    for (int cc = b + G * warp; cc < 2 * n; cc += G * QRCP_WARPS) {
      const bool isA = cc < n;
      if (isA && pos[cc] >= 0) continue;                        // finished columns (incl. pc) are left alone
      cplx* col = isA ? A + (size_t)cc * lda + j : E + (size_t)(cc - n) * lde + j;
      cplx w = cmake(0.0, 0.0);
#pragma unroll 4
      for (int r = lane; r < m; r += 32) {
        const cplx x = col[r];
        if (staged) st[r] = x;
        cfma_conj(w, vbuf[r], x);
      }
      w.x = warp_sum(w.x); w.y = warp_sum(w.y);
      const cplx f = cmul(ctau, w);
      double s = 0.0;
#pragma unroll 4
      for (int r = lane; r < m; r += 32) {
        cplx x = staged ? st[r] : col[r];
        const cplx d = cmul(f, vbuf[r]);
        x.x -= d.x; x.y -= d.y;
        col[r] = x;
        if (r > 0) s += cabs2(x);
      }
      if (isA) {
        s = warp_sum(s);
        if (lane == 0) vn[cc] = sqrt(s);
      }
      __syncwarp();
    }
    bar_target += G; grid_barrier_mono(bar, bar_target);
  }
}

// T[i, c] = R[i, c] / dabs[i] for i <= pos[c], else 0 (columns stayed in place, so T = D^-1 R P^T needs no scatter)
__global__ void qrcp_build_T_kernel(const cplx* __restrict__ A, int lda, int n, const double* __restrict__ dabs, const double* __restrict__ rdiag,
                                    const int* __restrict__ pos, cplx* __restrict__ T, int ldt) {
  for (int c = blockIdx.x; c < n; c += gridDim.x) {
    const int p = pos[c];
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      cplx v = cmake(0.0, 0.0);
      if (i < p) v = cscale(A[(size_t)c * lda + i], 1.0 / dabs[i]);
      else if (i == p) v = cmake(rdiag[p] / dabs[p], 0.0);
      T[(size_t)c * ldt + i] = v;
    }
  }
}

static size_t qrcp_smem_base(int n) { return (((size_t)n * (sizeof(cplx) + sizeof(int))) + 15) / 16 * 16; }
size_t qrcp_smem(int n, bool staged) { return qrcp_smem_base(n) + (staged ? sizeof(cplx) * (size_t)QRCP_WARPS * n : 0); }

// A (n x n, destroyed: R in the rows i <= pos[c] of every column c) -> QH = Q^H (n x n), dabs = |R_jj|, pos / perm, T = D^-1 R P^T.
// vn, rdiag: n doubles of scratch each, bar: one unsigned int (reset here).
int qrcp_udt(cudaStream_t st, cplx* A, int lda, int n, cplx* QH, int ldq, cplx* T, int ldt, double* dabs, double* vn, double* rdiag,
             int* perm, int* pos, unsigned int* bar, int num_sms) {
  static SmemMemo memo;
  size_t lim = 0;
  if (ensure_max_dynamic_smem(qrcp_kernel, memo, &lim)) return -1;
  int staged = qrcp_smem(n, true) <= lim ? 1 : 0;
  const size_t smem = qrcp_smem(n, staged != 0);
  if (smem > lim) { snprintf(g_errbuf, sizeof(g_errbuf), "qrcp: n = %d needs %zu bytes of shared memory (> %zu)", n, smem, lim); return -1; }
  if (set_identity(st, QH, ldq, n, num_sms)) return -1;
  CUDA_TRY(cudaMemsetAsync(bar, 0, sizeof(unsigned int), st));
  int grid = num_sms;
  if (grid > 2 * n) grid = 2 * n;
  void* params[] = {&A, &lda, &QH, &ldq, &n, &vn, &perm, &pos, &dabs, &rdiag, &bar, &staged};
  CUDA_TRY(cudaLaunchCooperativeKernel((const void*)qrcp_kernel, dim3(grid), dim3(QRCP_WARPS * 32), params, smem, st));
  g_launches++;
  qrcp_build_T_kernel<<<min(n, num_sms * 8), 256, 0, st>>>(A, lda, n, dabs, rdiag, pos, T, ldt);
  CUDA_TRY(cudaGetLastError());
  g_launches++;
  return 0;
}
