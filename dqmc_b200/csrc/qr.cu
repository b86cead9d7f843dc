#include "qr.cuh"
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

// =====================================================================================================
// Panel factorization: one thread-block cluster of QR_CL CTAs; each CTA keeps a slab of rows of the
// m x nb panel in shared memory (row-major, 32 columns per row, padded to 33).  Thread (c, g) owns column c
// on the rows {g, g+8, ...} of the slab.  Per column ONE cluster barrier: every CTA reduces the dot products
// of the current column with all other panel columns over its rows (rows below the diagonal only) and pushes
// them, plus row j if it owns it, into every CTA's shared memory (distributed shared memory stores).  After the
// barrier every CTA derives beta, tau, v and the update coefficients redundantly from local data.  The dots with
// the already-finished columns give V^H v_j, i.e. the compact-WY T factor, for free.
// =====================================================================================================
#define QR_LDA 33
__global__ void __cluster_dims__(QR_CL, 1, 1) __launch_bounds__(256)
qr_panel_kernel(cplx* __restrict__ A, int lda, int m, int nb, cplx* __restrict__ tau_out,
                double* __restrict__ dabs_out, cplx* __restrict__ Tout, long long* __restrict__ prof) {
  cg::cluster_group cl = cg::this_cluster();
  long long pc[6] = {0, 0, 0, 0, 0, 0}, tprev = clock64();
#define PSTAMP(k) do { if (prof) { long long t_ = clock64(); pc[k] += t_ - tprev; tprev = t_; } } while (0)
  const int rank = (int)cl.block_rank();
  const int rs = (m + QR_CL - 1) / QR_CL;
  const int r_begin = rank * rs;
  const int nloc = max(0, min(m, r_begin + rs) - r_begin);
  const int tid = threadIdx.x, g = tid >> 5, c = tid & 31;

  extern __shared__ __align__(16) unsigned char smem_raw[];
  cplx* a = reinterpret_cast<cplx*>(smem_raw);   // a[rl*QR_LDA + c]
  __shared__ cplx part[8][QR_NB];                // per-warp partial dots
  __shared__ cplx xch[2][QR_CL][QR_NB];          // [parity][source CTA][column]: partial dots of every CTA
  __shared__ cplx rowv[2][QR_NB];                // [parity][column]: row j of the panel (pushed by its owner)
  __shared__ cplx Tsm[QR_NB][QR_NB + 1];
  __shared__ cplx gsm[QR_NB];
  __shared__ cplx tpart[8][QR_NB];
  __shared__ cplx tau_s[QR_NB];
  __shared__ double beta_s[QR_NB];

  // global -> shared: consecutive threads along rows (coalesced), 32 columns
  for (int e = tid; e < nloc * QR_NB; e += blockDim.x) {
    const int rl = e % nloc, cc = e / nloc;
    a[rl * QR_LDA + cc] = cc < nb ? A[(size_t)cc * lda + r_begin + rl] : cmake(0.0, 0.0);
  }
  if (rank == 0)
    for (int e = tid; e < QR_NB * (QR_NB + 1); e += blockDim.x) (&Tsm[0][0])[e] = cmake(0.0, 0.0);
  __syncthreads();

  for (int j = 0; j < nb; ++j) {
    const int par = j & 1;
    // ---- phase A: t_c = sum_{r>j} conj(a[r][j]) a[r][c] over my rows (4 independent accumulators)
    {
      cplx acc0 = cmake(0.0, 0.0), acc1 = acc0, acc2 = acc0, acc3 = acc0;
      int rl = g;
      // rows are visited in increasing order; skip those with r <= j
      const int first = j + 1 - r_begin;                    // first local row with r > j
      if (first > rl) rl += ((first - rl + 7) / 8) * 8;
      for (; rl + 24 < nloc; rl += 32) {
        cfma_conj(acc0, a[rl * QR_LDA + j], a[rl * QR_LDA + c]);
        cfma_conj(acc1, a[(rl + 8) * QR_LDA + j], a[(rl + 8) * QR_LDA + c]);
        cfma_conj(acc2, a[(rl + 16) * QR_LDA + j], a[(rl + 16) * QR_LDA + c]);
        cfma_conj(acc3, a[(rl + 24) * QR_LDA + j], a[(rl + 24) * QR_LDA + c]);
      }
      for (; rl < nloc; rl += 8) cfma_conj(acc0, a[rl * QR_LDA + j], a[rl * QR_LDA + c]);
      part[g][c] = cadd(cadd(acc0, acc1), cadd(acc2, acc3));
    }
    PSTAMP(0);
    if (rank == 0 && j > 0) {
      // T(0:jp,jp) = -tau_jp * T(0:jp,0:jp) * g for the previous column jp = j-1 (zlarft, forward/columnwise), computed
      // in the shadow of this column's dot products; thread (i = c, k = g mod 8)
      const int jp = j - 1;
      cplx acc = cmake(0.0, 0.0);
      if (c < jp)
        for (int k = c + ((g - c) & 7); k < jp; k += 8) cfma(acc, Tsm[c][k], gsm[k]);
      tpart[g][c] = acc;
    }
    __syncthreads();
    {
      // every warp g sums the 8 partials of column c and pushes the result to CTA g of the cluster
      cplx t = part[0][c];
#pragma unroll
      for (int w = 1; w < 8; ++w) t = cadd(t, part[w][c]);
      cplx* dst = cl.map_shared_rank(&xch[0][0][0], g);
      dst[(par * QR_CL + rank) * QR_NB + c] = t;
      const int owner = j / rs;
      if (rank == owner) {
        cplx* dr = cl.map_shared_rank(&rowv[0][0], g);
        dr[par * QR_NB + c] = a[(j - r_begin) * QR_LDA + c];
      }
      if (rank == 0 && j > 0 && g == 1) {
        const int jp = j - 1;
        if (c < jp) {
          cplx tt = tpart[0][c];
#pragma unroll
          for (int w = 1; w < 8; ++w) tt = cadd(tt, tpart[w][c]);
          Tsm[c][jp] = cneg(cmul(tau_s[jp], tt));
        } else if (c == jp) {
          Tsm[jp][jp] = tau_s[jp];
        }
      }
    }
    PSTAMP(1);
    cl.sync();
    // ---- phase C: totals and reflector parameters (every thread, redundantly, from local shared memory)
    cplx tc = xch[par][0][c], tj = xch[par][0][j];
    if (prof) { if (tc.x == 1.2345e300) pc[5]++; }
    PSTAMP(2);
#pragma unroll
    for (int src = 1; src < QR_CL; ++src) { tc = cadd(tc, xch[par][src][c]); tj = cadd(tj, xch[par][src][j]); }
    const cplx alpha = rowv[par][j], rowc = rowv[par][c];
    const double nrm = sqrt(alpha.x * alpha.x + alpha.y * alpha.y + tj.x);
    double beta;
    cplx tau, scale;
    if (nrm == 0.0) {
      beta = 0.0; tau = cmake(0.0, 0.0); scale = cmake(0.0, 0.0);
    } else {
      beta = alpha.x >= 0.0 ? -nrm : nrm;
      const double ib = 1.0 / beta;
      tau = cmake((beta - alpha.x) * ib, -alpha.y * ib);
      const double ar = alpha.x - beta, id = 1.0 / (ar * ar + alpha.y * alpha.y);
      scale = cmake(ar * id, -alpha.y * id);
    }
    // w_c = conj(tau) * v^H a_c = conj(tau) * (a[j][c] + conj(scale) * t_c)      (c > j)
    cplx wc = rowc;
    cfma(wc, cconj(scale), tc);
    wc = cmul(cconj(tau), wc);
    if (rank == 0 && g == 0) {
      // g_c = V_c^H v_j = conj(V[j][c]) + scale * conj(t_c)   (c < j)
      cplx gg = cconj(rowc);
      cfma(gg, scale, cconj(tc));
      gsm[c] = c < j ? gg : cmake(0.0, 0.0);
      if (c == 0) { tau_s[j] = tau; beta_s[j] = beta; }
    }
    PSTAMP(3);
    // ---- phase D: a[r][c] -= v_r w_c (r >= j, c > j); column j <- v (below the diagonal) and beta (diagonal)
    {
      int rl0 = g;
      const int first = j - r_begin;                        // first local row with r >= j
      if (first > rl0) rl0 += ((first - rl0 + 7) / 8) * 8;
      if (c > j) {
        int rl = rl0;
        for (; rl + 24 < nloc; rl += 32) {
          cplx x0 = a[rl * QR_LDA + j], x1 = a[(rl + 8) * QR_LDA + j], x2 = a[(rl + 16) * QR_LDA + j], x3 = a[(rl + 24) * QR_LDA + j];
          cplx t0 = a[rl * QR_LDA + c], t1 = a[(rl + 8) * QR_LDA + c], t2 = a[(rl + 16) * QR_LDA + c], t3 = a[(rl + 24) * QR_LDA + c];
          x0 = (r_begin + rl == j) ? cmake(1.0, 0.0) : cmul(x0, scale);
          x1 = cmul(x1, scale); x2 = cmul(x2, scale); x3 = cmul(x3, scale);
          a[rl * QR_LDA + c] = csub(t0, cmul(x0, wc));
          a[(rl + 8) * QR_LDA + c] = csub(t1, cmul(x1, wc));
          a[(rl + 16) * QR_LDA + c] = csub(t2, cmul(x2, wc));
          a[(rl + 24) * QR_LDA + c] = csub(t3, cmul(x3, wc));
        }
        for (; rl < nloc; rl += 8) {
          cplx x0 = a[rl * QR_LDA + j];
          x0 = (r_begin + rl == j) ? cmake(1.0, 0.0) : cmul(x0, scale);
          a[rl * QR_LDA + c] = csub(a[rl * QR_LDA + c], cmul(x0, wc));
        }
      }
      __syncwarp();
      // column j itself: lane l of warp g rewrites row rl0 + 8*l (every row of the slab belongs to exactly one warp)
      for (int rl = rl0 + 8 * c; rl < nloc; rl += 8 * 32) {
        const int r = r_begin + rl;
        a[rl * QR_LDA + j] = (r == j) ? cmake(beta, 0.0) : cmul(a[rl * QR_LDA + j], scale);
      }
    }
    PSTAMP(4);
    __syncthreads();
  }
  if (prof && rank == 0 && tid == 0) { for (int q = 0; q < 5; ++q) prof[q] += pc[q]; prof[5] += nb; }
  if (rank == 0) {   // T column of the last reflector
    const int jp = nb - 1;
    cplx acc = cmake(0.0, 0.0);
    if (c < jp)
      for (int k = c + ((g - c) & 7); k < jp; k += 8) cfma(acc, Tsm[c][k], gsm[k]);
    tpart[g][c] = acc;
    __syncthreads();
    if (g == 0) {
      if (c < jp) {
        cplx tt = tpart[0][c];
#pragma unroll
        for (int w = 1; w < 8; ++w) tt = cadd(tt, tpart[w][c]);
        Tsm[c][jp] = cneg(cmul(tau_s[jp], tt));
      } else if (c == jp) {
        Tsm[jp][jp] = tau_s[jp];
      }
    }
    __syncthreads();
  }

  for (int e = tid; e < nloc * nb; e += blockDim.x) {
    const int rl = e % nloc, cc = e / nloc;
    A[(size_t)cc * lda + r_begin + rl] = a[rl * QR_LDA + cc];
  }
  if (rank == 0) {
    for (int e = tid; e < QR_NB * QR_NB; e += blockDim.x) {
      int i = e % QR_NB, k = e / QR_NB;
      Tout[e] = (i < nb && k < nb) ? Tsm[i][k] : cmake(0.0, 0.0);
    }
    if (tid < nb) { tau_out[tid] = tau_s[tid]; dabs_out[tid] = fabs(beta_s[tid]); }
  }
  cl.sync();  // no CTA may exit while others may still write into its shared memory
}

// =====================================================================================================
// Block-reflector application  C <- (I - V op(T) V^H) C  on 8-column blocks of C, streaming V and C
// from L2 in DMMA fragment order (no shared-memory staging of the operands).
// =====================================================================================================
__device__ __forceinline__ cplx vmask_load(const cplx* __restrict__ V, int ldv, int m, int r, int c) {
  if (r >= m || r < c) return cmake(0.0, 0.0);
  if (r == c) return cmake(1.0, 0.0);
  return V[(size_t)c * ldv + r];
}

__global__ void __launch_bounds__(256)
larfb_kernel(const cplx* __restrict__ V, int ldv, int m, const cplx* __restrict__ T, int conjT,
             cplx* __restrict__ C, int ldc, int ncols) {
  __shared__ cplx Tsm[QR_NB][QR_NB + 1];
  __shared__ double Wp[4][4][32][4];
  __shared__ cplx W1[QR_NB][8], W2[QR_NB][8];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int lo = lane >> 2, lk = lane & 3;
  const int c0 = blockIdx.x * 8;

  for (int e = tid; e < QR_NB * QR_NB; e += blockDim.x) Tsm[e % QR_NB][e / QR_NB] = T[e];

  // ---- phase 1: W = V^H C   (32 x 8), K = m split over the 8 warps
  double cr[4][2], ci[4][2];
#pragma unroll
  for (int it = 0; it < 4; ++it) cr[it][0] = cr[it][1] = ci[it][0] = ci[it][1] = 0.0;
  const int kchunk = (((m + 7) / 8) + 3) / 4 * 4;
  const int kbeg = warp * kchunk, kend = min(m, kbeg + kchunk);
  const bool colok = (c0 + lo) < ncols;
  for (int r0 = kbeg; r0 < kend; r0 += 8) {
    cplx av[2][4], bv[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int r = r0 + 4 * u + lk;
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int c = it * 8 + lo;
        av[u][it] = (r >= QR_NB && r < kend) ? V[(size_t)c * ldv + r] : (r < kend ? vmask_load(V, ldv, m, r, c) : cmake(0.0, 0.0));
      }
      bv[u] = (r < kend && colok) ? C[(size_t)(c0 + lo) * ldc + r] : cmake(0.0, 0.0);
    }
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
      for (int it = 0; it < 4; ++it) {   // A operand = conj(V)
        dmma884(cr[it][0], cr[it][1], av[u][it].x, bv[u].x);
        dmma884(cr[it][0], cr[it][1], av[u][it].y, bv[u].y);
        dmma884(ci[it][0], ci[it][1], av[u][it].x, bv[u].y);
        dmma884(ci[it][0], ci[it][1], -av[u][it].y, bv[u].x);
      }
  }
  if (warp >= 4) {
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      Wp[warp - 4][it][lane][0] = cr[it][0]; Wp[warp - 4][it][lane][1] = cr[it][1];
      Wp[warp - 4][it][lane][2] = ci[it][0]; Wp[warp - 4][it][lane][3] = ci[it][1];
    }
  }
  __syncthreads();
  if (warp < 4) {
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      Wp[warp][it][lane][0] += cr[it][0]; Wp[warp][it][lane][1] += cr[it][1];
      Wp[warp][it][lane][2] += ci[it][0]; Wp[warp][it][lane][3] += ci[it][1];
    }
  }
  __syncthreads();
  {
    const int i = tid >> 3, c = tid & 7;            // output W[i][c]
    const int it = i >> 3, ln = (i & 7) * 4 + (c >> 1), e = c & 1;
    double sr = 0.0, si = 0.0;
#pragma unroll
    for (int w = 0; w < 4; ++w) { sr += Wp[w][it][ln][e]; si += Wp[w][it][ln][2 + e]; }
    W1[i][c] = cmake(sr, si);
  }
  __syncthreads();
  // ---- phase 2: W <- op(T) W
  {
    const int i = tid >> 3, c = tid & 7;
    cplx acc = cmake(0.0, 0.0);
    if (conjT) { for (int k = 0; k < QR_NB; ++k) cfma_conj(acc, Tsm[k][i], W1[k][c]); }
    else       { for (int k = 0; k < QR_NB; ++k) cfma(acc, Tsm[i][k], W1[k][c]); }
    W2[i][c] = acc;
  }
  __syncthreads();
  // ---- phase 3: C -= V W   (m x 8), K = 32
  const int ntile = (m + 7) / 8;
  for (int rt = warp; rt < ntile; rt += 8) {
    const int r = rt * 8 + lo;
    cplx av[8];
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) {
      const int c = kk * 4 + lk;
      av[kk] = (rt * 8 >= QR_NB && r < m) ? V[(size_t)c * ldv + r] : vmask_load(V, ldv, m, r, c);
    }
    double dr[2] = {0.0, 0.0}, di[2] = {0.0, 0.0};
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) {
      const cplx b = W2[kk * 4 + lk][lo];
      dmma884(dr[0], dr[1], av[kk].x, b.x);
      dmma884(dr[0], dr[1], -av[kk].y, b.y);
      dmma884(di[0], di[1], av[kk].x, b.y);
      dmma884(di[0], di[1], av[kk].y, b.x);
    }
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int col = c0 + 2 * lk + e;
      if (r < m && col < ncols) {
        cplx* p = C + (size_t)col * ldc + r;
        cplx t = *p;
        t.x -= dr[e]; t.y -= di[e];
        *p = t;
      }
    }
  }
}

__global__ void set_identity_kernel(cplx* Q, int ldq, int n) {
  size_t tot = (size_t)n * n;
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < tot; e += (size_t)gridDim.x * blockDim.x) {
    int r = (int)(e % n), c = (int)(e / n);
    Q[(size_t)c * ldq + r] = cmake(r == c ? 1.0 : 0.0, 0.0);
  }
}

long long* g_qr_prof = nullptr;
static int launch_panel(cudaStream_t st, cplx* A, int lda, int m, int nb, cplx* tau, double* dabs, cplx* T) {
  const int rs = (m + QR_CL - 1) / QR_CL;
  const size_t smem = sizeof(cplx) * ((size_t)QR_LDA * rs + 8);
  static size_t smem_lim = 0;
  if (smem_lim == 0 && set_max_dynamic_smem(qr_panel_kernel, &smem_lim)) return -1;
  if (smem > smem_lim) { snprintf(g_errbuf, sizeof(g_errbuf), "qr panel: m=%d too large", m); return -1; }
  qr_panel_kernel<<<QR_CL, 256, smem, st>>>(A, lda, m, nb, tau, dabs, T, g_qr_prof);
  CUDA_TRY(cudaGetLastError());
  g_launches++;
  return 0;
}

static int launch_larfb(cudaStream_t st, const cplx* V, int ldv, int m, const cplx* T, int conjT, cplx* C, int ldc,
                        int ncols) {
  if (ncols <= 0) return 0;
  larfb_kernel<<<(ncols + 7) / 8, 256, 0, st>>>(V, ldv, m, T, conjT, C, ldc, ncols);
  CUDA_TRY(cudaGetLastError());
  g_launches++;
  return 0;
}

// Right-looking blocked QR with one-panel lookahead: while the cluster kernel factors panel k+1 on `st`, the block
// reflector of panel k is applied to the rest of the trailing matrix (and to the right-hand side) on `st2`.
int qr_factor(cudaStream_t st, cplx* A, int lda, int n, cplx* tau, double* dabs, cplx* tfac, cplx* rhs, int ldr,
              int nrhs, int num_sms, const QrAsync* as) {
  (void)num_sms;
  const bool la = as != nullptr && n > 2 * QR_NB;
  for (int j0 = 0, k = 0; j0 < n; j0 += QR_NB, ++k) {
    const int nb = min(QR_NB, n - j0), m = n - j0;
    cplx* P = A + (size_t)j0 * lda + j0;
    cplx* T = tfac + (size_t)k * QR_NB * QR_NB;
    if (launch_panel(st, P, lda, m, nb, tau + j0, dabs + j0, T)) return -1;
    const int ntrail = n - j0 - nb;
    if (!la) {
      if (launch_larfb(st, P, lda, m, T, 1, P + (size_t)nb * lda, lda, ntrail)) return -1;
      if (rhs && launch_larfb(st, P, lda, m, T, 1, rhs + j0, ldr, nrhs)) return -1;
      continue;
    }
    // columns of the next panel first (they must have received every earlier update: wait for st2's previous step)
    const int nnext = min(QR_NB, ntrail);
    if (k > 0) CUDA_TRY(cudaStreamWaitEvent(st, as->eB, 0));
    if (launch_larfb(st, P, lda, m, T, 1, P + (size_t)nb * lda, lda, nnext)) return -1;
    CUDA_TRY(cudaEventRecord(as->eA, st));
    CUDA_TRY(cudaStreamWaitEvent(as->st2, as->eA, 0));
    if (launch_larfb(as->st2, P, lda, m, T, 1, P + (size_t)(nb + nnext) * lda, lda, ntrail - nnext)) return -1;
    if (rhs && launch_larfb(as->st2, P, lda, m, T, 1, rhs + j0, ldr, nrhs)) return -1;
    CUDA_TRY(cudaEventRecord(as->eB, as->st2));
  }
  if (la) CUDA_TRY(cudaStreamWaitEvent(st, as->eB, 0));
  return 0;
}

int qr_form_q(cudaStream_t st, const cplx* A, int lda, int n, const cplx* tfac, cplx* Q, int ldq, int num_sms) {
  set_identity_kernel<<<num_sms * 4, 256, 0, st>>>(Q, ldq, n);
  CUDA_TRY(cudaGetLastError());
  g_launches++;
  const int nblk = (n + QR_NB - 1) / QR_NB;
  for (int k = nblk - 1; k >= 0; --k) {
    const int j0 = k * QR_NB, m = n - j0;
    if (launch_larfb(st, A + (size_t)j0 * lda + j0, lda, m, tfac + (size_t)k * QR_NB * QR_NB, 0,
                     Q + (size_t)j0 * ldq + j0, ldq, n - j0))
      return -1;
  }
  return 0;
}

// =====================================================================================================
// Triangular solve with many right-hand sides: every CTA owns 8 columns of Y in shared memory and runs the
// whole blocked back substitution for them (inverted 32x32 diagonal blocks, DMMA updates streaming R from L2).
// =====================================================================================================
__global__ void __launch_bounds__(32) trtri_diag_kernel(const cplx* __restrict__ A, int lda, int n, cplx* __restrict__ inv) {
  __shared__ cplx R[QR_NB][QR_NB + 1];
  const int kb = blockIdx.x, j0 = kb * QR_NB, nb = min(QR_NB, n - j0), j = threadIdx.x;
  for (int c = 0; c < QR_NB; ++c) {
    cplx v = cmake(0.0, 0.0);
    if (j < nb && c < nb && j <= c) v = A[(size_t)(j0 + c) * lda + j0 + j];
    if (j == c && j >= nb) v = cmake(1.0, 0.0);
    R[j][c] = v;
  }
  __syncwarp();
  cplx x[QR_NB];
#pragma unroll
  for (int i = 0; i < QR_NB; ++i) x[i] = cmake(0.0, 0.0);
  // column j of the inverse: back substitution of R x = e_j (fully unrolled so x stays in registers)
#pragma unroll
  for (int i = QR_NB - 1; i >= 0; --i) {
    if (i <= j) {
      cplx acc = cmake(i == j ? 1.0 : 0.0, 0.0);
#pragma unroll
      for (int k = i + 1; k < QR_NB; ++k)
        if (k <= j) { cplx t = cmul(R[i][k], x[k]); acc = csub(acc, t); }
      x[i] = cdiv(acc, R[i][i]);
    }
  }
#pragma unroll
  for (int i = 0; i < QR_NB; ++i) inv[(size_t)kb * QR_NB * QR_NB + (size_t)j * QR_NB + i] = x[i];
}

__global__ void __launch_bounds__(256)
trsm_kernel(const cplx* __restrict__ A, int lda, int n, cplx* __restrict__ Y, int ldy, int nrhs,
            const cplx* __restrict__ inv, const double* __restrict__ rowscale) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int n8 = (n + 7) / 8 * 8, lds = n8 + 4;
  cplx* Ys = reinterpret_cast<cplx*>(smem_raw);        // [8][lds]
  cplx* Xs = Ys + (size_t)8 * lds;                      // [32][8]
  cplx* Is = Xs + QR_NB * 8;                            // [32*32] inverse of the current diagonal block
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int lo = lane >> 2, lk = lane & 3;
  const int c0 = blockIdx.x * 8;
  for (int c = 0; c < 8; ++c)
    for (int r = tid; r < n8; r += blockDim.x)
      Ys[(size_t)c * lds + r] = (r < n && c0 + c < nrhs) ? Y[(size_t)(c0 + c) * ldy + r] : cmake(0.0, 0.0);
  const int nblk = (n + QR_NB - 1) / QR_NB;
  for (int e = tid; e < QR_NB * QR_NB; e += blockDim.x) Is[e] = inv[(size_t)(nblk - 1) * QR_NB * QR_NB + e];
  __syncthreads();
  for (int kb = nblk - 1; kb >= 0; --kb) {
    const int j0 = kb * QR_NB;
    const cplx* invn = inv + (size_t)(kb > 0 ? kb - 1 : 0) * QR_NB * QR_NB;
    const cplx pre0 = invn[tid], pre1 = invn[tid + 256], pre2 = invn[tid + 512], pre3 = invn[tid + 768];
    {   // x = inv(R_kk) y_k
      const int i = tid & 31, c = tid >> 5;
      cplx acc = cmake(0.0, 0.0);
#pragma unroll 8
      for (int k = 0; k < QR_NB; ++k) {
        const int r = j0 + k;
        if (r < n8) cfma(acc, Is[k * QR_NB + i], Ys[(size_t)c * lds + r]);
      }
      Xs[i * 8 + c] = acc;
    }
    __syncthreads();
    {
      const int i = tid & 31, c = tid >> 5;
      if (j0 + i < n8) Ys[(size_t)c * lds + j0 + i] = Xs[i * 8 + c];
    }
    // y[0:j0] -= R[0:j0, block kb] x
    const int ntile = j0 / 8;
    for (int rt = warp; rt < ntile; rt += 8) {
      const int r = rt * 8 + lo;
      cplx av[8];
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        const int c = j0 + kk * 4 + lk;
        av[kk] = c < n ? A[(size_t)c * lda + r] : cmake(0.0, 0.0);
      }
      double dr[2] = {0.0, 0.0}, di[2] = {0.0, 0.0};
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        const cplx b = Xs[(kk * 4 + lk) * 8 + lo];
        dmma884(dr[0], dr[1], av[kk].x, b.x);
        dmma884(dr[0], dr[1], -av[kk].y, b.y);
        dmma884(di[0], di[1], av[kk].x, b.y);
        dmma884(di[0], di[1], av[kk].y, b.x);
      }
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        cplx* p = Ys + (size_t)(2 * lk + e) * lds + r;
        cplx t = *p;
        t.x -= dr[e]; t.y -= di[e];
        *p = t;
      }
    }
    __syncthreads();
    Is[tid] = pre0; Is[tid + 256] = pre1; Is[tid + 512] = pre2; Is[tid + 768] = pre3;
    __syncthreads();
  }
  for (int c = 0; c < 8; ++c)
    if (c0 + c < nrhs)
      for (int r = tid; r < n; r += blockDim.x) {
        cplx t = Ys[(size_t)c * lds + r];
        if (rowscale) t = cscale(t, rowscale[r]);
        Y[(size_t)(c0 + c) * ldy + r] = t;
      }
}

int trsm_upper(cudaStream_t st, const cplx* A, int lda, int n, cplx* Y, int ldy, int nrhs, cplx* work,
               const double* rowscale, int num_sms) {
  (void)num_sms;
  const int nblk = (n + QR_NB - 1) / QR_NB;
  trtri_diag_kernel<<<nblk, 32, 0, st>>>(A, lda, n, work);
  CUDA_TRY(cudaGetLastError());
  g_launches++;
  const int n8 = (n + 7) / 8 * 8, lds = n8 + 4;
  const size_t smem = sizeof(cplx) * ((size_t)8 * lds + QR_NB * 8 + QR_NB * QR_NB);
  static size_t smem_lim = 0;
  if (smem_lim == 0 && set_max_dynamic_smem(trsm_kernel, &smem_lim)) return -1;
  if (smem > smem_lim) { snprintf(g_errbuf, sizeof(g_errbuf), "trsm: n=%d too large", n); return -1; }
  trsm_kernel<<<(nrhs + 7) / 8, 256, smem, st>>>(A, lda, n, Y, ldy, nrhs, work, rowscale);
  CUDA_TRY(cudaGetLastError());
  g_launches++;
  return 0;
}
