#include "qr.cuh"
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

// =====================================================================================================
// Panel factorization: one thread-block cluster of QR_CL CTAs; each CTA keeps a slab of rows of the
// m x nb panel in shared memory.  Per column ONE cluster barrier: every CTA publishes the partial dot
// products of the current column with all other panel columns (rows below the diagonal only) in its own
// shared memory, the others pull them through distributed shared memory.  From those sums every CTA
// derives beta, tau, v and the update coefficients redundantly, so no second exchange is needed; the
// dots with the already-finished columns give V^H v_j, i.e. the compact-WY T factor for free.
// =====================================================================================================
__global__ void __cluster_dims__(QR_CL, 1, 1) __launch_bounds__(256)
qr_panel_kernel(cplx* __restrict__ A, int lda, int m, int nb, cplx* __restrict__ tau_out,
                double* __restrict__ dabs_out, cplx* __restrict__ Tout) {
  cg::cluster_group cl = cg::this_cluster();
  const int rank = (int)cl.block_rank();
  const int rs = (m + QR_CL - 1) / QR_CL;
  const int r_begin = rank * rs;
  const int nloc = max(0, min(m, r_begin + rs) - r_begin);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  extern __shared__ __align__(16) unsigned char smem_raw[];
  cplx* a = reinterpret_cast<cplx*>(smem_raw);   // a[c*rs + rloc]
  cplx* vloc = a + (size_t)QR_NB * rs;           // [rs]
  __shared__ cplx xch[2][QR_NB];                 // my partial dots, double-buffered by column parity
  __shared__ cplx rowv[2][QR_NB];                // row j of the panel (owner CTA only)
  __shared__ cplx Ptot[QR_NB], rowj[QR_NB], wv[QR_NB];
  __shared__ cplx Tsm[QR_NB][QR_NB + 1];
  __shared__ cplx gsm[QR_NB];
  __shared__ cplx tau_s[QR_NB];
  __shared__ double beta_s[QR_NB];

  for (int c = 0; c < nb; ++c)
    for (int r = tid; r < nloc; r += blockDim.x) a[c * rs + r] = A[(size_t)c * lda + r_begin + r];
  if (rank == 0)
    for (int e = tid; e < QR_NB * (QR_NB + 1); e += blockDim.x) (&Tsm[0][0])[e] = cmake(0.0, 0.0);
  __syncthreads();

  for (int j = 0; j < nb; ++j) {
    const int par = j & 1;
    // ---- phase A: partial dots over my rows r > j
    for (int cc = warp; cc < nb; cc += 8) {
      cplx acc = cmake(0.0, 0.0);
      for (int rl = lane; rl < nloc; rl += 32) {
        if (r_begin + rl > j) {
          if (cc >= j) cfma_conj(acc, a[j * rs + rl], a[cc * rs + rl]);
          else cfma_conj(acc, a[cc * rs + rl], a[j * rs + rl]);
        }
      }
      acc.x = warp_sum(acc.x);
      acc.y = warp_sum(acc.y);
      if (lane == 0) xch[par][cc] = acc;
    }
    const int owner = j / rs;
    if (rank == owner && tid < nb) rowv[par][tid] = a[tid * rs + (j - r_begin)];
    cl.sync();
    // ---- phase B: pull the partials of all CTAs
    if (tid < nb) {
      cplx tot = cmake(0.0, 0.0);
#pragma unroll
      for (int src = 0; src < QR_CL; ++src) {
        const cplx* rx = cl.map_shared_rank(&xch[0][0], src);
        tot = cadd(tot, rx[par * QR_NB + tid]);
      }
      Ptot[tid] = tot;
      const cplx* rr = cl.map_shared_rank(&rowv[0][0], owner);
      rowj[tid] = rr[par * QR_NB + tid];
    }
    __syncthreads();
    // ---- phase C: reflector parameters (every thread, redundantly)
    const cplx alpha = rowj[j];
    const double nrm = sqrt(alpha.x * alpha.x + alpha.y * alpha.y + Ptot[j].x);
    double beta;
    cplx tau, scale;
    if (nrm == 0.0) {
      beta = 0.0; tau = cmake(0.0, 0.0); scale = cmake(0.0, 0.0);
    } else {
      beta = alpha.x >= 0.0 ? -nrm : nrm;
      tau = cmake((beta - alpha.x) / beta, -alpha.y / beta);
      scale = cdiv(cmake(1.0, 0.0), cmake(alpha.x - beta, alpha.y));
    }
    const cplx ctau = cconj(tau);
    for (int rl = tid; rl < nloc; rl += blockDim.x) {
      int r = r_begin + rl;
      vloc[rl] = r > j ? cmul(a[j * rs + rl], scale) : (r == j ? cmake(1.0, 0.0) : cmake(0.0, 0.0));
    }
    if (tid > j && tid < nb) {   // w_cc = v^H a_cc = a[j][cc] + conj(scale) * sum_{r>j} conj(x_r) a[r][cc]
      cplx w = rowj[tid];
      cfma(w, cconj(scale), Ptot[tid]);
      wv[tid] = cmul(ctau, w);
    }
    if (rank == 0 && tid < j) {  // g = V^H v_j = conj(V[j][t]) + scale * sum_{r>j} conj(V[r][t]) x_r
      cplx g = cconj(rowj[tid]);
      cfma(g, scale, Ptot[tid]);
      gsm[tid] = g;
    }
    if (tid == 0) { tau_s[j] = tau; beta_s[j] = beta; }
    __syncthreads();
    // ---- phase D: a_cc -= conj(tau) w_cc v  (rows r >= j), store v and beta in column j
    const int ncol = nb - j - 1;
    for (int e = tid; e < ncol * nloc; e += blockDim.x) {
      int cc = j + 1 + e / nloc, rl = e % nloc;
      if (r_begin + rl >= j) {
        cplx t = a[cc * rs + rl];
        cplx p = cmul(vloc[rl], wv[cc]);
        a[cc * rs + rl] = csub(t, p);
      }
    }
    for (int rl = tid; rl < nloc; rl += blockDim.x) {
      int r = r_begin + rl;
      if (r > j) a[j * rs + rl] = vloc[rl];
      else if (r == j) a[j * rs + rl] = cmake(beta, 0.0);
    }
    if (rank == 0 && warp == 7) {  // T(0:j,j) = -tau * T(0:j,0:j) * g ; T(j,j) = tau   (zlarft, forward/columnwise)
      if (lane < j) {
        cplx acc = cmake(0.0, 0.0);
        for (int k = lane; k < j; ++k) cfma(acc, Tsm[lane][k], gsm[k]);
        Tsm[lane][j] = cneg(cmul(tau, acc));
      }
      if (lane == 0) Tsm[j][j] = tau;
    }
    __syncthreads();
  }

  for (int c = 0; c < nb; ++c)
    for (int r = tid; r < nloc; r += blockDim.x) A[(size_t)c * lda + r_begin + r] = a[c * rs + r];
  if (rank == 0) {
    for (int e = tid; e < QR_NB * QR_NB; e += blockDim.x) {
      int i = e % QR_NB, k = e / QR_NB;
      Tout[e] = (i < nb && k < nb) ? Tsm[i][k] : cmake(0.0, 0.0);
    }
    if (tid < nb) { tau_out[tid] = tau_s[tid]; dabs_out[tid] = fabs(beta_s[tid]); }
  }
  cl.sync();  // no CTA may exit while others may still read its shared memory
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

static int launch_panel(cudaStream_t st, cplx* A, int lda, int m, int nb, cplx* tau, double* dabs, cplx* T) {
  const int rs = (m + QR_CL - 1) / QR_CL;
  const size_t smem = sizeof(cplx) * ((size_t)QR_NB * rs + rs);
  static size_t smem_lim = 0;
  if (smem_lim == 0 && set_max_dynamic_smem(qr_panel_kernel, &smem_lim)) return -1;
  if (smem > smem_lim) { snprintf(g_errbuf, sizeof(g_errbuf), "qr panel: m=%d too large", m); return -1; }
  qr_panel_kernel<<<QR_CL, 256, smem, st>>>(A, lda, m, nb, tau, dabs, T);
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

int qr_factor(cudaStream_t st, cplx* A, int lda, int n, cplx* tau, double* dabs, cplx* tfac, cplx* rhs, int ldr,
              int nrhs, int num_sms) {
  (void)num_sms;
  for (int j0 = 0, k = 0; j0 < n; j0 += QR_NB, ++k) {
    const int nb = min(QR_NB, n - j0), m = n - j0;
    cplx* P = A + (size_t)j0 * lda + j0;
    cplx* T = tfac + (size_t)k * QR_NB * QR_NB;
    if (launch_panel(st, P, lda, m, nb, tau + j0, dabs + j0, T)) return -1;
    if (launch_larfb(st, P, lda, m, T, 1, P + (size_t)nb * lda, lda, n - j0 - nb)) return -1;
    if (rhs && launch_larfb(st, P, lda, m, T, 1, rhs + j0, ldr, nrhs)) return -1;
  }
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
