#include "qr.cuh"
#include "zgemm.cuh"
#include <cstdlib>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

// =====================================================================================================
// Panel factorization: one thread-block cluster of QR_CL CTAs; each CTA owns a slab of rows of the m x nb
// panel and keeps it in REGISTERS: thread (c = lane, g = warp) holds column c on the local rows {g + NW*t}.
// The rows of column j that warp g needs are exactly the ones lane j of the same warp holds, so the current
// column travels through a per-warp shared-memory line (one predicated store, broadcast loads, __syncwarp) and
// the slab itself is never re-read from shared memory (the earlier shared-memory-resident version was bound by
// 3 passes x 64 KB of shared-memory traffic per column).  Per column ONE cluster barrier: every CTA reduces the
// dot products of the current column with all panel columns over its rows (rows below the diagonal only) and
// pushes them, plus row j if it owns it, into every CTA's shared memory (distributed shared memory stores).
// After the barrier every CTA derives beta, tau, v and the update coefficients redundantly from local data.
// The dots with the already-finished columns give V^H v_j, i.e. the compact-WY T factor, for free.
// =====================================================================================================
#define QR_LDA 33
// cluster rank that accumulates the compact-WY T factor: rank 0 (which owns only the 32-row diagonal block) while the
// other CTAs have long slabs, rank 1 once the slabs are short and rank 0 (masked inner loops, pushes row j) is the slowest
__host__ __device__ __forceinline__ int panel_t_rank(int m) { return (m - QR_NB) >= (QR_CL - 1) * 48 ? 0 : 1; }
#define QR_TX_BYTES ((QR_CL + 1) * QR_NB * 16)   // per column and CTA: QR_CL partial-dot vectors + row j

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
// 16-byte store into the shared memory of another CTA of the cluster, counted on that CTA's transaction barrier
__device__ __forceinline__ void st_async_c16(uint32_t raddr, cplx v, uint32_t rbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f64 [%0], {%1, %2}, [%3];"
               ::"r"(raddr), "d"(v.x), "d"(v.y), "r"(rbar) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "WAIT_LOOP:\n"
#ifdef QRV_ACQ
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%0], %1;\n"
#else
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
#endif
      "@P1 bra.uni WAIT_DONE;\n"
      "bra.uni WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(bar), "r"(parity) : "memory");
}

// Shared-memory state of one CTA of the panel cluster (static part)
template <int NW>
struct PanelSmem {
  cplx part[NW][QR_NB];               // per-warp partial dots
  cplx colbuf[NW][20];                // current column on the rows of warp g (zero on rows <= j)
  cplx rowl[QR_NB];                   // row j of the panel (rank 0)
  cplx xch[2][QR_CL][QR_NB];          // [parity][source CTA][column]: partial dots of every CTA
  cplx rowv[2][QR_NB];                // [parity][column]: row j of the panel (pushed by rank 0)
  cplx Tsm[QR_NB][QR_NB + 1];         // rank 0: compact-WY T; rows {g, g+8, g+16, g+24} belong to warp g
  cplx gsm[8][QR_NB];                 // V^H v_j, one private copy per warp (no CTA barrier between write and use)
  cplx tau_s[QR_NB];
  double beta_s[QR_NB];
  unsigned long long full[2];         // transaction barriers, one per parity
};

// The column loop for one CTA.  GENERAL = true is rank 0, which owns the first QR_NB rows of the panel (the diagonal
// block: rows at or above the diagonal are masked, row j is pushed to the cluster, the T factor is accumulated);
// GENERAL = false are the CTAs below, whose inner loops are bare multiply-adds.  A finished column c stays UNSCALED in
// registers (v = sc_c * x below the diagonal); the factor is applied to its dot products and at the final store.
template <int NW, int T, bool GENERAL, bool PROF>
__device__ __forceinline__ void panel_body(PanelSmem<NW>& S, cplx* a, cg::cluster_group& cl, cplx* __restrict__ A, int lda,
                                           int m, int nb, int r_begin, int nloc, cplx* __restrict__ tau_out,
                                           double* __restrict__ dabs_out, cplx* __restrict__ Tout, long long* __restrict__ prof) {
  long long pc[6] = {0, 0, 0, 0, 0, 0}, tprev = clock64();
#define PSTAMP(k) do { if (PROF && GENERAL && prof) { long long t_ = clock64(); pc[k] += t_ - tprev; tprev = t_; } } while (0)
  const int rank = (int)cl.block_rank();
  const int tid = threadIdx.x, g = tid >> 5, c = tid & 31;
  const bool doT = (rank == panel_t_rank(m));     // the CTA that accumulates the compact-WY T factor
  cplx x[T];
#pragma unroll
  for (int t = 0; t < T; ++t) {
    const int rl = g + NW * t;
    x[t] = rl < nloc ? a[rl * QR_LDA + c] : cmake(0.0, 0.0);
  }
  if (c == 0) {
#pragma unroll
    for (int t = 0; t < T; ++t) S.colbuf[g][t] = (!GENERAL || g + NW * t > 0) ? x[t] : cmake(0.0, 0.0);
  }
  // remote addresses of my slots in CTA g's exchange buffers (warp g pushes to CTA g)
  const uint32_t dst_rank = (uint32_t)(g < QR_CL ? g : 0);
  const uint32_t r_xch = mapa_u32(smem_u32(&S.xch[0][rank][c]), dst_rank);
  const uint32_t r_row = mapa_u32(smem_u32(&S.rowv[0][c]), dst_rank);
  const uint32_t r_bar = mapa_u32(smem_u32(&S.full[0]), dst_rank);
  const uint32_t l_bar = smem_u32(&S.full[0]);
  cl.sync();   // every CTA's barriers are initialised before anybody stores into them

  cplx tau_prev = cmake(0.0, 0.0);
  cplx sc = cmake(1.0, 0.0);                     // scale of my column once it is finished
  for (int j = 0; j < nb; ++j) {
    const int par = j & 1;
    if (tid == 0)
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(l_bar + 8 * par), "r"(QR_TX_BYTES) : "memory");
    // ---- phase A: t_c = sum_{r>j} conj(a[r][j]) a[r][c] over my rows
    cplx d[T];
    {
      cplx acc0 = cmake(0.0, 0.0), acc1 = acc0, acc2 = acc0, acc3 = acc0;
#pragma unroll
      for (int t = 0; t < T; ++t) d[t] = S.colbuf[g][t];
#pragma unroll
      for (int t = 0; t < T; ++t) {
        if ((t & 3) == 0) cfma_conj(acc0, d[t], x[t]);
        else if ((t & 3) == 1) cfma_conj(acc1, d[t], x[t]);
        else if ((t & 3) == 2) cfma_conj(acc2, d[t], x[t]);
        else cfma_conj(acc3, d[t], x[t]);
      }
      S.part[g][c] = cadd(cadd(acc0, acc1), cadd(acc2, acc3));
      if (GENERAL) {
#pragma unroll
        for (int t = 0; t < T; ++t)
          if (g + NW * t == j) S.rowl[c] = (c < j) ? cmul(x[t], sc) : x[t];
      }
    }
    PSTAMP(0);
    __syncthreads();
    if (g < QR_CL) {
      // warp g sums the NW partials of column c and pushes the result (and row j, if mine) to CTA g of the cluster
      cplx p0 = cadd(S.part[0][c], S.part[1][c]), p1 = cadd(S.part[2][c], S.part[3][c]);
      cplx p2 = cadd(S.part[4][c], S.part[5][c]), p3 = cadd(S.part[6][c], S.part[7][c]);
      if (NW > 8) {
        p0 = cadd(p0, cadd(S.part[8 % NW][c], S.part[9 % NW][c])); p1 = cadd(p1, cadd(S.part[10 % NW][c], S.part[11 % NW][c]));
        p2 = cadd(p2, cadd(S.part[12 % NW][c], S.part[13 % NW][c])); p3 = cadd(p3, cadd(S.part[14 % NW][c], S.part[15 % NW][c]));
      }
      cplx tot = cadd(cadd(p0, p1), cadd(p2, p3));
      if (c < j) tot = cmul(sc, tot);            // finished column: its rows are stored unscaled
      st_async_c16(r_xch + par * (QR_CL * QR_NB * 16), tot, r_bar + 8 * par);
      if (GENERAL) st_async_c16(r_row + par * (QR_NB * 16), S.rowl[c], r_bar + 8 * par);
    }
    PSTAMP(1);
#ifndef QRV_NOT
    if (doT && j > 0 && g < 8) {
      // T(0:jp,jp) = -tau_jp * T(0:jp,0:jp) * g for the previous column jp = j-1 (zlarft, forward/columnwise) in the
      // shadow of the exchange: warp g owns rows i = g + 8q; lane = (q, h) sums k = h, h+8, ... and the 8 h-lanes combine
      const int jp = j - 1, q = c >> 3, h = c & 7, i = g + 8 * q;
      cplx acc = cmake(0.0, 0.0);
      if (i < jp) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const int k = h + 8 * kk;
          if (k >= i && k < jp) cfma(acc, S.Tsm[i][k], S.gsm[g][k]);
        }
      }
#pragma unroll
      for (int o = 1; o < 8; o <<= 1) {
        acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o);
        acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
      }
      if (h == 0) {
        if (i < jp) S.Tsm[i][jp] = cneg(cmul(tau_prev, acc));
        else if (i == jp) S.Tsm[jp][jp] = tau_prev;
      }
      __syncwarp();
    }
#endif
    mbar_wait_cluster(l_bar + 8 * par, (uint32_t)((j >> 1) & 1));
    // ---- phase C: totals and reflector parameters (every thread, redundantly, from local shared memory)
    PSTAMP(2);
    cplx tc, tj;
    {
      const cplx c0 = cadd(S.xch[par][0][c], S.xch[par][1][c]), c1 = cadd(S.xch[par][2][c], S.xch[par][3][c]);
      const cplx c2 = cadd(S.xch[par][4][c], S.xch[par][5][c]), c3 = cadd(S.xch[par][6][c], S.xch[par][7][c]);
      const double j0 = S.xch[par][0][j].x + S.xch[par][1][j].x, j1 = S.xch[par][2][j].x + S.xch[par][3][j].x;
      const double j2 = S.xch[par][4][j].x + S.xch[par][5][j].x, j3 = S.xch[par][6][j].x + S.xch[par][7][j].x;
      tc = cadd(cadd(c0, c1), cadd(c2, c3));
      tj = cmake((j0 + j1) + (j2 + j3), 0.0);
    }
    const cplx alpha = S.rowv[par][j], rowc = S.rowv[par][c];
    const double aa = alpha.x * alpha.x + alpha.y * alpha.y;
    const double nrm2 = aa + tj.x;
    double beta;
    cplx tau, scale;
    if (nrm2 == 0.0) {
      beta = 0.0; tau = cmake(0.0, 0.0); scale = cmake(0.0, 0.0);
    } else {
#ifdef QRV_FAKEMATH
      const double nrm = nrm2 * 0.5;
      beta = alpha.x >= 0.0 ? -nrm : nrm;
      const double id = 0.3 * (aa + nrm2 + 2.0 * fabs(alpha.x) * nrm);
      const double ib = 0.3 * beta;
#else
      const double nrm = sqrt(nrm2);
      beta = alpha.x >= 0.0 ? -nrm : nrm;
      // |alpha - beta|^2 = 2 |alpha|^2 + t_j + 2 |Re alpha| nrm: independent of the division that gives 1/beta
      const double id = 1.0 / (aa + nrm2 + 2.0 * fabs(alpha.x) * nrm);
      const double ib = 1.0 / beta;
#endif
      const double ar = alpha.x - beta;
      tau = cmake((beta - alpha.x) * ib, -alpha.y * ib);
      scale = cmake(ar * id, -alpha.y * id);
    }
    // w_c = conj(tau) * v^H a_c = conj(tau) * (a[j][c] + conj(scale) * t_c)      (c > j)
    cplx wc = rowc;
    cfma(wc, cconj(scale), tc);
    wc = cmul(cconj(tau), wc);
    if (doT && g < 8) {
      // g_c = V_c^H v_j = conj(V[j][c]) + scale * conj(t_c)   (c < j)
      cplx gg = cconj(rowc);
      cfma(gg, scale, cconj(tc));
      S.gsm[g][c] = c < j ? gg : cmake(0.0, 0.0);
      if (tid == 0) { S.tau_s[j] = tau; S.beta_s[j] = beta; }
    }
    tau_prev = tau;
    if (c == j) sc = scale;
    PSTAMP(3);
    // ---- phase D: a[r][c] -= v_r w_c (r >= j, c > j) with v_r = scale * a[r][j] (r > j), v_j = 1:  x <- x + d * q
    {
      const cplx q = c > j ? cneg(cmul(scale, wc)) : cmake(0.0, 0.0);
      if (NW > 8) {   // 512-thread variant: 128 registers per thread, so the column is re-read instead of kept
#pragma unroll
        for (int t = 0; t < T; ++t) d[t] = S.colbuf[g][t];
      }
#pragma unroll
      for (int t = 0; t < T; ++t) {
        cfma(x[t], d[t], q);
        if (GENERAL) {
          const int r = g + NW * t;
          if (r == j && c >= j) x[t] = (c == j) ? cmake(beta, 0.0) : csub(x[t], wc);
        }
      }
      if (c == j + 1) {
#pragma unroll
        for (int t = 0; t < T; ++t) S.colbuf[g][t] = (!GENERAL || g + NW * t > j + 1) ? x[t] : cmake(0.0, 0.0);
      }
      __syncwarp();
    }
    PSTAMP(4);
  }
  if (PROF && GENERAL && prof && tid == 0) { for (int q = 0; q < 5; ++q) prof[q] += pc[q]; prof[5] += nb; }
  if (doT && g < 8) {   // T column of the last reflector
    const int jp = nb - 1, q = c >> 3, h = c & 7, i = g + 8 * q;
    cplx acc = cmake(0.0, 0.0);
    if (i < jp) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const int k = h + 8 * kk;
        if (k >= i && k < jp) cfma(acc, S.Tsm[i][k], S.gsm[g][k]);
      }
    }
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
      acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o);
      acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
    }
    if (h == 0) {
      if (i < jp) S.Tsm[i][jp] = cneg(cmul(tau_prev, acc));
      else if (i == jp) S.Tsm[jp][jp] = tau_prev;
    }
  }
#pragma unroll
  for (int t = 0; t < T; ++t) {
    const int rl = g + NW * t;
    if (rl < nloc) a[rl * QR_LDA + c] = (!GENERAL || rl > c) ? cmul(x[t], sc) : x[t];
  }
  __syncthreads();
  for (int e = tid; e < nloc * nb; e += NW * 32) {
    const int rl = e % nloc, cc = e / nloc;
    A[(size_t)cc * lda + r_begin + rl] = a[rl * QR_LDA + cc];
  }
  if (doT) {
    for (int e = tid; e < QR_NB * QR_NB; e += NW * 32) {
      int i = e % QR_NB, k = e / QR_NB;
      Tout[e] = (i < nb && k < nb) ? S.Tsm[i][k] : cmake(0.0, 0.0);
    }
    if (tid < nb) { tau_out[tid] = S.tau_s[tid]; dabs_out[tid] = fabs(S.beta_s[tid]); }
  }
  cl.sync();  // no CTA may exit while others may still write into its shared memory
}

// Row split: rank 0 owns rows [0, QR_NB) (or all m of them if fewer), the rest is dealt evenly to ranks 1..QR_CL-1.
__host__ __device__ __forceinline__ int panel_rows_below(int m) { return (max(0, m - QR_NB) + QR_CL - 2) / (QR_CL - 1); }

template <int NW, int T, bool PROF>
__global__ void __cluster_dims__(QR_CL, 1, 1) __launch_bounds__(NW * 32)
qr_panel_kernel(cplx* __restrict__ A, int lda, int m, int nb, cplx* __restrict__ tau_out,
                double* __restrict__ dabs_out, cplx* __restrict__ Tout, long long* __restrict__ prof) {
  cg::cluster_group cl = cg::this_cluster();
  const int rank = (int)cl.block_rank();
  const int rs1 = panel_rows_below(m);
  const int r_begin = rank == 0 ? 0 : min(m, QR_NB + (rank - 1) * rs1);
  const int nloc = rank == 0 ? min(m, QR_NB) : max(0, min(m, r_begin + rs1) - r_begin);
  const int tid = threadIdx.x;

  extern __shared__ __align__(16) unsigned char smem_raw[];
  cplx* a = reinterpret_cast<cplx*>(smem_raw);   // staging for the coalesced load / store: a[rl*QR_LDA + c]
  __shared__ __align__(16) PanelSmem<NW> S;

  // global -> shared: consecutive threads along rows (coalesced), 32 columns
  for (int e = tid; e < nloc * QR_NB; e += NW * 32) {
    const int rl = e % nloc, cc = e / nloc;
    a[rl * QR_LDA + cc] = cc < nb ? A[(size_t)cc * lda + r_begin + rl] : cmake(0.0, 0.0);
  }
  if (rank == panel_t_rank(m))
    for (int e = tid; e < QR_NB * (QR_NB + 1); e += NW * 32) (&S.Tsm[0][0])[e] = cmake(0.0, 0.0);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&S.full[0])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&S.full[1])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (rank == 0) panel_body<NW, QR_NB / NW, true, PROF>(S, a, cl, A, lda, m, nb, r_begin, nloc, tau_out, dabs_out, Tout, prof);
  else panel_body<NW, T, false, PROF>(S, a, cl, A, lda, m, nb, r_begin, nloc, tau_out, dabs_out, Tout, prof);
}

// =====================================================================================================
// PAIRED panel factorization for matrices with the antiunitary flavour symmetry S = [[A, B], [-conj(B), conj(A)]]
// (executable specification, formula by formula: oracle/experiments/paired_panel_spec.py).  The matrix is held as its LEFT
// HALF with pair-interleaved rows (rows 2i, 2i+1 = natural rows i, i+n/2 = one quaternion); column c+n/2 is the partner
// psi(x)[2i] = -conj(x[2i+1]), psi(x)[2i+1] = conj(x[2i]) of column c.  One pair-step eliminates a column AND its partner
// with the two mutually orthogonal reflectors v, psi(v):  H = 1 - tau (v v^H + psi(v) psi(v)^H), real tau -- so a
// factorization has n/2 sequential steps instead of n, each with the same exchange (one per step) and the same number of
// multiply-adds as a step of qr_panel_kernel.  A panel = 16 pair-steps = 32 reflectors V = [v_0, psi(v_0), v_1, ...], written
// out EXPLICITLY (m x 32, 2x2 blocks on the diagonal), so larfb_kernel / larfb_cluster_kernel apply it unchanged (vmode 1).
// Thread (lane, warp g): column lane & 15, component lane >> 4 (0: even rows "e", 1: odd rows "o") on the quaternion rows
// {g + 8 t} of the CTA's slab, in registers.  Raw dot products over the rows strictly below row pair j,
//     D1_c = sum conj(x_e) c_e + conj(x_o) c_o,      D2_c = sum x_e c_o - x_o c_e,
// are all a step needs besides row pair j of the panel; lanes c / c+16 end up with D1_c / D2_c of their column (one shuffle),
// which is exactly the 32-slot exchange vector of the unpaired kernel.
// =====================================================================================================
#define QP_NP 16       // pair-steps per panel
#define QP_TMAX 24     // quaternion rows per thread, upper bound
struct PairedSmem {
  cplx part[8][QR_NB];                // per-warp partial dots: slot c = D1_c, slot c+16 = D2_c
  cplx colbuf[8][2][QP_TMAX];         // current column on the quaternion rows of warp g: [component][t] (zero on rows <= j)
  cplx rowl[QR_NB];                   // row pair j of the panel (rank 0): slot c = e, slot c+16 = o
  cplx xch[2][QR_CL][QR_NB];          // [parity][source CTA][slot]
  cplx rowv[2][QR_NB];                // [parity][slot]: row pair j (pushed by rank 0)
  cplx Tsm[QR_NB][QR_NB + 1];         // compact-WY T; rows {2g, 2g+1, 2g+16, 2g+17} belong to warp g
  cplx gsm[2][QR_NB];                 // [parity of the step]: V^H v_j, written by warp 0, read by all in the next step's exchange shadow
  cplx ab[2][QR_NB];                  // update coefficients alpha / beta of the step per lane (warp 0 -> all warps)
  cplx sc4[4];                        // u_j and x_j on row pair j (e, o)
  cplx udiag[QR_NB];                  // v_j on its own row pair: slot j = e, slot j+16 = o
  double tau_s[QP_NP], nx_s[QP_NP], rn_s[QP_NP];
  unsigned long long full[2];
};

// 1/sqrt(a) and 1/a for normal a > 0: hardware approximation (2^-22 / 2^-23) + two Newton steps
__device__ __forceinline__ double rsqrt_nr(double a) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
  double e = fma(-(a * y), y, 1.0);
  y = fma(0.5 * y, e, y);
  e = fma(-(a * y), y, 1.0);
  return fma(0.5 * y, e, y);
}
__device__ __forceinline__ double rcp_nr(double a) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
  double e = fma(-a, y, 1.0);
  y = fma(y, e, y);
  e = fma(-a, y, 1.0);
  return fma(y, e, y);
}

// The two columns 2j, 2j+1 of T (zlarft, forward / columnwise) of pair-step j.  Warp g owns the rows {2g, 2g+1, 2g+16, 2g+17} (whole
// quaternion rows): column 2j is T(i, 2j) = -tau sum_{k=i}^{2j-1} T(i,k) gv(k); column 2j+1 follows from the quaternion structure
// of T (exact: paired_panel_spec.py), T(2c, 2j+1) = -conj(T(2c+1, 2j)), T(2c+1, 2j+1) = conj(T(2c, 2j)), T(2j, 2j+1) = 0.
__device__ __forceinline__ void paired_t_columns(PairedSmem& S, int g, int lane, int j, double tau, const cplx* gv) {
  const int q = lane >> 3, hh = lane & 7, i = 2 * g + (q & 1) + 16 * (q >> 1), col = 2 * j;
  cplx acc = cmake(0.0, 0.0);
  if (i < col) {
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const int k = hh + 8 * kk;
      if (k >= i && k < col) cfma(acc, S.Tsm[i][k], gv[k]);
    }
  }
#pragma unroll
  for (int o = 1; o < 8; o <<= 1) {
    acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o);
    acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
  }
  // lanes of row i and of its partner row i^1 are 8 apart (q ^ 1): fetch the partner's sum
  const cplx oth = cmake(__shfl_xor_sync(0xffffffffu, acc.x, 8), __shfl_xor_sync(0xffffffffu, acc.y, 8));
  if (hh == 0) {
    if (i < col) {
      S.Tsm[i][col] = cmake(-tau * acc.x, -tau * acc.y);
      // T(i, col+1) from T(i^1, col) = -tau oth:  i even: -conj(.),  i odd: conj(.)
      S.Tsm[i][col + 1] = (i & 1) ? cmake(-tau * oth.x, tau * oth.y) : cmake(tau * oth.x, -tau * oth.y);
    } else if (i == col) {
      S.Tsm[i][i] = cmake(tau, 0.0);
      S.Tsm[i][i + 1] = cmake(0.0, 0.0);
    } else if (i == col + 1) {
      S.Tsm[i][i] = cmake(tau, 0.0);
    }
  }
  __syncwarp();
}

template <int T, bool GENERAL, bool PROF>
__device__ __forceinline__ void panel_body_paired(PairedSmem& S, cg::cluster_group& cl, cplx* __restrict__ A, int lda,
                                                  cplx* __restrict__ Vout, int ldv, int m, int np, int r_begin, int nloc,
                                                  double* __restrict__ dabs_out, int dabs_dup, cplx* __restrict__ Tout,
                                                  long long* __restrict__ prof, int prof_rank) {
  constexpr int NW = 8;
  // PROF instantiation only: cycle stamps of thread 0 of cluster rank `prof_rank`: [0] dots [1] reduce + push [2] T columns
  // [3] wait for the exchange [4] parameters [5] update [6] epilogue (stores) [7] pair-steps
  long long pcyc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tprev = PROF ? clock64() : 0;
  const bool stamp = PROF && prof != nullptr && (int)cl.block_rank() == prof_rank && threadIdx.x == 0;
#define QSTAMP(k) do { if (PROF && stamp) { const long long t_ = clock64(); pcyc[k] += t_ - tprev; tprev = t_; } } while (0)
  // sub-stamps of phase C (prof[8..12]); `dep` makes the stamp wait for the value it is meant to time
  long long psub[5] = {0, 0, 0, 0, 0}, tsub = 0;
#define QSUB(k, dep) do { if (PROF && stamp) { if ((dep) == 1.2345e-300) tsub++; const long long t_ = clock64(); psub[(k) - 8] += t_ - tsub; tsub = t_; } } while (0)
  constexpr bool KEEP = (T <= 10);               // current column kept in registers between the dot and the update phase
  const int rank = (int)cl.block_rank();
  const int tid = threadIdx.x, g = tid >> 5, lane = tid & 31;
  const int col = lane & 15, comp = lane >> 4;
  const bool doT = (rank == panel_t_rank(m));
  cplx x[T];
  {
    const cplx* Ac = A + (size_t)col * lda + r_begin;
#pragma unroll
    for (int t = 0; t < T; ++t) {
      const int rl = 2 * (g + NW * t) + comp;
      x[t] = (rl < nloc && col < np) ? Ac[rl] : cmake(0.0, 0.0);
    }
  }
  if (col == 0) {
#pragma unroll
    for (int t = 0; t < T; ++t) S.colbuf[g][comp][t] = (!GENERAL || g + NW * t > 0) ? x[t] : cmake(0.0, 0.0);
  }
  const uint32_t dst_rank = (uint32_t)g;         // warp g pushes to CTA g (NW == QR_CL)
  const uint32_t r_xch = mapa_u32(smem_u32(&S.xch[0][rank][lane]), dst_rank);
  const uint32_t r_row = mapa_u32(smem_u32(&S.rowv[0][lane]), dst_rank);
  const uint32_t r_bar = mapa_u32(smem_u32(&S.full[0]), dst_rank);
  const uint32_t l_bar = smem_u32(&S.full[0]);
  cl.sync();   // every CTA's barriers are initialised before anybody stores into them

  double tau_prev = 0.0;
  double sc = 1.0;                               // 1/|x| of my column once it is finished (v = sc * u)
  if (PROF) tprev = clock64();
  for (int j = 0; j < np; ++j) {
    const int par = j & 1;
    if (tid == 0)
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(l_bar + 8 * par), "r"(QR_TX_BYTES) : "memory");
    // ---- phase A: raw dots of the current column with my column over my rows (rows <= j are zero in the column buffer)
    cplx dA[KEEP ? T : 1], dB[KEEP ? T : 1];
    (void)dA; (void)dB;
    {
      cplx p1a = cmake(0.0, 0.0), p1b = p1a, p2a = p1a, p2b = p1a;
#pragma unroll
      for (int t = 0; t < T; ++t) {
        const cplx ca = S.colbuf[g][comp][t], cb = S.colbuf[g][comp ^ 1][t];
        if constexpr (KEEP) { dA[t] = ca; dB[t] = cb; }
        if (t & 1) { cfma_conj(p1b, ca, x[t]); cfma(p2b, cb, x[t]); }
        else       { cfma_conj(p1a, ca, x[t]); cfma(p2a, cb, x[t]); }
      }
      cplx p1 = cadd(p1a, p1b), p2 = cadd(p2a, p2b);
      if (comp == 0) p2 = cneg(p2);              // D2 = sum x_e c_o - x_o c_e: the e-lanes hold -(x_o c_e)
      const cplx o1 = cmake(__shfl_xor_sync(0xffffffffu, p1.x, 16), __shfl_xor_sync(0xffffffffu, p1.y, 16));
      const cplx o2 = cmake(__shfl_xor_sync(0xffffffffu, p2.x, 16), __shfl_xor_sync(0xffffffffu, p2.y, 16));
      S.part[g][lane] = comp == 0 ? cadd(p1, o1) : cadd(p2, o2);
      if (GENERAL) {
#pragma unroll
        for (int t = 0; t < T; ++t)
          if (g + NW * t == j) S.rowl[lane] = (col < j) ? cscale(x[t], sc) : x[t];
      }
    }
    QSTAMP(0);
    __syncthreads();
    {
      // warp g sums the 8 partials of slot `lane` and pushes the result (and row pair j, if mine) to CTA g of the cluster
      const cplx p0 = cadd(S.part[0][lane], S.part[1][lane]), p1 = cadd(S.part[2][lane], S.part[3][lane]);
      const cplx p2 = cadd(S.part[4][lane], S.part[5][lane]), p3 = cadd(S.part[6][lane], S.part[7][lane]);
      cplx tot = cadd(cadd(p0, p1), cadd(p2, p3));
      if (col < j) tot = cscale(tot, sc);        // finished column: its rows are stored unscaled
      st_async_c16(r_xch + par * (QR_CL * QR_NB * 16), tot, r_bar + 8 * par);
      if (GENERAL) st_async_c16(r_row + par * (QR_NB * 16), S.rowl[lane], r_bar + 8 * par);
    }
    QSTAMP(1);
    if (doT && j > 0) {
      // the two T columns of the previous pair-step, in the shadow of the exchange
      paired_t_columns(S, g, lane, j - 1, tau_prev, S.gsm[par ^ 1]);
    }
    QSTAMP(2);
    mbar_wait_cluster(l_bar + 8 * par, (uint32_t)((j >> 1) & 1));
    QSTAMP(3);
    if (PROF) tsub = tprev;
    // ---- phase C: totals and reflector parameters.  Identical for every warp of the CTA: warp 0 alone computes them and
    // broadcasts the per-column update coefficients through shared memory (a latency-bound dependent chain, ~800 cycles; eight
    // warps evaluating it redundantly took as long and only added issue pressure).
    if (g == 0) {
      cplx tc;
      {
        const cplx c0 = cadd(S.xch[par][0][lane], S.xch[par][1][lane]), c1 = cadd(S.xch[par][2][lane], S.xch[par][3][lane]);
        const cplx c2 = cadd(S.xch[par][4][lane], S.xch[par][5][lane]), c3 = cadd(S.xch[par][6][lane], S.xch[par][7][lane]);
        tc = cadd(cadd(c0, c1), cadd(c2, c3));
      }
      QSUB(8, tc.x);
      const double tj = __shfl_sync(0xffffffffu, tc.x, j);        // D1 of the current column = |x|^2 below row pair j
      const cplx tother = cmake(__shfl_xor_sync(0xffffffffu, tc.x, 16), __shfl_xor_sync(0xffffffffu, tc.y, 16));
      const cplx D1 = comp == 0 ? tc : tother, D2 = comp == 0 ? tother : tc;
      const cplx xe0 = S.rowv[par][j], xo0 = S.rowv[par][j + 16];
      const cplx ce0 = S.rowv[par][col], co0 = S.rowv[par][col + 16];
      const double q2 = cabs2(xe0) + cabs2(xo0);
      const double nrm2 = q2 + tj;
      QSUB(9, nrm2 + ce0.x + D2.x);
      // The square roots and divisions of the reflector are the longest dependent chain of a step (sqrt, sqrt, 1/x, x/y one after
      // the other: ~1000 cycles measured).  Reciprocal square roots from the hardware approximation + two Newton steps (1-2 ulp,
      // no slow-path branches, the two chains interleave): |x| = nrm2 r, 1/|x| = r, |x_j| = q2 rq, |x|/|x_j| = |x| rq, and ONE
      // reciprocal 1 / (|x| + |x_j|).  A relative error eps in |x| leaves H unitary up to O(eps), like any rounding of tau.
      const bool nzx = nrm2 > 0.0, nzq = q2 > 0.0;
      const double r = nzx ? rsqrt_nr(nrm2) : 0.0, rq = nzq ? rsqrt_nr(q2) : 0.0;
      const double nx = nrm2 * r, q = q2 * rq;
      const double gi = nzx ? rcp_nr(nx + q) : 0.0;
      const double f = r * gi;                     // 1 / (|x| (|x| + |x_j|))
      const double rnj = r;                        // 1 / |x|
      const double tau = nx * gi;                  // |x| / (|x| + |x_j|)
      QSUB(10, f);
      // products with row pair j that do not depend on the scale: t1 = x_j^H c_j, t2 = psi(x_j)^H c_j (in parallel with the chain above)
      cplx t1 = cmake(0.0, 0.0), t2 = t1;
      cfma_conj(t1, xe0, ce0); cfma_conj(t1, xo0, co0);
      cfma(t2, xe0, co0); cfma(t2, cneg(xo0), ce0);
      // u_j = s1 x_j, s1 = 1 + |x| / |x_j|;  x_j = 0: u_j = (|x|, 0), i.e. u^H c = D1 + |x| c_e, psi(u)^H c = D2 + |x| c_o
      const double s1 = nzq ? fma(nx, rq, 1.0) : nx;
      if (!nzq) { t1 = ce0; t2 = co0; }
      const cplx ue0 = nzq ? cscale(xe0, s1) : cmake(nx, 0.0), uo0 = nzq ? cscale(xo0, s1) : cmake(0.0, 0.0);
      // uc = u^H c = D1 + s1 t1,   pc = psi(u)^H c = D2 + s1 t2
      const cplx uc = cmake(fma(s1, t1.x, D1.x), fma(s1, t1.y, D1.y)), pc = cmake(fma(s1, t2.x, D2.x), fma(s1, t2.y, D2.y));
      if (doT && comp == 0) {
        // Gram entries against the finished columns (their dots and row entries arrived scaled by 1/|x_c|)
        const bool fin = col < j;
        const cplx z = cmake(0.0, 0.0);
        S.gsm[par][2 * col] = fin ? cmake(rnj * uc.x, -rnj * uc.y) : z;        // v_c^H v_j        =  conj(uc)
        S.gsm[par][2 * col + 1] = fin ? cmake(-rnj * pc.x, -rnj * pc.y) : z;   // psi(v_c)^H v_j   = -pc
        // (the entries against psi(v_j), conj(pc) and uc, are not needed: column 2j+1 of T follows from column 2j)
      }
      // update coefficients of my column: alpha = -f uc, beta = +-f pc (e / o lanes), zero for the finished columns
      const bool upd = col > j;
      const double fb = comp == 0 ? f : -f;
      S.ab[0][lane] = upd ? cmake(-f * uc.x, -f * uc.y) : cmake(0.0, 0.0);
      S.ab[1][lane] = upd ? cmake(fb * pc.x, fb * pc.y) : cmake(0.0, 0.0);
      if (lane == 0) {
        S.tau_s[j] = tau; S.nx_s[j] = nx; S.rn_s[j] = rnj;
        S.udiag[j] = cscale(ue0, rnj); S.udiag[j + 16] = cscale(uo0, rnj);
        S.sc4[0] = ue0; S.sc4[1] = uo0; S.sc4[2] = xe0; S.sc4[3] = xo0;
      }
      QSUB(11, uc.x + pc.x);
    }
    __syncthreads();
    const cplx alpha = S.ab[0][lane], beta = S.ab[1][lane];
    tau_prev = S.tau_s[j];
    QSUB(12, alpha.x);
    QSTAMP(4);
    // ---- phase D: c += A alpha + conj(B) beta on the rows below row pair j (A = my component of the current column,
    // B = the other one), alpha = -f uc, beta = +-f pc (e / o lanes); row pair j itself with u_j (rank 0)
#pragma unroll
    for (int t = 0; t < T; ++t) {
      cplx ca, cb;
      if constexpr (KEEP) { ca = dA[t]; cb = dB[t]; }
      else { ca = S.colbuf[g][comp][t]; cb = S.colbuf[g][comp ^ 1][t]; }
      cfma(x[t], ca, alpha);
      cfma_conj(x[t], cb, beta);
      if (GENERAL) {
        if (g + NW * t == j && col >= j) {
          const cplx ue0 = S.sc4[0], uo0 = S.sc4[1];
          const cplx myU = comp == 0 ? ue0 : uo0, otU = comp == 0 ? uo0 : ue0;
          if (col == j) x[t] = comp == 0 ? csub(S.sc4[2], ue0) : csub(S.sc4[3], uo0);
          else { cfma(x[t], myU, alpha); cfma_conj(x[t], otU, beta); }
        }
      }
    }
    if (col == j) sc = S.rn_s[j];
    if (!KEEP) __syncwarp();                     // everybody has re-read the column buffer before it is overwritten
    if (col == j + 1) {
#pragma unroll
      for (int t = 0; t < T; ++t) S.colbuf[g][comp][t] = (!GENERAL || g + NW * t > j + 1) ? x[t] : cmake(0.0, 0.0);
    }
    __syncwarp();
    QSTAMP(5);
  }
  if (doT) {
    paired_t_columns(S, g, lane, np - 1, tau_prev, S.gsm[(np - 1) & 1]);
  }
  __syncthreads();   // udiag / T complete
  // ---- results straight from the registers: lanes c / c+16 hold the two rows of a quaternion row, i.e. every (column, row pair)
  // is one aligned 32-byte sector.  R: rows at or above the quaternion diagonal keep their values, eliminated entries are exact
  // zeros.  V = [v_0, psi(v_0), v_1, ...] explicit: column 2c = v_c, column 2c+1: rows (2q, 2q+1) = (-conj(v_o), conj(v_e)).
  const int q_begin = r_begin >> 1;              // global quaternion row of my first local one (within the panel)
  if (col < np) {
    cplx* Rc = A + (size_t)col * lda + r_begin;
    cplx* V0 = Vout + (size_t)(2 * col) * ldv + r_begin;
    cplx* V1 = V0 + ldv;
#pragma unroll
    for (int t = 0; t < T; ++t) {
      const int ql = g + NW * t, rl = 2 * ql + comp;
      if (rl < nloc) {
        const int qg = q_begin + ql;
        Rc[rl] = (qg <= col) ? x[t] : cmake(0.0, 0.0);
        cplx v = cmake(0.0, 0.0);
        if (qg > col) v = cscale(x[t], sc);
        else if (qg == col) v = S.udiag[col + 16 * comp];
        V0[rl] = v;
        V1[rl ^ 1] = comp == 0 ? cmake(v.x, -v.y) : cmake(-v.x, v.y);   // my value feeds the OTHER row of the pair in the partner column
      }
    }
  }
  if (doT) {
    for (int e = tid; e < QR_NB * QR_NB; e += NW * 32) {
      const int i = e % QR_NB, k = e / QR_NB;
      Tout[e] = (i < 2 * np && k < 2 * np && i <= k) ? S.Tsm[i][k] : cmake(0.0, 0.0);
    }
    if (tid < np) { dabs_out[tid] = S.nx_s[tid]; if (dabs_dup) dabs_out[dabs_dup + tid] = S.nx_s[tid]; }
  }
  cl.sync();  // no CTA may exit while others may still write into its shared memory
  QSTAMP(6);
  if (PROF && stamp) { for (int q = 0; q < 7; ++q) prof[q] += pcyc[q]; prof[7] += np; for (int q = 0; q < 5; ++q) prof[8 + q] += psub[q]; }
#undef QSTAMP
#undef QSUB
}

// Row split of the paired panel: rank 0 owns the first 32 interleaved rows (16 quaternion rows, the diagonal block), the rest
// is dealt to ranks 1..7 in whole quaternion rows.
__host__ __device__ __forceinline__ int paired_rows_below(int m) { return 2 * ((max(0, m - QR_NB) / 2 + QR_CL - 2) / (QR_CL - 1)); }

template <int T, bool PROF>
__global__ void __cluster_dims__(QR_CL, 1, 1) __launch_bounds__(256)
qr_panel_paired_kernel(cplx* __restrict__ A, int lda, cplx* __restrict__ Vout, int ldv, int m, int np,
                       double* __restrict__ dabs_out, int dabs_dup, cplx* __restrict__ Tout, long long* __restrict__ prof, int prof_rank) {
  cg::cluster_group cl = cg::this_cluster();
  const int rank = (int)cl.block_rank();
  const int rs1 = paired_rows_below(m);
  const int r_begin = rank == 0 ? 0 : min(m, QR_NB + (rank - 1) * rs1);
  const int nloc = rank == 0 ? min(m, QR_NB) : max(0, min(m, r_begin + rs1) - r_begin);
  const int tid = threadIdx.x;

  __shared__ __align__(16) PairedSmem S;

  if (rank == panel_t_rank(m))
    for (int e = tid; e < QR_NB * (QR_NB + 1); e += 256) (&S.Tsm[0][0])[e] = cmake(0.0, 0.0);
  if (tid < QR_NB) S.udiag[tid] = cmake(0.0, 0.0);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&S.full[0])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&S.full[1])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (rank == 0) panel_body_paired<2, true, PROF>(S, cl, A, lda, Vout, ldv, m, np, r_begin, nloc, dabs_out, dabs_dup, Tout, prof, prof_rank);
  else panel_body_paired<T, false, PROF>(S, cl, A, lda, Vout, ldv, m, np, r_begin, nloc, dabs_out, dabs_dup, Tout, prof, prof_rank);
}

// =====================================================================================================
// Block-reflector application  C <- (I - V op(T) V^H) C  on 8-column blocks of C, streaming V and C
// from L2 in DMMA fragment order (no shared-memory staging of the operands).
// =====================================================================================================
// vmode 0: V is stored LAPACK style in the factored panel (unit diagonal implied, R above it); vmode 1: V is an explicit
// m x 32 block, zeros included (the paired panel kernel writes it that way: 2x2 blocks on its diagonal)
__device__ __forceinline__ cplx vmask_load(const cplx* __restrict__ V, int ldv, int m, int r, int c, int vmode) {
  if (r >= m) return cmake(0.0, 0.0);
  if (!vmode) {
    if (r < c) return cmake(0.0, 0.0);
    if (r == c) return cmake(1.0, 0.0);
  }
  return V[(size_t)c * ldv + r];
}

__global__ void __launch_bounds__(256)
larfb_kernel(const cplx* __restrict__ V, int ldv, int m, const cplx* __restrict__ T, int conjT,
             cplx* __restrict__ C, int ldc, int ncols, int vmode, cplx* __restrict__ Cb, int ldcb, int ncolsb) {
  __shared__ cplx Tsm[QR_NB][QR_NB + 1];
  __shared__ double Wp[4][4][32][4];
  __shared__ cplx W1[QR_NB][8], W2[QR_NB][8];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int lo = lane >> 2, lk = lane & 3;
  // two target matrices in one launch (the trailing columns and the right-hand side of a factorization: together they fill
  // the GPU, one after the other they take two latency-bound waves): column blocks [0, nblk_a) belong to C, the rest to Cb
  const int nblk_a = (ncols + 7) / 8;
  int c0 = blockIdx.x * 8;
  if ((int)blockIdx.x >= nblk_a) { C = Cb; ldc = ldcb; ncols = ncolsb; c0 = ((int)blockIdx.x - nblk_a) * 8; }

  for (int e = tid; e < QR_NB * QR_NB; e += blockDim.x) Tsm[e % QR_NB][e / QR_NB] = T[e];

  // ---- phase 1: W = V^H C   (32 x 8), K = m split over the 8 warps; the fragments of the next 8-row step are
  // loaded (L2 latency) while the DMMAs of the current one run
  double cr[4][2], ci[4][2];
#pragma unroll
  for (int it = 0; it < 4; ++it) cr[it][0] = cr[it][1] = ci[it][0] = ci[it][1] = 0.0;
  const int kchunk = (((m + 7) / 8) + 3) / 4 * 4;
  const int kbeg = warp * kchunk, kend = min(m, kbeg + kchunk);
  const bool colok = (c0 + lo) < ncols;
  auto load_step = [&](int r0, cplx (&av)[2][4], cplx (&bv)[2]) {
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int r = r0 + 4 * u + lk;
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int c = it * 8 + lo;
        av[u][it] = (r >= QR_NB && r < kend) ? V[(size_t)c * ldv + r] : (r < kend ? vmask_load(V, ldv, m, r, c, vmode) : cmake(0.0, 0.0));
      }
      bv[u] = (r < kend && colok) ? C[(size_t)(c0 + lo) * ldc + r] : cmake(0.0, 0.0);
    }
  };
  auto mma_step = [&](const cplx (&av)[2][4], const cplx (&bv)[2]) {
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
      for (int it = 0; it < 4; ++it) {   // A operand = conj(V)
        dmma884(cr[it][0], cr[it][1], av[u][it].x, bv[u].x);
        dmma884(cr[it][0], cr[it][1], av[u][it].y, bv[u].y);
        dmma884(ci[it][0], ci[it][1], av[u][it].x, bv[u].y);
        dmma884(ci[it][0], ci[it][1], -av[u][it].y, bv[u].x);
      }
  };
  {
    cplx avA[2][4], bvA[2], avB[2][4], bvB[2];
    int r0 = kbeg;
    if (r0 < kend) load_step(r0, avA, bvA);
    while (r0 < kend) {
      if (r0 + 8 < kend) load_step(r0 + 8, avB, bvB);
      mma_step(avA, bvA);
      r0 += 8;
      if (r0 >= kend) break;
      if (r0 + 8 < kend) load_step(r0 + 8, avA, bvA);
      mma_step(avB, bvB);
      r0 += 8;
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
  // ---- phase 3: C -= V W   (m x 8), K = 32; next tile's V fragments and C values prefetched
  const int ntile = (m + 7) / 8;
  auto load_tile = [&](int rt, cplx (&av)[8], cplx (&cv)[2]) {
    const int r = rt * 8 + lo;
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) {
      const int c = kk * 4 + lk;
      av[kk] = (rt * 8 >= QR_NB && r < m) ? V[(size_t)c * ldv + r] : vmask_load(V, ldv, m, r, c, vmode);
    }
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int col = c0 + 2 * lk + e;
      cv[e] = (r < m && col < ncols) ? C[(size_t)col * ldc + r] : cmake(0.0, 0.0);
    }
  };
  auto do_tile = [&](int rt, const cplx (&av)[8], const cplx (&cv)[2]) {
    const int r = rt * 8 + lo;
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
      if (r < m && col < ncols) C[(size_t)col * ldc + r] = cmake(cv[e].x - dr[e], cv[e].y - di[e]);
    }
  };
  {
    cplx avA[8], cvA[2], avB[8], cvB[2];
    int rt = warp;
    if (rt < ntile) load_tile(rt, avA, cvA);
    while (rt < ntile) {
      if (rt + 8 < ntile) load_tile(rt + 8, avB, cvB);
      do_tile(rt, avA, cvA);
      rt += 8;
      if (rt >= ntile) break;
      if (rt + 8 < ntile) load_tile(rt + 8, avA, cvA);
      do_tile(rt, avB, cvB);
      rt += 8;
    }
  }
}

// Same operation for NARROW C (few column blocks): an 8-CTA cluster per 8-column block splits the rows, so that the
// two latency-bound sweeps over V (L2 loads in fragment order) are 8x shorter; the 32 x 8 partial products V^H C are
// exchanged through distributed shared memory.  Used for the columns of the next panel, which sit on the critical path
// of the factorization, and for the tail of the trailing matrix where one CTA per block would leave most SMs idle.
#define LC_CL 8
__global__ void __cluster_dims__(LC_CL, 1, 1) __launch_bounds__(256)
larfb_cluster_kernel(const cplx* __restrict__ V, int ldv, int m, const cplx* __restrict__ T, int conjT,
                     cplx* __restrict__ C, int ldc, int ncols, int vmode) {
  cg::cluster_group cl = cg::this_cluster();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cplx* slots = reinterpret_cast<cplx*>(smem_raw);           // [LC_CL][256]: partial W of every CTA of the cluster
  __shared__ cplx Tsm[QR_NB][QR_NB + 1];
  __shared__ double Wp[4][4][32][4];
  __shared__ cplx W1[QR_NB][8], W2[QR_NB][8];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int lo = lane >> 2, lk = lane & 3;
  const int rank = (int)cl.block_rank();
  const int c0 = (blockIdx.x / LC_CL) * 8;
  const int rchunk = (((m + LC_CL - 1) / LC_CL) + 7) / 8 * 8;
  const int rbeg = min(m, rank * rchunk), rend = min(m, rbeg + rchunk);

  for (int e = tid; e < QR_NB * QR_NB; e += blockDim.x) Tsm[e % QR_NB][e / QR_NB] = T[e];
  // split cluster barrier: "I am running" now, waited for right before the first store into another CTA's shared memory (a remote
  // store is only defined once its target CTA has started; the wait is free, phase 1 lies in between)
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");

  // ---- phase 1: partial W = V^H C over my rows (32 x 8); 8-row steps dealt to the warps
  double cr[4][2], ci[4][2];
#pragma unroll
  for (int it = 0; it < 4; ++it) cr[it][0] = cr[it][1] = ci[it][0] = ci[it][1] = 0.0;
  const bool colok = (c0 + lo) < ncols;
  for (int r0 = rbeg + 8 * warp; r0 < rend; r0 += 64) {
    cplx av[2][4], bv[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int r = r0 + 4 * u + lk;
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int c = it * 8 + lo;
        av[u][it] = (r >= QR_NB && r < rend) ? V[(size_t)c * ldv + r] : (r < rend ? vmask_load(V, ldv, m, r, c, vmode) : cmake(0.0, 0.0));
      }
      bv[u] = (r < rend && colok) ? C[(size_t)(c0 + lo) * ldc + r] : cmake(0.0, 0.0);
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
    const int i = tid >> 3, c = tid & 7;            // element W[i][c] of my partial goes to slot `rank` of every CTA
    const int it = i >> 3, ln = (i & 7) * 4 + (c >> 1), e = c & 1;
    double sr = 0.0, si = 0.0;
#pragma unroll
    for (int w = 0; w < 4; ++w) { sr += Wp[w][it][ln][e]; si += Wp[w][it][ln][2 + e]; }
    const cplx v = cmake(sr, si);
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");   // every CTA of the cluster has started
#pragma unroll
    for (int dst = 0; dst < LC_CL; ++dst) {
      cplx* rs = cl.map_shared_rank(slots, dst);
      rs[rank * 256 + tid] = v;
    }
  }
  cl.sync();
  {
    cplx s0 = cadd(slots[tid], slots[256 + tid]), s1 = cadd(slots[2 * 256 + tid], slots[3 * 256 + tid]);
    cplx s2 = cadd(slots[4 * 256 + tid], slots[5 * 256 + tid]), s3 = cadd(slots[6 * 256 + tid], slots[7 * 256 + tid]);
    W1[tid >> 3][tid & 7] = cadd(cadd(s0, s1), cadd(s2, s3));
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
  // ---- phase 3: C -= V W on my rows (K = 32)
  for (int r0 = rbeg + 8 * warp; r0 < rend; r0 += 64) {
    const int r = r0 + lo;
    cplx av[8];
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) {
      const int c = kk * 4 + lk;
      av[kk] = (r0 >= QR_NB && r < m) ? V[(size_t)c * ldv + r] : vmask_load(V, ldv, m, r, c, vmode);
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
      if (r < rend && col < ncols) {
        cplx* p = C + (size_t)col * ldc + r;
        cplx t = *p;
        t.x -= dr[e]; t.y -= di[e];
        *p = t;
      }
    }
  }
  cl.sync();   // nobody exits while its shared memory may still be written... (all remote stores precede the first sync)
}

// The same for the 16 columns of the NEXT panel of the paired factorization -- the one block-reflector application that sits on
// the critical path of the panel chain (once per panel).  larfb_cluster_kernel spends its 11 us on L2 round trips in sequence
// (V fragments, then C, per 8-row step and per phase); here every global load of a CTA (<= 128 rows: two 8-row steps per warp
// in each phase) is issued before the first use, the T product runs on four independent accumulators, and the trailing cluster
// barrier is dropped (all remote stores precede the first one).  m <= 1024, vmode 1 (explicit V).
__global__ void __cluster_dims__(LC_CL, 1, 1) __launch_bounds__(256)
larfb_narrow_kernel(const cplx* __restrict__ V, int ldv, int m, const cplx* __restrict__ T, int conjT,
                    cplx* __restrict__ C, int ldc, int ncols) {
  cg::cluster_group cl = cg::this_cluster();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cplx* slots = reinterpret_cast<cplx*>(smem_raw);           // [LC_CL][256]: partial W of every CTA of the cluster
  __shared__ cplx Tsm[QR_NB][QR_NB + 1];
  __shared__ double Wp[4][4][32][4];
  __shared__ cplx W1[QR_NB][8], W2[QR_NB][8];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int lo = lane >> 2, lk = lane & 3;
  const int rank = (int)cl.block_rank();
  const int c0 = (blockIdx.x / LC_CL) * 8;
  const int rchunk = (((m + LC_CL - 1) / LC_CL) + 7) / 8 * 8;          // <= 128
  const int rbeg = min(m, rank * rchunk), rend = min(m, rbeg + rchunk);
  const cplx zero = cmake(0.0, 0.0);

  // ---- every global load of this CTA, up front: phase-1 fragments (conj(V)^T, C), phase-3 fragments (V, C tile)
  cplx a1[2][2][4], b1[2][2], a3[2][8], cv[2][2];
  const bool colok1 = (c0 + lo) < ncols;
#pragma unroll
  for (int s2 = 0; s2 < 2; ++s2) {
    const int r0 = rbeg + 8 * warp + 64 * s2;
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int r = r0 + 4 * u + lk;
      const bool ok = r < rend;
#pragma unroll
      for (int it = 0; it < 4; ++it) a1[s2][u][it] = ok ? V[(size_t)(it * 8 + lo) * ldv + r] : zero;
      b1[s2][u] = (ok && colok1) ? C[(size_t)(c0 + lo) * ldc + r] : zero;
    }
    const int r3 = r0 + lo;
    const bool ok3 = r3 < rend;
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) a3[s2][kk] = ok3 ? V[(size_t)(kk * 4 + lk) * ldv + r3] : zero;
#pragma unroll
    for (int e = 0; e < 2; ++e) cv[s2][e] = (ok3 && c0 + 2 * lk + e < ncols) ? C[(size_t)(c0 + 2 * lk + e) * ldc + r3] : zero;
  }
  for (int e = tid; e < QR_NB * QR_NB; e += blockDim.x) Tsm[e % QR_NB][e / QR_NB] = T[e];
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");   // "I am running" (waited for before the first remote store)

  // ---- phase 1: partial W = V^H C over my rows (32 x 8)
  double cr[4][2], ci[4][2];
#pragma unroll
  for (int it = 0; it < 4; ++it) cr[it][0] = cr[it][1] = ci[it][0] = ci[it][1] = 0.0;
#pragma unroll
  for (int s2 = 0; s2 < 2; ++s2)
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
      for (int it = 0; it < 4; ++it) {   // A operand = conj(V)
        dmma884(cr[it][0], cr[it][1], a1[s2][u][it].x, b1[s2][u].x);
        dmma884(cr[it][0], cr[it][1], a1[s2][u][it].y, b1[s2][u].y);
        dmma884(ci[it][0], ci[it][1], a1[s2][u][it].x, b1[s2][u].y);
        dmma884(ci[it][0], ci[it][1], -a1[s2][u][it].y, b1[s2][u].x);
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
    const int i = tid >> 3, c = tid & 7;            // element W[i][c] of my partial goes to slot `rank` of every CTA
    const int it = i >> 3, ln = (i & 7) * 4 + (c >> 1), e = c & 1;
    double sr = 0.0, si = 0.0;
#pragma unroll
    for (int w = 0; w < 4; ++w) { sr += Wp[w][it][ln][e]; si += Wp[w][it][ln][2 + e]; }
    const cplx v = cmake(sr, si);
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");   // every CTA of the cluster has started
#pragma unroll
    for (int dst = 0; dst < LC_CL; ++dst) {
      cplx* rs = cl.map_shared_rank(slots, dst);
      rs[rank * 256 + tid] = v;
    }
  }
  cl.sync();
  {
    cplx s0 = cadd(slots[tid], slots[256 + tid]), s1 = cadd(slots[2 * 256 + tid], slots[3 * 256 + tid]);
    cplx s2 = cadd(slots[4 * 256 + tid], slots[5 * 256 + tid]), s3 = cadd(slots[6 * 256 + tid], slots[7 * 256 + tid]);
    W1[tid >> 3][tid & 7] = cadd(cadd(s0, s1), cadd(s2, s3));
  }
  __syncthreads();
  // ---- phase 2: W <- op(T) W, four independent accumulator chains
  {
    const int i = tid >> 3, c = tid & 7;
    cplx acc[4] = {zero, zero, zero, zero};
    if (conjT) {
#pragma unroll
      for (int k = 0; k < QR_NB; ++k) cfma_conj(acc[k & 3], Tsm[k][i], W1[k][c]);
    } else {
#pragma unroll
      for (int k = 0; k < QR_NB; ++k) cfma(acc[k & 3], Tsm[i][k], W1[k][c]);
    }
    W2[i][c] = cadd(cadd(acc[0], acc[1]), cadd(acc[2], acc[3]));
  }
  __syncthreads();
  // ---- phase 3: C -= V W on my rows (K = 32), operands already in registers
#pragma unroll
  for (int s2 = 0; s2 < 2; ++s2) {
    const int r = rbeg + 8 * warp + 64 * s2 + lo;
    double dr[2] = {0.0, 0.0}, di[2] = {0.0, 0.0}, er[2] = {0.0, 0.0}, ei[2] = {0.0, 0.0};
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) {
      const cplx b = W2[kk * 4 + lk][lo];
      dmma884(dr[0], dr[1], a3[s2][kk].x, b.x);
      dmma884(er[0], er[1], -a3[s2][kk].y, b.y);
      dmma884(di[0], di[1], a3[s2][kk].x, b.y);
      dmma884(ei[0], ei[1], a3[s2][kk].y, b.x);
    }
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int col = c0 + 2 * lk + e;
      if (r < rend && col < ncols) C[(size_t)col * ldc + r] = cmake(cv[s2][e].x - (dr[e] + er[e]), cv[s2][e].y - (di[e] + ei[e]));
    }
  }
  // no trailing cluster barrier: every remote store into this CTA's shared memory precedes the cl.sync() above
}

__global__ void set_identity_kernel(cplx* Q, int ldq, int n) {
  size_t tot = (size_t)n * n;
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < tot; e += (size_t)gridDim.x * blockDim.x) {
    int r = (int)(e % n), c = (int)(e / n);
    Q[(size_t)c * ldq + r] = cmake(r == c ? 1.0 : 0.0, 0.0);
  }
}

long long* g_qr_prof = nullptr;
int g_qr_prof_rank = 1;        // cluster rank whose thread 0 is stamped by the paired panel kernel's profile instantiation
template <int NW, int T>
static int launch_panel_t(cudaStream_t st, cplx* A, int lda, int m, int nb, cplx* tau, double* dabs, cplx* Tf, size_t smem) {
  static SmemMemo memo, memo_prof;
  size_t smem_lim = 0;
  if (ensure_max_dynamic_smem(qr_panel_kernel<NW, T, false>, memo, &smem_lim) || ensure_max_dynamic_smem(qr_panel_kernel<NW, T, true>, memo_prof, &smem_lim)) return -1;
  if (smem > smem_lim) { snprintf(g_errbuf, sizeof(g_errbuf), "qr panel: m=%d too large", m); return -1; }
  if (g_qr_prof) qr_panel_kernel<NW, T, true><<<QR_CL, NW * 32, smem, st>>>(A, lda, m, nb, tau, dabs, Tf, g_qr_prof);
  else qr_panel_kernel<NW, T, false><<<QR_CL, NW * 32, smem, st>>>(A, lda, m, nb, tau, dabs, Tf, nullptr);
  CUDA_TRY(cudaGetLastError());
  g_launches++;
  return 0;
}
static int launch_panel(cudaStream_t st, cplx* A, int lda, int m, int nb, cplx* tau, double* dabs, cplx* T) {
  const int rs = panel_rows_below(m);   // rows per CTA below the diagonal block
  const size_t smem = sizeof(cplx) * ((size_t)QR_LDA * max(rs, QR_NB) + 8);
  if (rs <= 8) return launch_panel_t<8, 1>(st, A, lda, m, nb, tau, dabs, T, smem);
  if (rs <= 16) return launch_panel_t<8, 2>(st, A, lda, m, nb, tau, dabs, T, smem);
  if (rs <= 32) return launch_panel_t<8, 4>(st, A, lda, m, nb, tau, dabs, T, smem);
  if (rs <= 64) return launch_panel_t<8, 8>(st, A, lda, m, nb, tau, dabs, T, smem);
  if (rs <= 96) return launch_panel_t<8, 12>(st, A, lda, m, nb, tau, dabs, T, smem);
  if (rs <= 128) return launch_panel_t<8, 16>(st, A, lda, m, nb, tau, dabs, T, smem);
  if (rs <= 144) return launch_panel_t<8, 18>(st, A, lda, m, nb, tau, dabs, T, smem);
  if (rs <= 192) return launch_panel_t<16, 12>(st, A, lda, m, nb, tau, dabs, T, smem);
  if (rs <= 224) return launch_panel_t<16, 14>(st, A, lda, m, nb, tau, dabs, T, smem);
  if (rs <= 256) return launch_panel_t<16, 16>(st, A, lda, m, nb, tau, dabs, T, smem);
  if (rs <= 320) return launch_panel_t<16, 20>(st, A, lda, m, nb, tau, dabs, T, smem);
  snprintf(g_errbuf, sizeof(g_errbuf), "qr panel: m=%d too large", m);
  return -1;
}

int g_larfb_cluster_max_cols = 256;   // column count up to which the row-split cluster kernel is used
static int launch_larfb(cudaStream_t st, const cplx* V, int ldv, int m, const cplx* T, int conjT, cplx* C, int ldc,
                        int ncols, int vmode = 0) {
  if (ncols <= 0) return 0;
  static const bool narrow_ok = getenv("DQMC_LARFB_NARROW") == nullptr || atoi(getenv("DQMC_LARFB_NARROW")) != 0;
  if (narrow_ok && vmode == 1 && ncols <= 16 && m >= 64 && m <= 1024) {
    static SmemMemo memo_n;
    const size_t smem = sizeof(cplx) * LC_CL * 256;
    if (ensure_dynamic_smem(larfb_narrow_kernel, memo_n, smem)) return -1;
    larfb_narrow_kernel<<<((ncols + 7) / 8) * LC_CL, 256, smem, st>>>(V, ldv, m, T, conjT, C, ldc, ncols);
    CUDA_TRY(cudaGetLastError());
    g_launches++;
    return 0;
  }
  if (ncols <= g_larfb_cluster_max_cols && m >= 64) {
    static SmemMemo memo;
    const size_t smem = sizeof(cplx) * LC_CL * 256;
    if (ensure_dynamic_smem(larfb_cluster_kernel, memo, smem)) return -1;
    larfb_cluster_kernel<<<((ncols + 7) / 8) * LC_CL, 256, smem, st>>>(V, ldv, m, T, conjT, C, ldc, ncols, vmode);
    CUDA_TRY(cudaGetLastError());
    g_launches++;
    return 0;
  }
  larfb_kernel<<<(ncols + 7) / 8, 256, 0, st>>>(V, ldv, m, T, conjT, C, ldc, ncols, vmode, nullptr, 0, 0);
  CUDA_TRY(cudaGetLastError());
  g_launches++;
  return 0;
}
// the same block reflector applied to two matrices (either may be empty) in one launch of the wide kernel
static int launch_larfb2(cudaStream_t st, const cplx* V, int ldv, int m, const cplx* T, int conjT, cplx* Ca, int ldca, int ncolsa,
                         cplx* Cb, int ldcb, int ncolsb, int vmode) {
  if (ncolsa <= 0) return launch_larfb(st, V, ldv, m, T, conjT, Cb, ldcb, ncolsb, vmode);
  if (ncolsb <= 0) return launch_larfb(st, V, ldv, m, T, conjT, Ca, ldca, ncolsa, vmode);
  if (ncolsa + ncolsb <= g_larfb_cluster_max_cols && m >= 64) {   // few columns: two launches of the row-split cluster kernel
    if (launch_larfb(st, V, ldv, m, T, conjT, Ca, ldca, ncolsa, vmode)) return -1;
    return launch_larfb(st, V, ldv, m, T, conjT, Cb, ldcb, ncolsb, vmode);
  }
  larfb_kernel<<<(ncolsa + 7) / 8 + (ncolsb + 7) / 8, 256, 0, st>>>(V, ldv, m, T, conjT, Ca, ldca, ncolsa, vmode, Cb, ldcb, ncolsb);
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

int qr_panels_only(cudaStream_t st, cplx* A, int lda, int n, cplx* tau, double* dabs, cplx* tfac) {
  for (int j0 = 0, k = 0; j0 < n; j0 += QR_NB, ++k)
    if (launch_panel(st, A + (size_t)j0 * lda + j0, lda, n - j0, min(QR_NB, n - j0), tau + j0, dabs + j0, tfac + (size_t)k * QR_NB * QR_NB))
      return -1;
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

// ---- paired factorization: driver.  AL: n x n/2 (left half, pair-interleaved rows), factored in place into the quaternion
// upper triangle R_L (exact zeros below); V: n x n scratch receiving the explicit reflector blocks (panel k at (32k, 32k));
// dabs[0:n/2] (and a copy in dabs[n/2:n]) = quaternion modulus of the diagonal of R; rhs (n x nrhs, pair-interleaved rows) <- Q^H rhs.
template <int T>
static int launch_panel_paired_t(cudaStream_t st, cplx* A, int lda, cplx* V, int ldv, int m, int np, double* dabs, int dup,
                                 cplx* Tf, size_t smem) {
  (void)smem;   // the slab goes from global memory straight into registers and back: no staging area
  if (g_qr_prof) qr_panel_paired_kernel<T, true><<<QR_CL, 256, 0, st>>>(A, lda, V, ldv, m, np, dabs, dup, Tf, g_qr_prof, g_qr_prof_rank);
  else qr_panel_paired_kernel<T, false><<<QR_CL, 256, 0, st>>>(A, lda, V, ldv, m, np, dabs, dup, Tf, nullptr, 0);
  CUDA_TRY(cudaGetLastError());
  g_launches++;
  return 0;
}
static int launch_panel_paired(cudaStream_t st, cplx* A, int lda, cplx* V, int ldv, int m, int np, double* dabs, int dup, cplx* T) {
  const int rs = paired_rows_below(m);            // interleaved rows per CTA below the diagonal block
  const size_t smem = sizeof(cplx) * ((size_t)QR_LDA * max(rs, QR_NB) + 8);
  const int tq = (rs / 2 + 7) / 8;                // quaternion rows per thread
#define PP(TT) return launch_panel_paired_t<TT>(st, A, lda, V, ldv, m, np, dabs, dup, T, smem)
  if (tq <= 1) PP(1);
  if (tq <= 2) PP(2);
  if (tq <= 3) PP(3);
  if (tq <= 4) PP(4);
  if (tq <= 5) PP(5);
  if (tq <= 6) PP(6);
  if (tq <= 7) PP(7);
  if (tq <= 8) PP(8);
  if (tq <= 9) PP(9);
  if (tq <= 10) PP(10);
  if (tq <= 12) PP(12);
  if (tq <= 14) PP(14);
  if (tq <= 16) PP(16);
  if (tq <= 20) PP(20);
#undef PP
  snprintf(g_errbuf, sizeof(g_errbuf), "paired qr panel: m=%d too large", m);
  return -1;
}

int qr_factor_paired(cudaStream_t st, cplx* AL, int lda, int n, cplx* V, int ldv, double* dabs, cplx* tfac, cplx* rhs, int ldr,
                     int nrhs, const QrAsync* as) {
  const int h = n / 2;
  if (n % 2 != 0 || h % QP_NP != 0) { snprintf(g_errbuf, sizeof(g_errbuf), "paired qr: n=%d is not a multiple of 32", n); return -1; }
  const bool la = as != nullptr && h > 2 * QP_NP;
  for (int j0 = 0, k = 0; j0 < h; j0 += QP_NP, ++k) {
    const int np = min(QP_NP, h - j0), r0 = 2 * j0, m = n - r0;
    cplx* P = AL + (size_t)j0 * lda + r0;
    cplx* Vp = V + (size_t)r0 * ldv + r0;
    cplx* T = tfac + (size_t)k * QR_NB * QR_NB;
    if (launch_panel_paired(st, P, lda, Vp, ldv, m, np, dabs + j0, h, T)) return -1;
    const int ntrail = h - j0 - np;
    if (!la) {
      if (launch_larfb(st, Vp, ldv, m, T, 1, P + (size_t)np * lda, lda, ntrail, 1)) return -1;
      if (rhs && launch_larfb(st, Vp, ldv, m, T, 1, rhs + r0, ldr, nrhs, 1)) return -1;
      continue;
    }
    const int nnext = min(QP_NP, ntrail);
    // the wide update of panel k needs V_k, T_k only and touches none of the next panel's columns: it may start as soon as the
    // panel is factored, next to the narrow update below (which must see the wide update of panel k-1)
    CUDA_TRY(cudaEventRecord(as->eA, st));
    if (k > 0) CUDA_TRY(cudaStreamWaitEvent(st, as->eB, 0));
    if (launch_larfb(st, Vp, ldv, m, T, 1, P + (size_t)np * lda, lda, nnext, 1)) return -1;
    CUDA_TRY(cudaStreamWaitEvent(as->st2, as->eA, 0));
    if (launch_larfb2(as->st2, Vp, ldv, m, T, 1, P + (size_t)(np + nnext) * lda, lda, ntrail - nnext, rhs ? rhs + r0 : nullptr, ldr,
                      rhs ? nrhs : 0, 1))
      return -1;
    CUDA_TRY(cudaEventRecord(as->eB, as->st2));
  }
  if (la) CUDA_TRY(cudaStreamWaitEvent(st, as->eB, 0));
  return 0;
}

int qr_panels_only_paired(cudaStream_t st, cplx* AL, int lda, int n, cplx* V, int ldv, double* dabs, cplx* tfac) {
  const int h = n / 2;
  for (int j0 = 0, k = 0; j0 < h; j0 += QP_NP, ++k)
    if (launch_panel_paired(st, AL + (size_t)j0 * lda + 2 * j0, lda, V + (size_t)(2 * j0) * ldv + 2 * j0, ldv, n - 2 * j0, min(QP_NP, h - j0),
                            dabs + j0, h, tfac + (size_t)k * QR_NB * QR_NB))
      return -1;
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

// Same for a QUATERNION upper triangle (the R of the paired factorization with rows and columns pair-interleaved): upper
// triangular in 2x2 blocks, diagonal blocks [[e, -conj(o)], [o, conj(e)]] (inverse = conjugate transpose / (|e|^2 + |o|^2)).
__global__ void __launch_bounds__(32) trtri_diag_quat_kernel(const cplx* __restrict__ A, int lda, int n, cplx* __restrict__ inv) {
  __shared__ cplx R[QR_NB][QR_NB + 1];
  const int kb = blockIdx.x, j0 = kb * QR_NB, nb = min(QR_NB, n - j0), j = threadIdx.x;
  for (int c = 0; c < QR_NB; ++c) {
    cplx v = cmake(0.0, 0.0);
    if (j < nb && c < nb && (j | 1) <= (c | 1)) v = A[(size_t)(j0 + c) * lda + j0 + j];
    if (j == c && j >= nb) v = cmake(1.0, 0.0);
    R[j][c] = v;
  }
  __syncwarp();
  cplx x[QR_NB];
#pragma unroll
  for (int i = 0; i < QR_NB; ++i) x[i] = cmake(0.0, 0.0);
  // column j of the inverse: block back substitution of R x = e_j, one row pair per step (fully unrolled: x in registers)
#pragma unroll
  for (int p = QR_NB / 2 - 1; p >= 0; --p) {
    const int i = 2 * p;
    if (i <= j) {
      cplx a0 = cmake(i == j ? 1.0 : 0.0, 0.0), a1 = cmake(i + 1 == j ? 1.0 : 0.0, 0.0);
#pragma unroll
      for (int k = i + 2; k < QR_NB; ++k)
        if (k <= (j | 1)) { a0 = csub(a0, cmul(R[i][k], x[k])); a1 = csub(a1, cmul(R[i + 1][k], x[k])); }
      // diagonal block [[r00, r01], [r10, r11]]: general 2x2 inverse (robust to the block being only approximately a quaternion)
      const cplx r00 = R[i][i], r01 = R[i][i + 1], r10 = R[i + 1][i], r11 = R[i + 1][i + 1];
      const cplx det = csub(cmul(r00, r11), cmul(r01, r10));
      x[i] = cdiv(csub(cmul(r11, a0), cmul(r01, a1)), det);
      x[i + 1] = cdiv(csub(cmul(r00, a1), cmul(r10, a0)), det);
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

__global__ void row_scale_kernel(cplx* __restrict__ Y, int ldy, int n, int nrhs, const double* __restrict__ rowscale) {
  const size_t total = (size_t)n * nrhs;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(e % n), c = (int)(e / n);
    cplx* p = Y + (size_t)c * ldy + r;
    *p = cscale(*p, rowscale[r]);
  }
}

static int trsm_block(cudaStream_t st, const cplx* A, int lda, int n, cplx* Y, int ldy, int nrhs, const cplx* inv,
                      const double* rowscale) {
  const int n8 = (n + 7) / 8 * 8, lds = n8 + 4;
  const size_t smem = sizeof(cplx) * ((size_t)8 * lds + QR_NB * 8 + QR_NB * QR_NB);
  static SmemMemo memo;
  size_t smem_lim = 0;
  if (ensure_max_dynamic_smem(trsm_kernel, memo, &smem_lim)) return -1;
  if (smem > smem_lim) { snprintf(g_errbuf, sizeof(g_errbuf), "trsm: n=%d too large", n); return -1; }
  trsm_kernel<<<(nrhs + 7) / 8, 256, smem, st>>>(A, lda, n, Y, ldy, nrhs, inv, rowscale);
  CUDA_TRY(cudaGetLastError());
  g_launches++;
  return 0;
}

// Y <- R^-1 Y (optionally with the rows of the result scaled).  Above TRSM_BS rows the back substitution is blocked at the
// host level: the per-CTA kernel (which streams its whole triangle from L2 for every 8 right-hand sides) solves TRSM_BS-row
// diagonal blocks, the rectangular part of R goes through ZGEMM.
#define TRSM_BS 512
int trsm_upper(cudaStream_t st, const cplx* A, int lda, int n, cplx* Y, int ldy, int nrhs, cplx* work,
               const double* rowscale, int num_sms, int quat) {
  const int nblk = (n + QR_NB - 1) / QR_NB;
  if (quat) trtri_diag_quat_kernel<<<nblk, 32, 0, st>>>(A, lda, n, work);
  else trtri_diag_kernel<<<nblk, 32, 0, st>>>(A, lda, n, work);
  CUDA_TRY(cudaGetLastError());
  g_launches++;
  static const bool split = getenv("DQMC_TRSM_NOSPLIT") == nullptr;
  if (n <= TRSM_BS || !split) return trsm_block(st, A, lda, n, Y, ldy, nrhs, work, rowscale);
  const cplx one = cmake(1.0, 0.0), mone = cmake(-1.0, 0.0);
  for (int j0 = (n - 1) / TRSM_BS * TRSM_BS; j0 >= 0; j0 -= TRSM_BS) {
    const int len = min(TRSM_BS, n - j0);
    if (trsm_block(st, A + (size_t)j0 * lda + j0, lda, len, Y + j0, ldy, nrhs, work + (size_t)(j0 / QR_NB) * QR_NB * QR_NB, nullptr))
      return -1;
    if (j0 > 0 && zgemm(st, OP_N, OP_N, j0, nrhs, len, mone, A + (size_t)j0 * lda, lda, Y + j0, ldy, one, Y, ldy, num_sms)) return -1;
  }
  if (rowscale) {
    row_scale_kernel<<<num_sms * 4, 256, 0, st>>>(Y, ldy, n, nrhs, rowscale);
    CUDA_TRY(cudaGetLastError());
    g_launches++;
  }
  return 0;
}
