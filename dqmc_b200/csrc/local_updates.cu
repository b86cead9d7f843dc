#include "local_updates.cuh"

// =====================================================================================================
// Persistent local-update kernel (one launch per time slice).
//
// All CTAs run the N-site Metropolis loop in lock step and take every accept/reject decision redundantly
// from identical data, so a rejected proposal costs no communication.  CTA b owns rows [b*rpc, (b+1)*rpc) of
// the pending update factor A (n x 4k, stored transposed, "At") and the same columns of B (4k x n, "Bm").
// An accepted proposal makes every CTA append its slice of the new 4 columns/rows; consumers spin on the
// data itself (the buffers are pre-filled with a NaN sentinel), so there is no barrier per accept.  After
// kmax accepts (or at the end of the slice) G += A*B is flushed with DMMA tiles between two grid barriers.
//
// Inside a CTA the work of site i+1 that does not depend on the decision at site i is pipelined behind it:
//   warp 0      decision for site i  (M = 1 + Delta (1 - G_eff), det, accept/reject, M^-1)
//   warps 1-3   proposal for site i+1 under the three possible outcomes of site i
//               (rejected / accepted without a draw / accepted with a draw): new field value, dS, exp(-dS), Delta
//   warps 4-7   prefetch of the rows/columns of A, B and G that site i+1 needs, partial G_eff
// =====================================================================================================

#include "lu_common.cuh"

// PROF = true adds the cycle counters of tools/lu_profile.py; the production instantiation carries none of that code (the site
// loop's instruction stream is as large as the instruction cache, every instruction less counts).
template <bool PROF>
__global__ void __launch_bounds__(256) local_updates_kernel(LUArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int n = a.n, N = a.nsites, ldk = 4 * a.kmax, rpc = a.rpc;
  const int lds = ldk + 2;                                  // shared-memory row stride: +2 spreads the 4 rows of a site over the banks
  const size_t bufstride = (size_t)ldk * n;                 // one At / Bm buffer
  cplx* sp = reinterpret_cast<cplx*>(smem_raw);
  cplx* Aown = sp; sp += (size_t)rpc * lds;                 // [rpc][ldk] my rows of A
  cplx* Bown = sp; sp += (size_t)rpc * lds;                 // [rpc][ldk] my columns of B
  cplx* As4 = sp; sp += 2 * 4 * lds;                        // [2][4][ldk] rows site+kN of A (double-buffered by site parity)
  cplx* Bs4 = sp; sp += 2 * 4 * lds;                        // [2][4][ldk] cols site+kN of B
  cplx* gcol = sp; sp += (size_t)2 * rpc * 4;               // [2][rpc][4] G[r, site+kN] for my rows
  cplx* grow = sp; sp += (size_t)2 * rpc * 4;               // [2][rpc][4] G[site+kN, c] for my cols
  cplx* FA = sp; sp += 2 * 64 * 36;                         // flush staging: 2 k-chunks x 64 rows x 32 k (+4 pad)
  cplx* FB = sp; sp += 2 * 64 * 36;
  double* fs = reinterpret_cast<double*>(sp);               // [3N] field of this slice
  double* tn = fs + 3 * N;                                  // [3N] phi(l+1) + phi(l-1)
  double* uw = tn + 3 * N;                                  // [4N] this slice's window of the uniform stream
  int* nbr = reinterpret_cast<int*>(uw + 4 * N);            // [4N] spatial neighbours
  __shared__ cplx g4e[2][16], g4r[16], Gx1[16], Gx2[16], X1s[16], X2s[16], T1s[16], T2s[16];
  __shared__ cplx Mm[16], Cof[16], Minv[16], gcc[64 * 4], grc[64 * 4];
  __shared__ Prep prep[2][3];
  __shared__ int s_accept, s_scn, s_exh;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int row0 = blockIdx.x * rpc;
  const int nown = max(0, min(n, row0 + rpc) - row0);
  const int sl = a.slice;
  const int sl_later = (sl + 1) % a.nslices, sl_earlier = (sl + a.nslices - 1) % a.nslices;
  const long long pos0 = *a.pos;
  const int navail = (int)max(0LL, min((long long)4 * N, a.nunif - pos0));

  for (int e = tid; e < 3 * N; e += blockDim.x) {
    fs[e] = a.hs[(size_t)3 * N * sl + e];
    tn[e] = a.hs[(size_t)3 * N * sl_later + e] + a.hs[(size_t)3 * N * sl_earlier + e];
  }
  for (int e = tid; e < 4 * N; e += blockDim.x) {
    nbr[e] = a.nbr[e];
    uw[e] = e < navail ? a.unif[pos0 + e] : 0.0;
  }
  if (tid == 0) s_exh = 0;
  unsigned int* const bar_ctr = a.bar + 2 + a.bar_parity;   // monotonic barrier counter of this launch (zero at launch)
  unsigned int bar_target = 0;
  if (a.bar_mode && blockIdx.x == 0 && tid == 0) a.bar[2 + (1 - a.bar_parity)] = 0;   // the next launch's counter
  int off = 0;                                              // stream position relative to pos0
  long long nacc = 0;
  double dS_sum = 0.0;
  int kc = 0, np = 0, batch = 0, nonreal = 0, s_cur = 0;
  // optional cycle profile of CTA 0: arrival stamps before the two CTA barriers of an iteration (a clock read right
  // after bar.sync would capture the barrier's issue, not its release)
  const bool prof = PROF && (a.prof != nullptr) && blockIdx.x == 0;
  __shared__ long long stampA[8], stampB[8], stampP[8];
  __shared__ long long p_role[8], p_role1[8];
  if (tid < 8) { p_role[tid] = 0; p_role1[tid] = 0; }
  long long pf[3] = {0, 0, 0};
  long long p_s1 = 0, p_rest_acc = 0, p_rest_rej = 0, p_flush = 0, n_fl = 0, rel_prev = 0;
  __syncthreads();
  // every CTA has read the field, the neighbour sums and the stream position: only now may CTA 0 write accepted field values
  // back into hs (and *pos at the end).  Co-residency does not mean simultaneous start.
  if (a.bar_mode) { bar_target += gridDim.x; grid_barrier_mono(bar_ctr, bar_target); } else grid_barrier(a.bar, gridDim.x);
  const long long t_begin = clock64();
  rel_prev = t_begin;

  // G-derived data of `site` into buffer bb (tp = index among the nth participating threads)
  auto fetch_G = [&](int site, int bb, int tp, int nth) {
    if (tp < 16) g4r[tp] = ldcg2(a.G + (size_t)(site + (tp >> 2) * N) * n + site + (tp & 3) * N);   // [r + 4c]
    for (int e = tp; e < nown * 4; e += nth) {
      const int rl = e >> 2, k = e & 3;
      gcol[bb * rpc * 4 + e] = ldcg2(a.G + (size_t)(site + k * N) * n + row0 + rl);
      grow[bb * rpc * 4 + e] = ldcg2(a.G + (size_t)(row0 + rl) * n + site + k * N);
    }
  };

  // ---- prologue: site 0
  fetch_G(0, 0, tid, 256);
  if (warp == 1) do_prep(a, fs, tn, nbr, uw, 0, navail, 0, -1, 0.0, 0.0, 0.0, &prep[0][0], &s_exh);
  __syncthreads();
  if (tid < 16) g4e[0][tid] = g4r[tid];
  __syncthreads();

  for (int i = 0; i < N; ++i) {
    const int b = i & 1, nb = b ^ 1;
    const cplx* Atb = a.At + batch * bufstride;
    const cplx* Bmb = a.Bm + batch * bufstride;
    const bool have_next = (i + 1 < N);
    // ================= stage 1: three roles in parallel =================
    if (warp == 0) {
      const Prep& P = prep[b][s_cur];
      const int r = (lane >> 2) & 3, c = lane & 3;
      if (lane < 16) {   // M = 1 + Delta * (1 - G_eff)
        cplx g0 = g4e[b][0 + 4 * c], g1 = g4e[b][1 + 4 * c], g2 = g4e[b][2 + 4 * c], g3 = g4e[b][3 + 4 * c];
        g0 = cmake((c == 0 ? 1.0 : 0.0) - g0.x, -g0.y);
        g1 = cmake((c == 1 ? 1.0 : 0.0) - g1.x, -g1.y);
        g2 = cmake((c == 2 ? 1.0 : 0.0) - g2.x, -g2.y);
        g3 = cmake((c == 3 ? 1.0 : 0.0) - g3.x, -g3.y);
        const cplx t0 = cmul(P.D[r * 4 + 0], g0), t1 = cmul(P.D[r * 4 + 1], g1);
        const cplx t2 = cmul(P.D[r * 4 + 2], g2), t3 = cmul(P.D[r * 4 + 3], g3);
        cplx m = cadd(cadd(t0, t1), cadd(t2, t3));
        if (r == c) m.x += 1.0;
        Mm[r * 4 + c] = m;
      }
      __syncwarp();
      if (lane < 16) {   // cofactor (r,c)
        const int r0 = (r == 0) ? 1 : 0, r1 = (r <= 1) ? 2 : 1, r2 = (r <= 2) ? 3 : 2;
        const int c0 = (c == 0) ? 1 : 0, c1 = (c <= 1) ? 2 : 1, c2 = (c <= 2) ? 3 : 2;
        cplx d = det3(Mm[r0 * 4 + c0], Mm[r0 * 4 + c1], Mm[r0 * 4 + c2], Mm[r1 * 4 + c0], Mm[r1 * 4 + c1], Mm[r1 * 4 + c2],
                      Mm[r2 * 4 + c0], Mm[r2 * 4 + c1], Mm[r2 * 4 + c2]);
        Cof[r * 4 + c] = ((r + c) & 1) ? cneg(d) : d;
      }
      __syncwarp();
      const cplx p0 = cmul(Mm[0], Cof[0]), p1 = cmul(Mm[1], Cof[1]), p2 = cmul(Mm[2], Cof[2]), p3 = cmul(Mm[3], Cof[3]);
      const cplx det = cadd(cadd(p0, p1), cadd(p2, p3));     // expansion along row 0
      const double p_acc = P.e_dS * det.x;
      int acc_flag, scn;
      if (p_acc > 1.0) { acc_flag = 1; scn = 1; }
      else { acc_flag = (P.u3 < p_acc) ? 1 : 0; scn = acc_flag ? 2 : 0; }
      if (lane == 0) { s_accept = acc_flag; s_scn = scn; }
      if (acc_flag) {
        if (lane < 16) {
          const double id = 1.0 / (det.x * det.x + det.y * det.y);
          const cplx dinv = cmake(det.x * id, -det.y * id);
          Minv[r * 4 + c] = cmul(Cof[c * 4 + r], dinv);
        }
        nacc++;
        dS_sum += P.mlog;
      }
      if (fabs(det.y) > 1e-4 * fabs(det.x)) nonreal++;
    } else if (warp <= 3) {
      if (have_next) {
        const int sc = warp - 1;                                  // 0 rejected, 1 accepted (no draw), 2 accepted (draw)
        const Prep& Pc = prep[b][s_cur];
        do_prep(a, fs, tn, nbr, uw, off + (sc == 1 ? 3 : 4), navail, i + 1, sc == 0 ? -1 : i, Pc.nw[0], Pc.nw[1], Pc.nw[2],
                &prep[nb][sc], &s_exh);
      }
    } else {
      const int tp = tid - 128;
      // ---- loads: everything is issued before anything is waited for (volatile asm keeps the issue order; the values
      // go to shared memory only after the A/B loads below are in flight)
      // (G) plain L2 loads: 4x4 block of site i+1, the two cross blocks between sites i+1 and i, my rows/cols of G
      cplx gq = cmake(0.0, 0.0), gcv = gq, grv = gq;
      if (have_next) {
        const int site = i + 1;
        const cplx* pq = nullptr;
        if (tp < 16) pq = a.G + (size_t)(site + (tp >> 2) * N) * n + site + (tp & 3) * N;                                   // [r + 4c]
        else if (tp < 32) { const int q = tp - 16; pq = a.G + (size_t)(i + (q & 3) * N) * n + site + (q >> 2) * N; }        // [r*4+k] = G[i+1+rN, i+kN]
        else if (tp < 48) { const int q = tp - 32; pq = a.G + (size_t)(site + (q & 3) * N) * n + i + (q >> 2) * N; }        // [k*4+c] = G[i+kN, i+1+cN]
        if (pq) gq = ld_cg_issue(pq);
        if (tp < nown * 4) {
          const int rl = tp >> 2, k = tp & 3;
          gcv = ld_cg_issue(a.G + (size_t)(site + k * N) * n + row0 + rl);
          grv = ld_cg_issue(a.G + (size_t)(row0 + rl) * n + site + k * N);
        }
      }
      // (A/B) all np pending columns for the rows/cols of site i+1, published by their owner CTAs (spin on the NaN
      // sentinel).  The current site's newest columns never come from memory: stage 2 derives them locally.
      // thread tp owns one (factor w, row/col k) pair and the pending columns p = tp/8 + 16u
      {
        const int w8 = (tp >> 2) & 1, k8 = tp & 3, p8 = tp >> 3;
        const cplx* sbase = (w8 ? Bmb : Atb) + (size_t)(i + 1 + k8 * N) * ldk;
        cplx* dbase = (w8 ? Bs4 : As4) + (nb * 4 + k8) * lds;
        unsigned long long vx[4], vy[4];   // np <= 64 (kmax <= 16): four window columns per thread
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (have_next && p8 + 16 * u < np)
            asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(vx[u]), "=l"(vy[u]) : "l"(sbase + p8 + 16 * u) : "memory");
        if (have_next) {   // G values: into shared memory while the A/B loads are in flight
          if (tp < 16) g4r[tp] = gq; else if (tp < 32) Gx1[tp - 16] = gq; else if (tp < 48) Gx2[tp - 32] = gq;
          if (tp < nown * 4) { gcol[nb * rpc * 4 + tp] = gcv; grow[nb * rpc * 4 + tp] = grv; }
          for (int e = tp + 128; e < nown * 4; e += 128) {
            const int rl = e >> 2, k = e & 3;
            gcol[nb * rpc * 4 + e] = ldcg2(a.G + (size_t)(i + 1 + k * N) * n + row0 + rl);
            grow[nb * rpc * 4 + e] = ldcg2(a.G + (size_t)(row0 + rl) * n + i + 1 + k * N);
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (have_next && p8 + 16 * u < np) {
            while (vx[u] == LU_SENT || vy[u] == LU_SENT)
              asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(vx[u]), "=l"(vy[u]) : "l"(sbase + p8 + 16 * u) : "memory");
            dbase[p8 + 16 * u] = make_double2(__longlong_as_double((long long)vx[u]), __longlong_as_double((long long)vy[u]));
          }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      // ---- partial sums over the pending columns: G_eff block of site i+1 and the two cross blocks (48 outputs x 2 threads)
      if (have_next && tp < 96) {
        const int o = tp >> 1, h = tp & 1, which = o >> 4, idx = o & 15;
        const cplx *pa, *pb;
        cplx base;
        if (which == 0) { pa = As4 + (nb * 4 + (idx & 3)) * lds; pb = Bs4 + (nb * 4 + (idx >> 2)) * lds; base = g4r[idx]; }
        else if (which == 1) { pa = As4 + (nb * 4 + (idx >> 2)) * lds; pb = Bs4 + (b * 4 + (idx & 3)) * lds; base = Gx1[idx]; }
        else { pa = As4 + (b * 4 + (idx >> 2)) * lds; pb = Bs4 + (nb * 4 + (idx & 3)) * lds; base = Gx2[idx]; }
        cplx acc0 = cmake(0.0, 0.0), acc1 = acc0;
        int p = h;
#pragma unroll 1
        for (; p + 2 < np; p += 4) { cfma(acc0, pa[p], pb[p]); cfma(acc1, pa[p + 2], pb[p + 2]); }
        if (p < np) cfma(acc0, pa[p], pb[p]);
        cplx acc = cadd(acc0, acc1);
        acc.x += __shfl_xor_sync(0xffffffffu, acc.x, 1);
        acc.y += __shfl_xor_sync(0xffffffffu, acc.y, 1);
        if (h == 0) {
          acc = cadd(acc, base);
          if (which == 0) g4e[nb][idx] = acc; else if (which == 1) X1s[idx] = acc; else X2s[idx] = acc;
        }
      }
    }
    if (prof && lane == 0) stampP[warp] = clock64();
    if (warp < 4) {
      // speculative part of a possible accept (does not depend on the decision): G_eff[r, i+kN] - delta for my rows and
      // G_eff[i+kN, c] for my columns, 2 threads per dot product over the pending columns
      const int half = nown * 4;
      for (int t0 = 0; t0 < 2 * half; t0 += 64) {
        const int t = t0 + (tid >> 1), h = tid & 1;
        const bool act = t < 2 * half;
        const bool isB = t >= half;
        const int tt = isB ? t - half : t;
        const int rl = tt >> 2, kq = tt & 3;
        cplx acc0 = cmake(0.0, 0.0), acc1 = acc0;
        if (act) {
          const cplx* own = (isB ? Bown : Aown) + (size_t)rl * lds;
          const cplx* site = (isB ? As4 : Bs4) + (b * 4 + kq) * lds;
          int p = h;
#pragma unroll 1
          for (; p + 2 < np; p += 4) { cfma(acc0, own[p], site[p]); cfma(acc1, own[p + 2], site[p + 2]); }
          if (p < np) cfma(acc0, own[p], site[p]);
        }
        cplx acc = cadd(acc0, acc1);
        acc.x += __shfl_xor_sync(0xffffffffu, acc.x, 1);
        acc.y += __shfl_xor_sync(0xffffffffu, acc.y, 1);
        if (act && h == 0) {
          if (!isB) {
            cplx v = cadd(gcol[b * rpc * 4 + tt], acc);
            if (row0 + rl == i + kq * N) v.x -= 1.0;
            gcc[tt] = v;
          } else {
            grc[tt] = cadd(grow[b * rpc * 4 + tt], acc);
          }
        }
      }
    }
    if (prof && lane == 0) stampA[warp] = clock64();
    __syncthreads();
    const int accepted = s_accept, scn = s_scn;
    off += (scn == 1) ? 3 : 4;
    // ================= stage 2: accepted -> append my slice of the new columns of A / rows of B =================
    if (accepted) {
      const Prep& P = prep[b][s_cur];
      if (tid < 3) {
        fs[3 * i + tid] = P.nw[tid];
        if (blockIdx.x == 0) a.hs[(size_t)3 * N * sl + 3 * i + tid] = P.nw[tid];
      }
      if (warp == 0) {
        // G_eff block of site i+1 gets the new rank-4 term without waiting for anybody:
        //   += (G_eff[i+1+rN, i+kN] M^-1) (Delta G_eff[i+kN, i+1+cN])
        if (have_next) {
          const int r = (lane >> 2) & 3, c = lane & 3;
          cplx acc = cmake(0.0, 0.0);
          if (lane < 16) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) cfma(acc, X1s[r * 4 + kk], Minv[kk * 4 + c]);
            T1s[r * 4 + c] = acc;
          } else {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) cfma(acc, P.D[r * 4 + kk], X2s[kk * 4 + c]);
            T2s[r * 4 + c] = acc;
          }
          // T1 = rows i+1+rN of the new A columns, T2 = columns i+1+cN of the new B rows: the next site's copies are
          // complete without a round trip through memory
          if (lane < 16) As4[(nb * 4 + r) * lds + np + c] = acc;
          else Bs4[(nb * 4 + c) * lds + np + r] = acc;
          __syncwarp();
          if (lane < 16) {
            const int rr = lane & 3, cc = lane >> 2;       // g4e layout [r + 4c]
            cplx g = g4e[nb][lane];
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) cfma(g, T1s[rr * 4 + kk], T2s[kk * 4 + cc]);
            g4e[nb][lane] = g;
          }
        }
      } else {
        // warps 1-7: my slice of the new columns:  A_new[r,:] = (G_eff[r, i+kN] - delta) M^-1,  B_new[:,c] = Delta G_eff[i+kN, c]
        const int t2 = tid - 32;
        cplx* Atw = a.At + batch * bufstride;
        cplx* Bmw = a.Bm + batch * bufstride;
        for (int e = t2; e < nown * 8; e += 224) {
          const bool isB = e >= nown * 4;
          const int tt = isB ? e - nown * 4 : e;
          const int rl = tt >> 2, kq = tt & 3;
          cplx acc = cmake(0.0, 0.0);
          if (!isB) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) cfma(acc, gcc[rl * 4 + kk], Minv[kk * 4 + kq]);
            Aown[(size_t)rl * lds + np + kq] = acc;
            st_pub(Atw + (size_t)(row0 + rl) * ldk + np + kq, acc);
          } else {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) cfma(acc, P.D[kq * 4 + kk], grc[rl * 4 + kk]);
            Bown[(size_t)rl * lds + np + kq] = acc;
            st_pub(Bmw + (size_t)(row0 + rl) * ldk + np + kq, acc);
          }
        }
      }
      kc++;
      np += 4;
    }
    // ================= flush: G += A B over the pending 4*kc columns =================
    const bool do_flush = (kc == a.kmax || (i == N - 1 && kc > 0));
    if (do_flush) {
      n_fl++;
      long long tf0 = clock64();
      if (a.bar_mode) { bar_target += gridDim.x; grid_barrier_mono(bar_ctr, bar_target); } else grid_barrier(a.bar, gridDim.x);
      long long tf1 = clock64();
      const int K = 4 * kc;
      const int lo = lane >> 2, lk = lane & 3;
      const int wm = warp & 1, wn = warp >> 1;           // 2 x 4 warps, warp tile 32 x 16
      // Antiunitary flavour symmetry of the O(3) model, G = [[A, B], [-conj(B), conj(A)]] (flavour blocks (1,2 | 3,4);
      // oracle/experiments/antiunitary_symmetry.py): every accepted update preserves it, so the flush computes the upper
      // half of G only and writes the lower half as its mirror image - half the DMMAs, one tile per CTA at n = 1024.
      const int hN = n >> 1;
      const bool sym = a.sym != 0;
      const int mrows = sym ? hN : n;                    // rows of G this flush computes (the last tile row may be partial)
      const int tiles_m = (n + 63) / 64, tiles_r = (mrows + 63) / 64, ntiles = tiles_r * tiles_m;
      const int nch = (K + 31) / 32;                     // k-chunks of 32 (at most kmax*4/32)
      // operands staged with cp.async, one commit group per k-chunk (A and B together); a CTA's tiles that share their
      // row block keep the A chunks; the G tile is fetched into registers before the DMMAs start
      int tm_loaded = -1;
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int tm0 = (t % tiles_r) * 64, tn0 = (t / tiles_r) * 64;
        const bool loadA = (tm0 != tm_loaded);
        tm_loaded = tm0;
        for (int ch = 0; ch < nch; ++ch) {
          const int k0 = ch * 32;
          cplx* fa = FA + ch * (64 * 36);
          cplx* fb = FB + ch * (64 * 36);
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int e = tid + 256 * u, rr = e >> 5, kk = e & 31;
            const bool kok = (k0 + kk) < K;
            if (loadA) { const bool ok = kok && tm0 + rr < n; cp_async16(fa + rr * 36 + kk, ok ? Atb + (size_t)(tm0 + rr) * ldk + k0 + kk : Atb, ok); }
            { const bool ok = kok && tn0 + rr < n; cp_async16(fb + rr * 36 + kk, ok ? Bmb + (size_t)(tn0 + rr) * ldk + k0 + kk : Bmb, ok); }
          }
          asm volatile("cp.async.commit_group;" ::: "memory");
        }
        // 3M complex product: S1 = Ar Br, S2 = Ai Bi, S3 = (Ar + Ai)(Br + Bi);  Re = S1 - S2,  Im = S3 - S1 - S2
        // (three real DMMAs per complex tile product instead of four; the flush is DMMA-bound)
        double s1[4][2][2], s2[4][2][2], s3[4][2][2];
#pragma unroll
        for (int x = 0; x < 4; ++x)
#pragma unroll
          for (int y = 0; y < 2; ++y)
#pragma unroll
            for (int e = 0; e < 2; ++e) s1[x][y][e] = s2[x][y][e] = s3[x][y][e] = 0.0;
        for (int ch = 0; ch < nch; ++ch) {
          if (ch + 1 < nch) asm volatile("cp.async.wait_group 1;" ::: "memory");
          else asm volatile("cp.async.wait_group 0;" ::: "memory");
          __syncthreads();
          const cplx* fa = FA + ch * (64 * 36);
          const cplx* fb = FB + ch * (64 * 36);
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            cplx av[4], bv[2];
#pragma unroll
            for (int x = 0; x < 4; ++x) av[x] = fa[(wm * 32 + x * 8 + lo) * 36 + ks * 4 + lk];
#pragma unroll
            for (int y = 0; y < 2; ++y) bv[y] = fb[(wn * 16 + y * 8 + lo) * 36 + ks * 4 + lk];
#pragma unroll
            double asum[4], bsum[2];
#pragma unroll
            for (int x = 0; x < 4; ++x) asum[x] = av[x].x + av[x].y;
#pragma unroll
            for (int y = 0; y < 2; ++y) bsum[y] = bv[y].x + bv[y].y;
#pragma unroll
            for (int x = 0; x < 4; ++x)
#pragma unroll
              for (int y = 0; y < 2; ++y) {
                dmma884(s1[x][y][0], s1[x][y][1], av[x].x, bv[y].x);
                dmma884(s2[x][y][0], s2[x][y][1], av[x].y, bv[y].y);
                dmma884(s3[x][y][0], s3[x][y][1], asum[x], bsum[y]);
              }
          }
        }
        __syncthreads();                                   // staging buffers free for the next tile's loads
        // G tile: read late (the accumulators of the 3M product leave no room to hold it during the DMMAs), all loads in flight
        // before the first add
        cplx gv[4][2][2];
#pragma unroll
        for (int x = 0; x < 4; ++x)
#pragma unroll
          for (int y = 0; y < 2; ++y)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int row = tm0 + wm * 32 + x * 8 + lo, col = tn0 + wn * 16 + y * 8 + 2 * lk + e;
              gv[x][y][e] = (row < mrows && col < n) ? ldcg2(a.G + (size_t)col * n + row) : cmake(0.0, 0.0);
            }
#pragma unroll
        for (int x = 0; x < 4; ++x)
#pragma unroll
          for (int y = 0; y < 2; ++y)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int row = tm0 + wm * 32 + x * 8 + lo, col = tn0 + wn * 16 + y * 8 + 2 * lk + e;
              if (row < mrows && col < n) {
                const cplx v = cmake(gv[x][y][e].x + (s1[x][y][e] - s2[x][y][e]),
                                     gv[x][y][e].y + (s3[x][y][e] - s1[x][y][e] - s2[x][y][e]));
                a.G[(size_t)col * n + row] = v;
                if (sym) {   // (r, c) -> (r + n/2, c + n/2) = conj(v) for c < n/2;  (r + n/2, c - n/2) = -conj(v) otherwise
                  const bool left = col < hN;
                  a.G[(size_t)(left ? col + hN : col - hN) * n + row + hN] = left ? cmake(v.x, -v.y) : cmake(-v.x, v.y);
                }
              }
            }
      }
      long long tf2 = clock64();
      if (a.bar_mode) { bar_target += gridDim.x; grid_barrier_mono(bar_ctr, bar_target); } else grid_barrier(a.bar, gridDim.x);
      long long tf3 = clock64();
      if (prof && tid == 0) { pf[0] += tf1 - tf0; pf[1] += tf2 - tf1; pf[2] += tf3 - tf2; }
      {   // re-arm my rows of the buffer just consumed (nobody reads it again before the flush after next)
        cplx* Atw = a.At + batch * bufstride;
        cplx* Bmw = a.Bm + batch * bufstride;
        for (int e = tid; e < nown * K; e += blockDim.x) {
          const int rl = e / K, p = e - rl * K;
          st_sent(Atw + (size_t)(row0 + rl) * ldk + p);
          st_sent(Bmw + (size_t)(row0 + rl) * ldk + p);
        }
      }
      batch ^= 1;
      kc = 0;
      np = 0;
      if (have_next) {   // G changed: refresh what was prefetched for site i+1
        fetch_G(i + 1, nb, tid, 256);
        __syncthreads();
        if (tid < 16) g4e[nb][tid] = g4r[tid];
      }
    }
    s_cur = scn;
    if (prof && lane == 0) stampB[warp] = clock64();
    __syncthreads();
    if (prof && tid == 0) {
      long long relA = stampA[0], relB = stampB[0];
      for (int w = 0; w < 8; ++w) { relA = max(relA, stampA[w]); relB = max(relB, stampB[w]); p_role[w] += stampA[w] - rel_prev; p_role1[w] += stampP[w] - rel_prev; }
      p_s1 += relA - rel_prev;
      if (do_flush) p_flush += relB - relA; else if (accepted) p_rest_acc += relB - relA; else p_rest_rej += relB - relA;
      rel_prev = relB;
    }
  }
  if (prof && tid == 0) {
    // [0] total, [1] stage 1, [2] stage 2 of accepted sites (no flush), [3] iterations with a flush (stage 2 + flush),
    // [4] #flushes, [5] accepts, [6] stage 2 of rejected sites, [8+w] stage-1 role time of warp w
    a.prof[0] = clock64() - t_begin; a.prof[1] = p_s1; a.prof[2] = p_rest_acc; a.prof[3] = p_flush; a.prof[4] = n_fl;
    a.prof[5] = nacc; a.prof[6] = p_rest_rej;
    for (int w = 0; w < 8; ++w) { a.prof[8 + w] = p_role[w]; a.prof[16 + w] = p_role1[w]; }
    a.prof[24] = pf[0]; a.prof[25] = pf[1]; a.prof[26] = pf[2];
  }

  if (blockIdx.x == 0 && tid == 0) {
    *a.pos = pos0 + off;
    *a.accepted += nacc;
    *a.dS += dS_sum;
    if (s_exh) a.flags[0] = 1;
    if (nonreal) a.flags[1] += nonreal;
  }
}

int local_updates_grid(int n, int num_sms, int* rpc) {
  int grid = n / 8;
  if (grid > num_sms) grid = num_sms;
  if (grid < 1) grid = 1;
  int r = (n + grid - 1) / grid;
  grid = (n + r - 1) / r;
  *rpc = r;
  return grid;
}

size_t local_updates_smem(const LUArgs& a) {
  const int ldk = 4 * a.kmax;
  const int lds = ldk + 2;
  return sizeof(cplx) * ((size_t)2 * a.rpc * lds + 16 * lds + 16 * a.rpc + 4 * 64 * 36) + sizeof(double) * 10 * a.nsites +
         sizeof(int) * 4 * a.nsites;
}

int launch_local_updates(cudaStream_t st, const LUArgs& a, int grid) {
  const size_t smem = local_updates_smem(a);
  static SmemMemo memo, memo_prof;
  size_t smem_lim = 0;
  if (ensure_max_dynamic_smem(local_updates_kernel<false>, memo, &smem_lim) || ensure_max_dynamic_smem(local_updates_kernel<true>, memo_prof, &smem_lim)) return -1;
  if (smem > smem_lim) { snprintf(g_errbuf, sizeof(g_errbuf), "local_updates: shared memory %zu > %zu", smem, smem_lim); return -1; }
  if (a.rpc > 64) { snprintf(g_errbuf, sizeof(g_errbuf), "local_updates: rows per CTA %d > 64", a.rpc); return -1; }
  LUArgs args = a;
  void* params[] = {&args};
  const void* kern = a.prof ? (const void*)local_updates_kernel<true> : (const void*)local_updates_kernel<false>;
  CUDA_TRY(cudaLaunchCooperativeKernel(kern, dim3(grid), dim3(256), params, smem, st));
  g_launches++;
  return 0;
}
