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

#define LU_SENT 0xFFFFFFFFFFFFFFFFull

struct Prep {
  double nw[3];      // proposed field value
  double e_dS;       // exp(-dS)
  double mlog;       // -log(exp(-dS))   (local_updates.jl:34)
  double u3;         // the accept draw (used only if p_acc <= 1)
  cplx D[16];        // Delta = e^{+dtau V(old)} e^{-dtau V(new)} - 1, row-major
};

__device__ __forceinline__ cplx ld_valid(const cplx* p) {
  unsigned long long x, y;
  do {
    asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(x), "=l"(y) : "l"(p) : "memory");
  } while (x == LU_SENT || y == LU_SENT);
  return make_double2(__longlong_as_double((long long)x), __longlong_as_double((long long)y));
}
__device__ __forceinline__ void st_pub(cplx* p, cplx v) {
  asm volatile("st.volatile.global.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
}
__device__ __forceinline__ void st_sent(cplx* p) {
  asm volatile("st.volatile.global.v2.u64 [%0], {%1, %1};" ::"l"(p), "l"(LU_SENT) : "memory");
}
__device__ __forceinline__ cplx ldcg2(const cplx* p) { return __ldcg(reinterpret_cast<const double2*>(p)); }

// 4x4 complex e^{-power*dtau*V(op)} (interactions.jl:102-141), element (r,c):  [C S 0 R; cS C -R 0; 0 -R C cS; R 0 S C]
__device__ __forceinline__ cplx evop_elem(int r, int c, double C, cplx S, double R) {
  if (r == c) return cmake(C, 0.0);
  switch (r * 4 + c) {
    case 1: case 14: return S;
    case 4: case 11: return cconj(S);
    case 3: case 12: return cmake(R, 0.0);
    case 6: case 9: return cmake(-R, 0.0);
    default: return cmake(0.0, 0.0);
  }
}

__device__ __forceinline__ cplx det3(cplx a, cplx b, cplx c, cplx d, cplx e, cplx f, cplx g, cplx h, cplx i) {
  cplx t1 = csub(cmul(e, i), cmul(f, h));
  cplx t2 = csub(cmul(d, i), cmul(f, g));
  cplx t3 = csub(cmul(d, h), cmul(e, g));
  return cadd(csub(cmul(a, t1), cmul(b, t2)), cmul(c, t3));
}

// cosh(x) and sinh(x)/x-style pieces from one expm1 (x >= 0, small): accurate without cancellation
__device__ __forceinline__ void cosh_sinh(double x, double* ch, double* sh) {
  const double em1 = expm1(x);
  const double q = em1 / (em1 + 1.0);
  *sh = 0.5 * (em1 + q);
  *ch = 1.0 + 0.5 * em1 * q;
}

// One warp evaluates the proposal at `site` (proposal draws at unif[posp..posp+2], accept draw at posp+3).
// prev_site / prev_new: a site whose field value must be read as prev_new instead of fs[] (or -1).
__device__ __forceinline__ void do_prep(const LUArgs& a, const double* fs, int site, long long posp, int prev_site,
                                        double pn1, double pn2, double pn3, int sl_earlier, int sl_later, Prep* out,
                                        int* exhausted) {
  const int lane = threadIdx.x & 31;
  const int N = a.nsites;
  double u0 = 0.0, u1 = 0.0, u2 = 0.0, u3 = 0.0;
  if (posp + 4 <= a.nunif) { u0 = a.unif[posp]; u1 = a.unif[posp + 1]; u2 = a.unif[posp + 2]; u3 = a.unif[posp + 3]; }
  else *exhausted = 1;
  const double o1 = fs[3 * site], o2 = fs[3 * site + 1], o3 = fs[3 * site + 2];
  // randuniform (dqmc_framework.jl:628): -b + 2*b*rand(); no FMA contraction so the field stays bit-identical
  const double b2 = __dmul_rn(2.0, a.box);
  const double n1 = __dadd_rn(o1, __dadd_rn(-a.box, __dmul_rn(b2, u0)));
  const double n2 = __dadd_rn(o2, __dadd_rn(-a.box, __dmul_rn(b2, u1)));
  const double n3 = __dadd_rn(o3, __dadd_rn(-a.box, __dmul_rn(b2, u2)));
  // calc_boson_action_diff (action.jl:57-101)
  const double d1 = n1 - o1, d2 = n2 - o2, d3 = n3 - o3;
  const double osq = o1 * o1 + o2 * o2 + o3 * o3, nsq = n1 * n1 + n2 * n2 + n3 * n3;
  const double sq_diff = nsq - osq, pow4_diff = nsq * nsq - osq * osq;
  double dS;
  if (!a.edrun) {
    const double* he = a.hs + 3 * ((size_t)site + (size_t)N * sl_earlier);
    const double* hl = a.hs + 3 * ((size_t)site + (size_t)N * sl_later);
    const double t1 = hl[0] + he[0], t2 = hl[1] + he[1], t3 = hl[2] + he[2];
    double s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
    for (int nb = 0; nb < 4; ++nb) {
      const int j = a.nbr[4 * site + nb];
      const bool sub = (j == prev_site);
      s1 += sub ? pn1 : fs[3 * j];
      s2 += sub ? pn2 : fs[3 * j + 1];
      s3 += sub ? pn3 : fs[3 * j + 2];
    }
    dS = a.inv_dtau_c2 * (sq_diff - (t1 * d1 + t2 * d2 + t3 * d3));
    dS += 0.5 * a.dtau * (4.0 * sq_diff - 2.0 * (s1 * d1 + s2 * d2 + s3 * d3));
    dS += a.dtau * (0.5 * a.r * sq_diff + 0.25 * a.u * pow4_diff);
  } else {
    dS = a.dtau * (0.5 * a.r * sq_diff);
  }
  const double e_dS = exp(-dS);
  // interaction_matrix_exp_op!: old value with power -1, new value with power +1
  const double on = sqrt(osq), nn = sqrt(nsq);
  double C1, s1h, C2, s2h;
  cosh_sinh(a.lam_dtau * on, &C1, &s1h);
  cosh_sinh(a.lam_dtau * nn, &C2, &s2h);
  const double sh1 = -s1h / on, sh2 = s2h / nn;
  const cplx S1 = cmake(-o1 * sh1, o2 * sh1), S2 = cmake(-n1 * sh2, n2 * sh2);
  const double R1 = -o3 * sh1, R2 = -n3 * sh2;
  if (lane < 16) {
    const int r = lane >> 2, c = lane & 3;
    cplx acc = cmake(r == c ? -1.0 : 0.0, 0.0);
#pragma unroll
    for (int k = 0; k < 4; ++k) cfma(acc, evop_elem(r, k, C1, S1, R1), evop_elem(k, c, C2, S2, R2));
    out->D[lane] = acc;
  }
  if (lane == 0) {
    out->nw[0] = n1; out->nw[1] = n2; out->nw[2] = n3;
    out->e_dS = e_dS; out->mlog = -log(e_dS); out->u3 = u3;
  }
}

__global__ void __launch_bounds__(256) local_updates_kernel(LUArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int n = a.n, N = a.nsites, ldk = 4 * a.kmax, rpc = a.rpc;
  const size_t bufstride = (size_t)ldk * n;                 // one At / Bm buffer
  cplx* sp = reinterpret_cast<cplx*>(smem_raw);
  cplx* Aown = sp; sp += (size_t)rpc * ldk;                 // [rpc][ldk] my rows of A
  cplx* Bown = sp; sp += (size_t)rpc * ldk;                 // [rpc][ldk] my columns of B
  cplx* As4 = sp; sp += 2 * 4 * ldk;                        // [2][4][ldk] rows site+kN of A (double-buffered by site parity)
  cplx* Bs4 = sp; sp += 2 * 4 * ldk;                        // [2][4][ldk] cols site+kN of B
  cplx* gcol = sp; sp += (size_t)2 * rpc * 4;               // [2][rpc][4] G[r, site+kN] for my rows
  cplx* grow = sp; sp += (size_t)2 * rpc * 4;               // [2][rpc][4] G[site+kN, c] for my cols
  cplx* FA = sp; sp += 64 * 36;                             // flush staging: 64 rows x 32 k (+4 pad)
  cplx* FB = sp; sp += 64 * 36;
  double* fs = reinterpret_cast<double*>(sp);               // [3*N] field of this slice
  __shared__ cplx g4e[2][16], g4r[16], Mm[16], Cof[16], Minv[16], gcc[64 * 4], grc[64 * 4];
  __shared__ Prep prep[2][3];
  __shared__ int s_accept, s_scn, s_exh;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int row0 = blockIdx.x * rpc;
  const int nown = max(0, min(n, row0 + rpc) - row0);
  const int sl = a.slice;
  const int sl_later = (sl + 1) % a.nslices, sl_earlier = (sl + a.nslices - 1) % a.nslices;

  for (int e = tid; e < 3 * N; e += blockDim.x) fs[e] = a.hs[(size_t)3 * N * sl + e];
  if (tid == 0) s_exh = 0;
  long long pos = *a.pos;
  long long nacc = 0;
  double dS_sum = 0.0;
  int kc = 0, np = 0, batch = 0, nonreal = 0, s_cur = 0;
  // optional cycle profile of CTA 0: arrival stamps before the two CTA barriers of an iteration (a clock read right
  // after bar.sync would capture the barrier's issue, not its release)
  const bool prof = (a.prof != nullptr) && blockIdx.x == 0;
  __shared__ long long stampA[8], stampB[8];
  long long p_role[8] = {0, 0, 0, 0, 0, 0, 0, 0}, p_s1 = 0, p_rest_acc = 0, p_rest_rej = 0, p_flush = 0, n_fl = 0, rel_prev = 0;
  __syncthreads();
  const long long t_begin = clock64();
  rel_prev = t_begin;

  // G-derived data of `site` into buffer bb (all threads of the CTA or the 128 prefetch threads; tp = local index)
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
  if (warp == 1) do_prep(a, fs, 0, pos, -1, 0.0, 0.0, 0.0, sl_earlier, sl_later, &prep[0][0], &s_exh);
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
        cplx t0, t1, t2, t3;
        {
          cplx g0 = g4e[b][0 + 4 * c], g1 = g4e[b][1 + 4 * c], g2 = g4e[b][2 + 4 * c], g3 = g4e[b][3 + 4 * c];
          g0 = cmake((c == 0 ? 1.0 : 0.0) - g0.x, -g0.y);
          g1 = cmake((c == 1 ? 1.0 : 0.0) - g1.x, -g1.y);
          g2 = cmake((c == 2 ? 1.0 : 0.0) - g2.x, -g2.y);
          g3 = cmake((c == 3 ? 1.0 : 0.0) - g3.x, -g3.y);
          t0 = cmul(P.D[r * 4 + 0], g0); t1 = cmul(P.D[r * 4 + 1], g1);
          t2 = cmul(P.D[r * 4 + 2], g2); t3 = cmul(P.D[r * 4 + 3], g3);
        }
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
        if (lane < 16) Minv[r * 4 + c] = cdiv(Cof[c * 4 + r], det);
        nacc++;
        dS_sum += P.mlog;
      }
      if (fabs(det.y) > 1e-4 * fabs(det.x)) nonreal++;
    } else if (warp <= 3) {
      if (have_next) {
        const int sc = warp - 1;                                  // 0 rejected, 1 accepted (no draw), 2 accepted (draw)
        const Prep& Pc = prep[b][s_cur];
        const long long posp = pos + (sc == 1 ? 3 : 4);
        do_prep(a, fs, i + 1, posp, sc == 0 ? -1 : i, Pc.nw[0], Pc.nw[1], Pc.nw[2], sl_earlier, sl_later, &prep[nb][sc], &s_exh);
      }
    } else {
      if (have_next) {
        const int tp = tid - 128, site = i + 1;
        for (int e = tp; e < 8 * np; e += 128) {
          const int w = e / (4 * np), rem = e - w * 4 * np, k = rem / np, p = rem - k * np;
          const cplx* X = w ? Bmb : Atb;
          cplx* Xs = (w ? Bs4 : As4) + (nb * 4 + k) * ldk;
          Xs[p] = ld_valid(X + (size_t)(site + k * N) * ldk + p);
        }
        fetch_G(site, nb, tp, 128);
        asm volatile("bar.sync 1, 128;" ::: "memory");
        const int o = tp >> 3, q = tp & 7, r = o & 3, c = o >> 2;   // 16 outputs x 8 partial lanes
        cplx acc = cmake(0.0, 0.0);
        for (int p = q; p < np; p += 8) cfma(acc, As4[(nb * 4 + r) * ldk + p], Bs4[(nb * 4 + c) * ldk + p]);
#pragma unroll
        for (int s = 4; s > 0; s >>= 1) {
          acc.x += __shfl_xor_sync(0xffffffffu, acc.x, s);
          acc.y += __shfl_xor_sync(0xffffffffu, acc.y, s);
        }
        if (q == 0) g4e[nb][o] = cadd(g4r[o], acc);
      }
    }
    if (prof && lane == 0) stampA[warp] = clock64();
    __syncthreads();
    const int accepted = s_accept, scn = s_scn;
    pos += (scn == 1) ? 3 : 4;
    // ================= stage 2: accepted -> append my slice of the new columns of A / rows of B =================
    if (accepted) {
      const Prep& P = prep[b][s_cur];
      if (tid < 3) {
        fs[3 * i + tid] = P.nw[tid];
        if (blockIdx.x == 0) a.hs[(size_t)3 * N * sl + 3 * i + tid] = P.nw[tid];
      }
      {   // G_eff[r, i+kN] for my rows and G_eff[i+kN, c] for my columns: 4 threads per dot product over the pending columns
        const int task = tid >> 2, q = tid & 3;
        const int half = nown * 4;
        for (int t0 = 0; t0 < 2 * half; t0 += 64) {
          const int t = t0 + task;
          const bool act = t < 2 * half;
          const bool isB = t >= half;
          const int tt = isB ? t - half : t;
          const int rl = tt >> 2, k = tt & 3;
          cplx acc = cmake(0.0, 0.0);
          if (act) {
            const cplx* own = (isB ? Bown : Aown) + (size_t)rl * ldk;
            const cplx* site = (isB ? As4 : Bs4) + (b * 4 + k) * ldk;
            for (int p = q; p < np; p += 4) cfma(acc, own[p], site[p]);
          }
          acc.x += __shfl_xor_sync(0xffffffffu, acc.x, 1); acc.y += __shfl_xor_sync(0xffffffffu, acc.y, 1);
          acc.x += __shfl_xor_sync(0xffffffffu, acc.x, 2); acc.y += __shfl_xor_sync(0xffffffffu, acc.y, 2);
          if (act && q == 0) {
            if (!isB) {
              cplx v = cadd(gcol[b * rpc * 4 + tt], acc);
              if (row0 + rl == i + k * N) v.x -= 1.0;
              gcc[tt] = v;
            } else {
              grc[tt] = cadd(grow[b * rpc * 4 + tt], acc);
            }
          }
        }
      }
      __syncthreads();
      cplx* Atw = a.At + batch * bufstride;
      cplx* Bmw = a.Bm + batch * bufstride;
      for (int e = tid; e < nown * 8; e += blockDim.x) {
        const bool isB = e >= nown * 4;
        const int tt = isB ? e - nown * 4 : e;
        const int rl = tt >> 2, k = tt & 3;
        cplx acc = cmake(0.0, 0.0);
        if (!isB) {   // A_new[r,:] = (G_eff[r, i+kN] - delta) M^-1
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) cfma(acc, gcc[rl * 4 + kk], Minv[kk * 4 + k]);
          Aown[(size_t)rl * ldk + np + k] = acc;
          st_pub(Atw + (size_t)(row0 + rl) * ldk + np + k, acc);
        } else {      // B_new[:,c] = Delta G_eff[i+kN, c]
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) cfma(acc, P.D[k * 4 + kk], grc[rl * 4 + kk]);
          Bown[(size_t)rl * ldk + np + k] = acc;
          st_pub(Bmw + (size_t)(row0 + rl) * ldk + np + k, acc);
        }
      }
      if (have_next) {   // the 4 new columns for the rows/columns of site i+1 (spin until their owners have published them)
        if (tid < 32) {
          const int w = tid >> 4, sub = tid & 15, k = sub >> 2, kp = sub & 3;
          const cplx* X = w ? Bmw : Atw;
          cplx* Xs = (w ? Bs4 : As4) + (nb * 4 + k) * ldk;
          Xs[np + kp] = ld_valid(X + (size_t)(i + 1 + k * N) * ldk + np + kp);
        }
        __syncthreads();
        if (tid < 16) {
          const int r = tid & 3, c = tid >> 2;
          cplx acc = g4e[nb][tid];
#pragma unroll
          for (int kp = 0; kp < 4; ++kp) cfma(acc, As4[(nb * 4 + r) * ldk + np + kp], Bs4[(nb * 4 + c) * ldk + np + kp]);
          g4e[nb][tid] = acc;
        }
      }
      kc++;
      np += 4;
    }
    // ================= flush: G += A B over the pending 4*kc columns =================
    const bool do_flush = (kc == a.kmax || (i == N - 1 && kc > 0));
    if (do_flush) {
      n_fl++;
      grid_barrier(a.bar, gridDim.x);
      const int K = 4 * kc;
      const int lo = lane >> 2, lk = lane & 3;
      const int wm = warp & 1, wn = warp >> 1;           // 2 x 4 warps, warp tile 32 x 16
      const int tiles_m = (n + 63) / 64, ntiles = tiles_m * tiles_m;
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int tm0 = (t % tiles_m) * 64, tn0 = (t / tiles_m) * 64;
        double cr[4][2][2], ci[4][2][2];
#pragma unroll
        for (int x = 0; x < 4; ++x)
#pragma unroll
          for (int y = 0; y < 2; ++y) cr[x][y][0] = cr[x][y][1] = ci[x][y][0] = ci[x][y][1] = 0.0;
        for (int k0 = 0; k0 < K; k0 += 32) {
          __syncthreads();
          for (int e = tid; e < 64 * 32; e += blockDim.x) {
            const int rr = e >> 5, kk = e & 31;
            const bool kok = (k0 + kk) < K;
            FA[rr * 36 + kk] = (kok && tm0 + rr < n) ? ldcg2(Atb + (size_t)(tm0 + rr) * ldk + k0 + kk) : cmake(0.0, 0.0);
            FB[rr * 36 + kk] = (kok && tn0 + rr < n) ? ldcg2(Bmb + (size_t)(tn0 + rr) * ldk + k0 + kk) : cmake(0.0, 0.0);
          }
          __syncthreads();
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            cplx av[4], bv[2];
#pragma unroll
            for (int x = 0; x < 4; ++x) av[x] = FA[(wm * 32 + x * 8 + lo) * 36 + ks * 4 + lk];
#pragma unroll
            for (int y = 0; y < 2; ++y) bv[y] = FB[(wn * 16 + y * 8 + lo) * 36 + ks * 4 + lk];
#pragma unroll
            for (int x = 0; x < 4; ++x)
#pragma unroll
              for (int y = 0; y < 2; ++y) {
                dmma884(cr[x][y][0], cr[x][y][1], av[x].x, bv[y].x);
                dmma884(cr[x][y][0], cr[x][y][1], -av[x].y, bv[y].y);
                dmma884(ci[x][y][0], ci[x][y][1], av[x].x, bv[y].y);
                dmma884(ci[x][y][0], ci[x][y][1], av[x].y, bv[y].x);
              }
          }
        }
#pragma unroll
        for (int x = 0; x < 4; ++x)
#pragma unroll
          for (int y = 0; y < 2; ++y)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int row = tm0 + wm * 32 + x * 8 + lo, col = tn0 + wn * 16 + y * 8 + 2 * lk + e;
              if (row < n && col < n) {
                cplx* p = a.G + (size_t)col * n + row;
                cplx v = ldcg2(p);
                v.x += cr[x][y][e]; v.y += ci[x][y][e];
                *p = v;
              }
            }
      }
      grid_barrier(a.bar, gridDim.x);
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
      for (int w = 0; w < 8; ++w) { relA = max(relA, stampA[w]); relB = max(relB, stampB[w]); p_role[w] += stampA[w] - rel_prev; }
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
    for (int w = 0; w < 8; ++w) a.prof[8 + w] = p_role[w];
  }

  if (blockIdx.x == 0 && tid == 0) {
    *a.pos = pos;
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
  return sizeof(cplx) * ((size_t)2 * a.rpc * ldk + 16 * ldk + 16 * a.rpc + 2 * 64 * 36) + sizeof(double) * 3 * a.nsites;
}

int launch_local_updates(cudaStream_t st, const LUArgs& a, int grid) {
  const size_t smem = local_updates_smem(a);
  static size_t smem_lim = 0;
  if (smem_lim == 0 && set_max_dynamic_smem(local_updates_kernel, &smem_lim)) return -1;
  if (smem > smem_lim) { snprintf(g_errbuf, sizeof(g_errbuf), "local_updates: shared memory %zu > %zu", smem, smem_lim); return -1; }
  if (a.rpc > 64) { snprintf(g_errbuf, sizeof(g_errbuf), "local_updates: rows per CTA %d > 64", a.rpc); return -1; }
  LUArgs args = a;
  void* params[] = {&args};
  CUDA_TRY(cudaLaunchCooperativeKernel((const void*)local_updates_kernel, dim3(grid), dim3(256), params, smem, st));
  g_launches++;
  return 0;
}
