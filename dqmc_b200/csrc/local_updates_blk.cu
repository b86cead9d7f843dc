#include "lu_common.cuh"

// =====================================================================================================
// Block-lookahead local-update kernel (one launch per time slice), the default for rows-per-CTA <= 16.
//
// As in local_updates.cu all CTAs run the Metropolis loop of the slice in lock step, take every decision redundantly
// from identical data, and CTA b owns rows [b*rpc, (b+1)*rpc) of the pending factor A (n x 4k) and the same columns of
// B (4k x n) of the delayed update G_eff = G + A B.  What changes is how a decision gets its 4x4 block of G_eff:
//
//   * the sites are processed in blocks of LU_BS = 8 consecutive sites.  At the start of a block every CTA gathers, for
//     the 32 rows/columns {s + k N} of the block's sites (the "window"), the pending columns of A and B (published by
//     their owners; consumers spin on a NaN sentinel) and the matching entries of G, and forms with DMMA tiles
//         Sw = G_eff[window, window]   (32 x 32; upper half + mirror when G has the antiunitary flavour symmetry)
//         GC = G_eff[own rows, window] (rpc x 32)          GR = G_eff[window, own columns] (32 x rpc);
//   * inside the block nothing leaves the SM: warp 0 decides site j from Sw, warps 1-3 evaluate the proposal of site
//     j+1 under the three possible outcomes of site j (which keeps the reference's conditional RNG consumption,
//     local_updates.jl:31); an accepted proposal is applied as an exact rank-4 update to the not yet visited part of
//     Sw, GC, GR, and the CTA's slice of the new columns of A / rows of B -- (GC[:, site] - delta) M^-1 and
//     Delta GR[site, :] -- is appended locally and published for the other CTAs' next gathers and for the flush;
//   * a block boundary with more than kmax accepted updates pending (or the end of the slice) flushes G += A B with
//     DMMA tiles between two grid barriers (3M products, symmetric half), exactly as before.
//
// One L2 round trip and two CTA barriers per SITE become one gather per BLOCK; the per-site critical path is the
// decision itself.  Reference: local_updates.jl:1-95 (local_updates, calc_detratio, update_greens!).
// =====================================================================================================

#define LU_BS 8          // sites per block
#define LU_W 32          // window size = 4 flavours x LU_BS
#define LU_SWLD 33       // row stride of Sw / GC (complex): odd -> column reads are conflict-free
#define LU_FCH 16        // k-chunk of the flush staging
#define LU_FLD 20        // row stride of the flush staging (16 + 4: conflict-free DMMA fragment loads)
#define LU_FST 3         // stages of the flush staging ring

__host__ __device__ __forceinline__ int lu_ld(int k) { return ((k + 7) & ~7) + 4; }   // stride = 4 (mod 8): conflict-free fragments

struct BlkLayout {
  int rpt, ldo, ldb;
  size_t off_Aown, off_Bown, off_fs, off_union, total;   // bytes
  size_t u_Ablk, u_Bblk, u_Sw, u_GC, u_GR, u_U, u_V, u_UA, u_VB, u_window_end;   // offsets inside the union (bytes)
  size_t u_FA, u_FB, u_flush_end;
};

__host__ __device__ inline BlkLayout blk_layout(int rpc, int kmax, int nsites) {
  BlkLayout L;
  L.rpt = (rpc + 7) & ~7;
  L.ldo = lu_ld(4 * (kmax + LU_BS));
  L.ldb = lu_ld(4 * kmax);
  size_t o = 0;
  L.off_Aown = o; o += sizeof(cplx) * (size_t)L.rpt * L.ldo;
  L.off_Bown = o; o += sizeof(cplx) * (size_t)L.rpt * L.ldo;
  L.off_fs = o; o += sizeof(double) * 10 * (size_t)nsites + sizeof(int) * 4 * (size_t)nsites;
  o = (o + 15) & ~(size_t)15;
  L.off_union = o;
  size_t u = 0;
  L.u_Ablk = u; u += sizeof(cplx) * LU_W * L.ldb;
  L.u_Bblk = u; u += sizeof(cplx) * LU_W * L.ldb;
  L.u_Sw = u; u += sizeof(cplx) * LU_W * LU_SWLD;
  L.u_GC = u; u += sizeof(cplx) * (size_t)L.rpt * LU_SWLD;
  L.u_GR = u; u += sizeof(cplx) * LU_W * (size_t)(L.rpt + 1);
  L.u_U = u; u += sizeof(cplx) * LU_W * 4;
  L.u_V = u; u += sizeof(cplx) * 4 * LU_W;
  L.u_UA = u; u += sizeof(cplx) * (size_t)L.rpt * 4;
  L.u_VB = u; u += sizeof(cplx) * 4 * (size_t)L.rpt;
  L.u_window_end = u;
  size_t f = 0;
  L.u_FA = f; f += sizeof(cplx) * LU_FST * 64 * LU_FLD;
  L.u_FB = f; f += sizeof(cplx) * LU_FST * 64 * LU_FLD;
  L.u_flush_end = f;
  L.total = L.off_union + (u > f ? u : f);
  return L;
}

// two 8x8 complex tiles  D_t += X_t[0:8, 0:K) * Y_t[0:8, 0:K)^T  (rows contiguous in k) on DMMA.8x8x4, conventional 4-multiplication
// complex product, eight independent accumulator chains.  d[t][0..1] real parts, d[t][2..3] imaginary parts of D[lane/4][2*(lane%4)+{0,1}].
__device__ __forceinline__ void tile_pair_mma(const cplx* X0, int ldx0, const cplx* Y0, int ldy0, const cplx* X1, int ldx1, const cplx* Y1,
                                              int ldy1, bool two, int K, double (&d)[2][4]) {
  const int lane = threadIdx.x & 31, r = lane >> 2, k = lane & 3;
  const cplx* px0 = X0 + r * ldx0 + k; const cplx* py0 = Y0 + r * ldy0 + k;
  const cplx* px1 = X1 + r * ldx1 + k; const cplx* py1 = Y1 + r * ldy1 + k;
  double a0[4] = {0, 0, 0, 0}, b0[4] = {0, 0, 0, 0}, a1[4] = {0, 0, 0, 0}, b1[4] = {0, 0, 0, 0};   // a: Xr Yr | Xr Yi ; b: -Xi Yi | Xi Yr
#pragma unroll 2
  for (int p = 0; p < K; p += 4) {
    const cplx x0 = px0[p], y0 = py0[p];
    dmma884(a0[0], a0[1], x0.x, y0.x);
    dmma884(b0[0], b0[1], -x0.y, y0.y);
    dmma884(a0[2], a0[3], x0.x, y0.y);
    dmma884(b0[2], b0[3], x0.y, y0.x);
    if (two) {
      const cplx x1 = px1[p], y1 = py1[p];
      dmma884(a1[0], a1[1], x1.x, y1.x);
      dmma884(b1[0], b1[1], -x1.y, y1.y);
      dmma884(a1[2], a1[3], x1.x, y1.y);
      dmma884(b1[2], b1[3], x1.y, y1.x);
    }
  }
#pragma unroll
  for (int e = 0; e < 4; ++e) { d[0][e] = a0[e] + b0[e]; d[1][e] = a1[e] + b1[e]; }
}

template <bool PROF>
__global__ void __launch_bounds__(256) lu_block_kernel(LUArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int n = a.n, N = a.nsites, rpc = a.rpc, kmax = a.kmax;
  const int ldk = 4 * (kmax + LU_BS);                       // global row stride of At / Bm (pending capacity: kmax + LU_BS updates)
  const size_t bufstride = (size_t)ldk * n;
  const BlkLayout Lo = blk_layout(rpc, kmax, N);
  const int rpt = Lo.rpt, ldo = Lo.ldo, ldb = Lo.ldb, grld = rpt + 1;
  cplx* Aown = reinterpret_cast<cplx*>(smem_raw + Lo.off_Aown);   // [rpt][ldo] my rows of A
  cplx* Bown = reinterpret_cast<cplx*>(smem_raw + Lo.off_Bown);   // [rpt][ldo] my columns of B
  double* fs = reinterpret_cast<double*>(smem_raw + Lo.off_fs);   // [3N] field of this slice
  double* tn = fs + 3 * N;                                        // [3N] phi(l+1) + phi(l-1)
  double* uw = tn + 3 * N;                                        // [4N] this slice's window of the uniform stream
  int* nbr = reinterpret_cast<int*>(uw + 4 * N);                  // [4N] spatial neighbours
  unsigned char* un = smem_raw + Lo.off_union;
  cplx* Ablk = reinterpret_cast<cplx*>(un + Lo.u_Ablk);           // [32][ldb] pending columns of A for the window rows
  cplx* Bblk = reinterpret_cast<cplx*>(un + Lo.u_Bblk);           // [32][ldb] pending rows of B for the window columns
  cplx* Sw = reinterpret_cast<cplx*>(un + Lo.u_Sw);               // [32][33]  G_eff[window, window]
  cplx* GC = reinterpret_cast<cplx*>(un + Lo.u_GC);               // [rpt][33] G_eff[own rows, window]
  cplx* GR = reinterpret_cast<cplx*>(un + Lo.u_GR);               // [32][rpt+1] G_eff[window, own cols]
  cplx* Us = reinterpret_cast<cplx*>(un + Lo.u_U);                // [32][4]
  cplx* Vs = reinterpret_cast<cplx*>(un + Lo.u_V);                // [4][32]
  cplx* UAs = reinterpret_cast<cplx*>(un + Lo.u_UA);              // [rpt][4]
  cplx* VBs = reinterpret_cast<cplx*>(un + Lo.u_VB);              // [4][rpt]
  cplx* FA = reinterpret_cast<cplx*>(un + Lo.u_FA);               // flush staging (aliases the window: dead during a flush)
  cplx* FB = reinterpret_cast<cplx*>(un + Lo.u_FB);
  __shared__ cplx Mm[16], Cof[16], g4s[16], X1s[16], X2s[16], T1s[16], T2s[16];
  __shared__ cplx Minv2[2][16], Dl2[2][16];                       // M^-1 and Delta of the site being applied (double-buffered by site parity)
  __shared__ Prep prep[2][3];
  __shared__ int s_accept2[2], s_scn2[2], s_exh;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int row0 = blockIdx.x * rpc;
  const int nown = max(0, min(n, row0 + rpc) - row0);
  const int sl = a.slice;
  const int sl_later = (sl + 1) % a.nslices, sl_earlier = (sl + a.nslices - 1) % a.nslices;
  const long long pos0 = *a.pos;
  const int navail = (int)max(0LL, min((long long)4 * N, a.nunif - pos0));
  const int hN = n >> 1;
  const bool sym = a.sym != 0;

  for (int e = tid; e < 3 * N; e += blockDim.x) {
    fs[e] = a.hs[(size_t)3 * N * sl + e];
    tn[e] = a.hs[(size_t)3 * N * sl_later + e] + a.hs[(size_t)3 * N * sl_earlier + e];
  }
  for (int e = tid; e < 4 * N; e += blockDim.x) {
    nbr[e] = a.nbr[e];
    uw[e] = e < navail ? a.unif[pos0 + e] : 0.0;
  }
  for (int e = tid; e < rpt * ldo; e += blockDim.x) { Aown[e] = cmake(0.0, 0.0); Bown[e] = cmake(0.0, 0.0); }   // padding rows stay zero
  if (tid == 0) s_exh = 0;
  unsigned int* const bar_ctr = a.bar + 2 + a.bar_parity;   // monotonic barrier counter of this launch (zero at launch)
  unsigned int bar_target = 0;
  if (blockIdx.x == 0 && tid == 0) a.bar[2 + (1 - a.bar_parity)] = 0;   // the next launch's counter
  int off = 0;                                              // stream position relative to pos0
  long long nacc = 0;
  double dS_sum = 0.0;
  int kc = 0, np = 0, batch = 0, nonreal = 0, s_cur = 0;
  long long pf[5] = {0, 0, 0, 0, 0};                         // PROF, per flush: barrier 1, staging + DMMA, G read-modify-write, barrier 2, re-arm
  long long pr[8] = {0, 0, 0, 0, 0, 0, 0, 0};               // PROF: [0] gather [1] form [2] stage 1 [3] stage 2 [4] flush [5] #flush
  __syncthreads();
  // every CTA has read the field, the neighbour sums and the stream position: only now may CTA 0 write accepted field values
  // back into hs (and *pos at the end).  Co-residency does not mean simultaneous start.
  bar_target += gridDim.x; grid_barrier_mono(bar_ctr, bar_target);
  const long long t_begin = PROF ? clock64() : 0;

  if (warp == 1) do_prep(a, fs, tn, nbr, uw, 0, navail, 0, -1, 0.0, 0.0, 0.0, &prep[0][0], &s_exh);
  __syncthreads();

  for (int s0 = 0; s0 < N; s0 += LU_BS) {
    const int nb = min(LU_BS, N - s0);                        // sites of this block
    cplx* Atw = a.At + batch * bufstride;
    cplx* Bmw = a.Bm + batch * bufstride;
    // ================= block start: gather the window, form Sw / GC / GR =================
    {
      long long t0 = PROF ? clock64() : 0;
      {
        // (1) pending columns: thread = (window index w, 8 threads per row), all loads in flight before the first check
        const int w = tid >> 3, p0 = tid & 7;
        const bool wok = (w & 7) < nb;
        const int grow = s0 + (w & 7) + (w >> 3) * N;        // global row (for A) = global column (for B) of window index w
        const cplx* srcA = Atw + (size_t)grow * ldk;
        const cplx* srcB = Bmw + (size_t)grow * ldk;
        unsigned long long ax[8], ay[8], bx[8], by[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int p = p0 + 8 * u;
          if (wok && p < np) {
            asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(ax[u]), "=l"(ay[u]) : "l"(srcA + p) : "memory");
            asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(bx[u]), "=l"(by[u]) : "l"(srcB + p) : "memory");
          }
        }
        // (2) G entries (G changes only in flushes, which end with a grid barrier: plain L2 loads)
        cplx gw[4], gc[2], gr[2];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int e = tid + 256 * u, ww = e & 31, wc = e >> 5;
          const bool ok = ((ww & 7) < nb) && ((wc & 7) < nb) && (!sym || ww < 16);
          const int gr_ = s0 + (ww & 7) + (ww >> 3) * N, gc_ = s0 + (wc & 7) + (wc >> 3) * N;
          gw[u] = ok ? ldcg2(a.G + (size_t)gc_ * n + gr_) : cmake(0.0, 0.0);
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int e = tid + 256 * u;
          {
            const int rl = e & (rpt - 1), wc = e / rpt;       // GC[rl][wc]
            const bool ok = wc < LU_W && rl < nown && (wc & 7) < nb;
            const int gc_ = s0 + (wc & 7) + (wc >> 3) * N;
            gc[u] = ok ? ldcg2(a.G + (size_t)gc_ * n + row0 + rl) : cmake(0.0, 0.0);
          }
          {
            const int ww = e & 31, cl = e >> 5;               // GR[ww][cl]
            const bool ok = cl < rpt && cl < nown && (ww & 7) < nb;
            const int gr_ = s0 + (ww & 7) + (ww >> 3) * N;
            gr[u] = ok ? ldcg2(a.G + (size_t)(row0 + cl) * n + gr_) : cmake(0.0, 0.0);
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) { const int e = tid + 256 * u; Sw[(e & 31) * LU_SWLD + (e >> 5)] = gw[u]; }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int e = tid + 256 * u;
          if (e / rpt < LU_W) GC[(e & (rpt - 1)) * LU_SWLD + e / rpt] = gc[u];
          if ((e >> 5) < rpt) GR[(e & 31) * grld + (e >> 5)] = gr[u];
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int p = p0 + 8 * u;
          if (p < np) {
            cplx va = cmake(0.0, 0.0), vb = va;
            if (wok) {
              while (ax[u] == LU_SENT || ay[u] == LU_SENT)
                asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(ax[u]), "=l"(ay[u]) : "l"(srcA + p) : "memory");
              while (bx[u] == LU_SENT || by[u] == LU_SENT)
                asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(bx[u]), "=l"(by[u]) : "l"(srcB + p) : "memory");
              va = make_double2(__longlong_as_double((long long)ax[u]), __longlong_as_double((long long)ay[u]));
              vb = make_double2(__longlong_as_double((long long)bx[u]), __longlong_as_double((long long)by[u]));
            }
            Ablk[w * ldb + p] = va;
            Bblk[w * ldb + p] = vb;
          }
        }
      }
      __syncthreads();
      long long t1 = PROF ? clock64() : 0;
      if (np > 0) {
        // (3) Sw += Ablk Bblk^T, GC += Aown Bblk^T, GR += Ablk Bown^T on DMMA tiles; tile list: Sw (upper half if sym), GC, GR
        const int rt = rpt >> 3;
        const int nsw = sym ? 8 : 16, ngc = rt * 4, ntl = nsw + 2 * ngc;
        auto tile_ptrs = [&](int t, const cplx*& X, int& ldx, const cplx*& Y, int& ldy, cplx*& D, int& ldd) {
          if (t < nsw) { const int tm = t >> 2, tnn = t & 3; X = Ablk + tm * 8 * ldb; ldx = ldb; Y = Bblk + tnn * 8 * ldb; ldy = ldb; D = Sw + tm * 8 * LU_SWLD + tnn * 8; ldd = LU_SWLD; }
          else if (t < nsw + ngc) { const int q = t - nsw, tm = q >> 2, tnn = q & 3; X = Aown + tm * 8 * ldo; ldx = ldo; Y = Bblk + tnn * 8 * ldb; ldy = ldb; D = GC + tm * 8 * LU_SWLD + tnn * 8; ldd = LU_SWLD; }
          else { const int q = t - nsw - ngc, tm = q & 3, tnn = q >> 2; X = Ablk + tm * 8 * ldb; ldx = ldb; Y = Bown + tnn * 8 * ldo; ldy = ldo; D = GR + tm * 8 * grld + tnn * 8; ldd = grld; }
        };
        for (int t = warp; t < ntl; t += 16) {
          const bool two = (t + 8) < ntl;
          const cplx *X0, *Y0, *X1, *Y1; cplx *D0, *D1; int lx0, ly0, lx1, ly1, ld0, ld1;
          tile_ptrs(t, X0, lx0, Y0, ly0, D0, ld0);
          tile_ptrs(two ? t + 8 : t, X1, lx1, Y1, ly1, D1, ld1);
          double d[2][4];
          tile_pair_mma(X0, lx0, Y0, ly0, X1, lx1, Y1, ly1, two, np, d);
          const int r = lane >> 2, c2 = 2 * (lane & 3);
          { cplx* q = D0 + r * ld0 + c2; q[0].x += d[0][0]; q[0].y += d[0][2]; q[1].x += d[0][1]; q[1].y += d[0][3]; }
          if (two) { cplx* q = D1 + r * ld1 + c2; q[0].x += d[1][0]; q[0].y += d[1][2]; q[1].x += d[1][1]; q[1].y += d[1][3]; }
        }
        __syncthreads();
      }
      if (sym) {   // lower half of the window from the upper one:  (w, c) -> (w+16, c+16) = conj, (w+16, c-16) = -conj
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int e = tid + 256 * u, ww = e >> 5, wc = e & 31;
          const cplx v = Sw[ww * LU_SWLD + wc];
          const bool left = wc < 16;
          Sw[(ww + 16) * LU_SWLD + (left ? wc + 16 : wc - 16)] = left ? cmake(v.x, -v.y) : cmake(-v.x, v.y);
        }
        __syncthreads();
      }
      if (PROF) { const long long t2 = clock64(); pr[0] += t1 - t0; pr[1] += t2 - t1; }
    }
    long long ts0 = PROF ? clock64() : 0;
    // ================= the sites of the block: three roles, no CTA-wide barrier =================
    //   warp 0      decides site j from its private copy g4s of the site's 4x4 block of G_eff, then derives the block of site
    //               j+1 itself (Sw as of site j-1 plus the rank-4 term of site j), so the next decision never waits for the
    //               window update of the current one;
    //   warps 1-3   proposal of site j+1 under the three possible outcomes of site j;
    //   warps 4-7   window update of site j (U, V, rank-4 DMMA tiles on Sw, GC, GR; my slice of the new columns of A / rows of
    //               B appended and published), overlapped with the decision of site j+1.
    // Named barriers: P (warps 0-3, once per site), GO (warp 0 arrives, warps 4-7 wait: "site j decided"), DONE (warps 4-7
    // arrive, warp 0 waits: "window exact up to site j"), UPD (warps 4-7).
    if (warp == 0 && lane < 16) g4s[lane] = Sw[((lane >> 2) * 8) * LU_SWLD + (lane & 3) * 8];     // block of site s0: [r*4+c]
    __syncwarp();
    for (int j = 0; j < nb; ++j) {
      const int i = s0 + j, b = i & 1, nbuf = b ^ 1, jb = j & 1;
      const bool have_next = (i + 1 < N);
      int accepted, scn;
      if (warp == 0) {
        const Prep& P = prep[b][s_cur];
        const int r = (lane >> 2) & 3, c = lane & 3;
        if (lane < 16) {   // M = 1 + Delta * (1 - G_eff[site block])
          cplx g0 = g4s[0 + c], g1 = g4s[4 + c], g2 = g4s[8 + c], g3 = g4s[12 + c];
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
        if (p_acc > 1.0) { accepted = 1; scn = 1; }
        else { accepted = (P.u3 < p_acc) ? 1 : 0; scn = accepted ? 2 : 0; }
        if (lane == 0) { s_accept2[jb] = accepted; s_scn2[jb] = scn; }
        cplx minv = cmake(0.0, 0.0);
        if (accepted) {
          if (lane < 16) {
            const double id = 1.0 / (det.x * det.x + det.y * det.y);
            const cplx dinv = cmake(det.x * id, -det.y * id);
            minv = cmul(Cof[c * 4 + r], dinv);
            Minv2[jb][r * 4 + c] = minv;
            Dl2[jb][lane] = P.D[lane];
          }
          nacc++;
          dS_sum += P.mlog;
        }
        if (fabs(det.y) > 1e-4 * fabs(det.x)) nonreal++;
        __syncwarp();
        __syncwarp();   // named barriers are .aligned: the whole warp executes them together
        if (j > 0) asm volatile("bar.sync 3, 160;" ::: "memory");        // DONE(j-1): the window is exact up to site j-1
        // what the block of site j+1 needs from the window, read BEFORE the update of site j may start
        cplx blk = cmake(0.0, 0.0), x1 = blk, x2 = blk;
        if (j + 1 < nb && lane < 16) {
          blk = Sw[(r * 8 + j + 1) * LU_SWLD + c * 8 + j + 1];
          x1 = Sw[(r * 8 + j + 1) * LU_SWLD + c * 8 + j];                 // X1[r][kk = c] = G_eff[row r of site j+1, col kk of site j]
          x2 = Sw[(r * 8 + j) * LU_SWLD + c * 8 + j + 1];                 // X2[kk = r][c] = G_eff[row kk of site j, col c of site j+1]
        }
        __syncwarp();   // named barriers are .aligned: the whole warp executes them together
        asm volatile("bar.arrive 2, 160;" ::: "memory");                  // GO(j)
        if (j + 1 < nb) {
          if (accepted) {   // g4(j+1) = blk + (X1 M^-1)(Delta X2)
            if (lane < 16) { X1s[lane] = x1; X2s[lane] = x2; }
            __syncwarp();
            if (lane < 16) {
              cplx acc = cmake(0.0, 0.0);
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) cfma(acc, X1s[r * 4 + kk], Minv2[jb][kk * 4 + c]);
              T1s[lane] = acc;
            } else {
              cplx acc = cmake(0.0, 0.0);
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) cfma(acc, P.D[r * 4 + kk], X2s[kk * 4 + c]);
              T2s[lane - 16] = acc;
            }
            __syncwarp();
            if (lane < 16) {
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) cfma(blk, T1s[r * 4 + kk], T2s[kk * 4 + c]);
            }
          }
          if (lane < 16) g4s[lane] = blk;
        }
        const double nw_mine = lane < 3 ? P.nw[lane] : 0.0;                // read before P(j): the proposal warps reuse prep[b] after it
        __syncwarp();   // named barriers are .aligned: the whole warp executes them together
        asm volatile("bar.sync 1, 128;" ::: "memory");                    // P(j)
        // the accepted value becomes the field only now: the proposals of site i+1 (which may still load fs[3i..] for the rejected
        // scenario) are complete, and site i is no neighbour of site i+2 whose proposals start here
        if (accepted && lane < 3) {
          fs[3 * i + lane] = nw_mine;
          if (blockIdx.x == 0) a.hs[(size_t)3 * N * sl + 3 * i + lane] = nw_mine;
        }
      } else if (warp <= 3) {
        if (have_next) {
          const int sc = warp - 1;                                  // 0 rejected, 1 accepted (no draw), 2 accepted (draw)
          const Prep& Pc = prep[b][s_cur];
          do_prep(a, fs, tn, nbr, uw, off + (sc == 1 ? 3 : 4), navail, i + 1, sc == 0 ? -1 : i, Pc.nw[0], Pc.nw[1], Pc.nw[2],
                  &prep[nbuf][sc], &s_exh);
        }
        __syncwarp();   // named barriers are .aligned: the whole warp executes them together
        asm volatile("bar.sync 1, 128;" ::: "memory");                    // P(j)
        scn = s_scn2[jb];
        accepted = scn != 0;
      } else {
        const int t4 = tid - 128;                                         // 0..127
        __syncwarp();   // named barriers are .aligned: the whole warp executes them together
        asm volatile("bar.sync 2, 160;" ::: "memory");                    // GO(j)
        accepted = s_accept2[jb];
        scn = 0;
        if (accepted) {
          const cplx* Mi = Minv2[jb];
          const cplx* Dl = Dl2[jb];
          // (a) U = (Sw[:, site] - delta) M^-1 (32 x 4), V = Delta Sw[site, :] (4 x 32); UA / VB = the same for my rows / columns
          //     = my slice of the new columns of A / rows of B
          {
            const int ww = t4 >> 2, k = t4 & 3;
            cplx accU = cmake(0.0, 0.0), accV = accU;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              cplx g = Sw[ww * LU_SWLD + kk * 8 + j];
              if (ww == kk * 8 + j) g.x -= 1.0;
              cfma(accU, g, Mi[kk * 4 + k]);
              cfma(accV, Dl[k * 4 + kk], Sw[(kk * 8 + j) * LU_SWLD + ww]);
            }
            Us[ww * 4 + k] = accU;
            Vs[k * LU_W + ww] = accV;
          }
          for (int e = t4; e < 8 * rpt; e += 128) {
            const int q = e % (4 * rpt), rl = q >> 2, k = q & 3;
            cplx acc = cmake(0.0, 0.0);
            if (e < 4 * rpt) {
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) {
                cplx g = GC[rl * LU_SWLD + kk * 8 + j];
                if (row0 + rl == i + kk * N) g.x -= 1.0;
                cfma(acc, g, Mi[kk * 4 + k]);
              }
              if (rl >= nown) acc = cmake(0.0, 0.0);
              UAs[rl * 4 + k] = acc;
              Aown[rl * ldo + np + k] = acc;
              if (rl < nown) st_pub(Atw + (size_t)(row0 + rl) * ldk + np + k, acc);
            } else {
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) cfma(acc, Dl[k * 4 + kk], GR[(kk * 8 + j) * grld + rl]);
              if (rl >= nown) acc = cmake(0.0, 0.0);
              VBs[k * rpt + rl] = acc;
              Bown[rl * ldo + np + k] = acc;
              if (rl < nown) st_pub(Bmw + (size_t)(row0 + rl) * ldk + np + k, acc);
            }
          }
          if (j + 1 < nb) {
            __syncwarp();   // named barriers are .aligned: the whole warp executes them together
            asm volatile("bar.sync 4, 128;" ::: "memory");                // UPD: U, V, UA, VB complete
            // (b) the whole window follows the update exactly: Sw += U V, GC += UA V, GR += U VB, one DMMA k-step (k = 4) per tile
            const int rt = rpt >> 3;
            const int nsw = sym ? 8 : 16, ngc = rt * 4, ntl = nsw + 2 * ngc;
            const int r = lane >> 2, kq = lane & 3, c2 = 2 * kq;
            for (int t = warp - 4; t < ntl; t += 4) {
              const cplx* X; const cplx* Y; cplx* D; int ldy, ldd; bool mirror = false;
              if (t < nsw) { const int tm = t >> 2, tnn = t & 3; X = Us + tm * 8 * 4; Y = Vs + tnn * 8; ldy = LU_W; D = Sw + tm * 8 * LU_SWLD + tnn * 8; ldd = LU_SWLD; mirror = sym; }
              else if (t < nsw + ngc) { const int qq = t - nsw, tm = qq >> 2, tnn = qq & 3; X = UAs + tm * 8 * 4; Y = Vs + tnn * 8; ldy = LU_W; D = GC + tm * 8 * LU_SWLD + tnn * 8; ldd = LU_SWLD; }
              else { const int qq = t - nsw - ngc, tm = qq & 3, tnn = qq >> 2; X = Us + tm * 8 * 4; Y = VBs + tnn * 8; ldy = rpt; D = GR + tm * 8 * grld + tnn * 8; ldd = grld; }
              const cplx x = X[r * 4 + kq], y = Y[kq * ldy + r];
              cplx* q0 = D + r * ldd + c2;
              cplx v0 = q0[0], v1 = q0[1];
              dmma884(v0.x, v1.x, x.x, y.x);
              dmma884(v0.y, v1.y, x.x, y.y);
              dmma884(v0.x, v1.x, -x.y, y.y);
              dmma884(v0.y, v1.y, x.y, y.x);
              q0[0] = v0; q0[1] = v1;
              if (mirror) {   // t < 8: window rows tm*8 + r < 16
                const int ww = (t >> 2) * 8 + r, wc = (t & 3) * 8 + c2;        // wc, wc + 1 lie in the same half
                const bool left = wc < 16;
                cplx* m0 = Sw + (ww + 16) * LU_SWLD + (left ? wc + 16 : wc - 16);
                m0[0] = left ? cmake(v0.x, -v0.y) : cmake(-v0.x, v0.y);
                m0[1] = left ? cmake(v1.x, -v1.y) : cmake(-v1.x, v1.y);
              }
            }
          }
        }
        __syncwarp();   // named barriers are .aligned: the whole warp executes them together
        asm volatile("bar.arrive 3, 160;" ::: "memory");                  // DONE(j)
      }
      // bookkeeping every role keeps for itself
      if (accepted) { kc++; np += 4; }
      if (warp <= 3) { off += (scn == 1) ? 3 : 4; s_cur = scn; }
    }
    __syncwarp();   // named barriers are .aligned: the whole warp executes them together
    if (warp == 0) asm volatile("bar.sync 3, 160;" ::: "memory");          // DONE(nb-1)
    __syncthreads();
    long long ts2 = PROF ? clock64() : 0;
    const int i = s0 + nb - 1;                                             // last site of the block
    // ================= flush at a block end: G += A B over the pending 4*kc columns =================
    const bool do_flush = kc > 0 && (kc > kmax || i == N - 1);
    if (do_flush) {
      long long tf[6] = {0, 0, 0, 0, 0, 0};
      if (PROF) tf[0] = clock64();
      bar_target += gridDim.x; grid_barrier_mono(bar_ctr, bar_target);     // also orders every thread's stage-2 smem traffic before the staging reuse
      if (PROF) tf[1] = clock64();
      const cplx* Atb = Atw;
      const cplx* Bmb = Bmw;
      const int K = 4 * kc;
      const int lo = lane >> 2, lk = lane & 3;
      const int wm = warp & 1, wn = warp >> 1;           // 2 x 4 warps, warp tile 32 x 16
      // Antiunitary flavour symmetry G = [[A, B], [-conj(B), conj(A)]]: every accepted update preserves it, so the flush computes
      // the upper half of G only and writes the lower half as its mirror image - half the DMMAs.
      const int mrows = sym ? hN : n;
      const int tiles_m = (n + 63) / 64, tiles_r = (mrows + 63) / 64, ntiles = tiles_r * tiles_m;
      const int nch = (K + LU_FCH - 1) / LU_FCH;
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int tm0 = (t % tiles_r) * 64, tn0 = (t / tiles_r) * 64;
        auto issue = [&](int ch) {
          const int k0 = ch * LU_FCH;
          cplx* fa = FA + (ch % LU_FST) * (64 * LU_FLD);
          cplx* fb = FB + (ch % LU_FST) * (64 * LU_FLD);
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int e = tid + 256 * u, rr = e >> 4, kk = e & 15;
            const bool kok = (k0 + kk) < K;
            { const bool ok = kok && tm0 + rr < n; cp_async16(fa + rr * LU_FLD + kk, ok ? Atb + (size_t)(tm0 + rr) * ldk + k0 + kk : Atb, ok); }
            { const bool ok = kok && tn0 + rr < n; cp_async16(fb + rr * LU_FLD + kk, ok ? Bmb + (size_t)(tn0 + rr) * ldk + k0 + kk : Bmb, ok); }
          }
          asm volatile("cp.async.commit_group;" ::: "memory");
        };
        issue(0);
        if (nch > 1) issue(1);
        // 3M complex product: S1 = Ar Br, S2 = Ai Bi, S3 = (Ar + Ai)(Br + Bi);  Re = S1 - S2,  Im = S3 - S1 - S2
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
          __syncthreads();                                 // chunk ch landed for everybody; everybody is done with chunk ch-1
          if (ch + 2 < nch) issue(ch + 2);                 // into the buffer chunk ch-1 used
          const cplx* fa = FA + (ch % LU_FST) * (64 * LU_FLD);
          const cplx* fb = FB + (ch % LU_FST) * (64 * LU_FLD);
#pragma unroll
          for (int ks = 0; ks < LU_FCH / 4; ++ks) {
            if (ch * LU_FCH + ks * 4 >= K) break;            // the zero-padded tail of the last chunk
            cplx av[4], bv[2];
#pragma unroll
            for (int x = 0; x < 4; ++x) av[x] = fa[(wm * 32 + x * 8 + lo) * LU_FLD + ks * 4 + lk];
#pragma unroll
            for (int y = 0; y < 2; ++y) bv[y] = fb[(wn * 16 + y * 8 + lo) * LU_FLD + ks * 4 + lk];
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
        __syncthreads();                                   // staging ring free for the next tile's loads
        if (PROF) tf[2] = clock64();
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
      if (PROF) tf[3] = clock64();
      bar_target += gridDim.x; grid_barrier_mono(bar_ctr, bar_target);
      if (PROF) tf[4] = clock64();
      {   // re-arm my rows of the buffer just consumed (nobody reads it again before the flush after next): thread = (row, 32 threads per row)
        cplx* Atr = a.At + batch * bufstride;
        cplx* Bmr = a.Bm + batch * bufstride;
        for (int rl = warp; rl < nown; rl += 8)
          for (int p = lane; p < K; p += 32) {
            st_sent(Atr + (size_t)(row0 + rl) * ldk + p);
            st_sent(Bmr + (size_t)(row0 + rl) * ldk + p);
          }
      }
      if (PROF) { tf[5] = clock64(); for (int q = 0; q < 5; ++q) pf[q] += tf[q + 1] - tf[q]; }
      batch ^= 1;
      kc = 0;
      np = 0;
      if (PROF) pr[5]++;
    }
    if (PROF) { const long long ts3 = clock64(); pr[2] += ts2 - ts0; pr[4] += ts3 - ts2; }
  }
  if (PROF && a.prof != nullptr && blockIdx.x == 0 && tid == 0) {
    // [0] total, [1] gather, [2] form, [3] site loops of the blocks, [4] -, [5] flush, [6] #flushes, [7] accepts
    a.prof[0] = clock64() - t_begin; a.prof[1] = pr[0]; a.prof[2] = pr[1]; a.prof[3] = pr[2]; a.prof[4] = pr[3]; a.prof[5] = pr[4];
    a.prof[6] = pr[5]; a.prof[7] = nacc;
    for (int q = 0; q < 5; ++q) a.prof[8 + q] = pf[q];
  }

  if (blockIdx.x == 0 && tid == 0) {
    *a.pos = pos0 + off;
    *a.accepted += nacc;
    *a.dS += dS_sum;
    if (s_exh) a.flags[0] = 1;
    if (nonreal) a.flags[1] += nonreal;
  }
}

size_t lu_block_smem(const LUArgs& a) { return blk_layout(a.rpc, a.kmax, a.nsites).total; }

int launch_lu_block(cudaStream_t st, const LUArgs& a, int grid) {
  const size_t smem = lu_block_smem(a);
  static SmemMemo memo, memo_prof;
  size_t smem_lim = 0;
  if (ensure_max_dynamic_smem(lu_block_kernel<false>, memo, &smem_lim) || ensure_max_dynamic_smem(lu_block_kernel<true>, memo_prof, &smem_lim)) return -1;
  if (smem > smem_lim) { snprintf(g_errbuf, sizeof(g_errbuf), "local_updates (block kernel): shared memory %zu > %zu", smem, smem_lim); return -1; }
  if (a.rpc > 16 || a.kmax > 16 || a.kmax < 1) { snprintf(g_errbuf, sizeof(g_errbuf), "local_updates (block kernel): rpc %d / kmax %d out of range", a.rpc, a.kmax); return -1; }
  LUArgs args = a;
  void* params[] = {&args};
  const void* kern = a.prof ? (const void*)lu_block_kernel<true> : (const void*)lu_block_kernel<false>;
  CUDA_TRY(cudaLaunchCooperativeKernel(kern, dim3(grid), dim3(256), params, smem, st));
  g_launches++;
  return 0;
}
