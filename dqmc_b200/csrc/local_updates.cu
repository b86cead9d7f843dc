#include "local_updates.cuh"

// 4x4 complex e^{-power*dtau*V(op)} (interactions.jl:102-141), element (r,c).
__device__ __forceinline__ cplx evop_elem(int r, int c, double C, cplx S, double R) {
  // [C S 0 R; cS C -R 0; 0 -R C cS; R 0 S C]
  if (r == c) return cmake(C, 0.0);
  const int code = r * 4 + c;
  switch (code) {
    case 1: return S;                 // (0,1)
    case 3: return cmake(R, 0.0);     // (0,3)
    case 4: return cconj(S);          // (1,0)
    case 6: return cmake(-R, 0.0);    // (1,2)
    case 9: return cmake(-R, 0.0);    // (2,1)
    case 11: return cconj(S);         // (2,3)
    case 12: return cmake(R, 0.0);    // (3,0)
    case 14: return S;                // (3,2)
    default: return cmake(0.0, 0.0);
  }
}

__device__ __forceinline__ cplx det3(cplx a, cplx b, cplx c, cplx d, cplx e, cplx f, cplx g, cplx h, cplx i) {
  // | a b c ; d e f ; g h i |
  cplx t1 = csub(cmul(e, i), cmul(f, h));
  cplx t2 = csub(cmul(d, i), cmul(f, g));
  cplx t3 = csub(cmul(d, h), cmul(e, g));
  return cadd(csub(cmul(a, t1), cmul(b, t2)), cmul(c, t3));
}

// All CTAs run the site loop in lock step and take every accept/reject decision redundantly from the same
// data, so a rejected proposal needs no communication at all; an accepted one needs a single grid barrier
// (each CTA publishes its rows of the new A columns / its columns of the new B rows).
__global__ void __launch_bounds__(256) local_updates_kernel(LUArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int n = a.n, N = a.nsites, ldk = 4 * a.kmax, rpc = a.rpc;
  cplx* Aown = reinterpret_cast<cplx*>(smem_raw);          // [rpc][ldk]
  cplx* Bown = Aown + (size_t)rpc * ldk;                    // [rpc][ldk]
  cplx* As4 = Bown + (size_t)rpc * ldk;                     // [4][ldk]  rows i+kN of A
  cplx* Bs4 = As4 + 4 * ldk;                                // [4][ldk]  cols i+kN of B
  cplx* gcol = Bs4 + 4 * ldk;                               // [rpc][4]  G[r, i+kN] for my rows
  cplx* grow = gcol + (size_t)rpc * 4;                      // [rpc][4]  G[i+kN, c] for my cols
  double* fs = reinterpret_cast<double*>(grow + (size_t)rpc * 4);   // [3*N] field of this slice
  __shared__ cplx g4[16], E1[16], E2[16], Dl[16], Mm[16], Cof[16], Minv[16];
  __shared__ int s_accept;
  __shared__ double s_newop[3];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int row0 = blockIdx.x * rpc;
  const int nown = max(0, min(n, row0 + rpc) - row0);
  const int sl = a.slice;
  const int sl_later = (sl + 1) % a.nslices, sl_earlier = (sl + a.nslices - 1) % a.nslices;

  for (int e = tid; e < 3 * N; e += blockDim.x) fs[e] = a.hs[(size_t)3 * N * sl + e];
  long long pos = *a.pos;
  long long nacc = 0;
  double dS_sum = 0.0;
  int kc = 0, nonreal = 0, exhausted = 0;
  __syncthreads();

  for (int i = 0; i < N; ++i) {
    const int np = 4 * kc;
    // ---- gather what the proposal (and a possible accept) needs
    for (int e = tid; e < 4 * np; e += blockDim.x) {
      int k = e / np, p = e % np;
      As4[k * ldk + p] = a.At[(size_t)(i + k * N) * ldk + p];
      Bs4[k * ldk + p] = a.Bm[(size_t)(i + k * N) * ldk + p];
    }
    if (tid < 16) g4[tid] = a.G[(size_t)(i + (tid >> 2) * N) * n + i + (tid & 3) * N];   // g4[r + 4c] = G[i+rN, i+cN]
    for (int e = tid; e < nown * 4; e += blockDim.x) {
      int rl = e >> 2, k = e & 3;
      gcol[e] = a.G[(size_t)(i + k * N) * n + row0 + rl];
      grow[e] = a.G[(size_t)(row0 + rl) * n + i + k * N];
    }
    __syncthreads();
    // effective 4x4 block: g += A[i+rN, :] B[:, i+cN]
    {
      const int o = tid >> 4, q = tid & 15;          // 16 outputs x 16 partial lanes
      const int r = o & 3, c = o >> 2;
      cplx acc = cmake(0.0, 0.0);
      for (int p = q; p < np; p += 16) cfma(acc, As4[r * ldk + p], Bs4[c * ldk + p]);
#pragma unroll
      for (int s = 8; s > 0; s >>= 1) {
        acc.x += __shfl_xor_sync(0xffffffffu, acc.x, s);
        acc.y += __shfl_xor_sync(0xffffffffu, acc.y, s);
      }
      __syncthreads();
      if (q == 0) g4[o] = cadd(g4[o], acc);
    }
    __syncthreads();
    // ---- decision (warp 0; lanes 0..15 own one matrix element each, scalars are computed redundantly)
    if (warp == 0) {
      const int r = lane & 3, c = (lane >> 2) & 3;
      double u0 = 0.0, u1 = 0.0, u2 = 0.0, u3 = 0.0;
      if (pos + 4 <= a.nunif) { u0 = a.unif[pos]; u1 = a.unif[pos + 1]; u2 = a.unif[pos + 2]; u3 = a.unif[pos + 3]; }
      else exhausted = 1;
      const double o1 = fs[3 * i], o2 = fs[3 * i + 1], o3 = fs[3 * i + 2];
      // randuniform (dqmc_framework.jl:628): -b + 2*b*rand(), no FMA contraction so the field stays bit-identical
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
        const double* he = a.hs + 3 * ((size_t)i + (size_t)N * sl_earlier);
        const double* hl = a.hs + 3 * ((size_t)i + (size_t)N * sl_later);
        const double t1 = hl[0] + he[0], t2 = hl[1] + he[1], t3 = hl[2] + he[2];
        double s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
        for (int nb = 0; nb < 4; ++nb) {
          const int j = a.nbr[4 * i + nb];
          s1 += fs[3 * j]; s2 += fs[3 * j + 1]; s3 += fs[3 * j + 2];
        }
        dS = a.inv_dtau_c2 * (sq_diff - (t1 * d1 + t2 * d2 + t3 * d3));
        dS += 0.5 * a.dtau * (4.0 * sq_diff - 2.0 * (s1 * d1 + s2 * d2 + s3 * d3));
        dS += a.dtau * (0.5 * a.r * sq_diff + 0.25 * a.u * pow4_diff);
      } else {
        dS = a.dtau * (0.5 * a.r * sq_diff);
      }
      const double e_dS = exp(-dS);
      // interaction_matrix_exp_op!: old with power -1, new with power +1
      const double on = sqrt(osq), nn = sqrt(nsq);
      const double sh1 = -sinh(a.lam_dtau * on) / on, C1 = cosh(a.lam_dtau * on);
      const double sh2 = sinh(a.lam_dtau * nn) / nn, C2 = cosh(a.lam_dtau * nn);
      const cplx S1 = cmake(-o1 * sh1, o2 * sh1), S2 = cmake(-n1 * sh2, n2 * sh2);
      const double R1 = -o3 * sh1, R2 = -n3 * sh2;
      if (lane < 16) { E1[r * 4 + c] = evop_elem(r, c, C1, S1, R1); E2[r * 4 + c] = evop_elem(r, c, C2, S2, R2); }
      __syncwarp();
      if (lane < 16) {   // delta = E1*E2 - 1
        cplx acc = cmake(r == c ? -1.0 : 0.0, 0.0);
#pragma unroll
        for (int k = 0; k < 4; ++k) cfma(acc, E1[r * 4 + k], E2[k * 4 + c]);
        Dl[r * 4 + c] = acc;
      }
      __syncwarp();
      if (lane < 16) {   // M = 1 + delta*(1 - g)
        cplx acc = cmake(r == c ? 1.0 : 0.0, 0.0);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          cplx gk = g4[k + 4 * c];
          cplx omg = cmake((k == c ? 1.0 : 0.0) - gk.x, -gk.y);
          cfma(acc, Dl[r * 4 + k], omg);
        }
        Mm[r * 4 + c] = acc;
      }
      __syncwarp();
      if (lane < 16) {   // cofactor (r,c)
        int rr[3], cc[3];
        for (int t = 0, q = 0; t < 4; ++t) if (t != r) rr[q++] = t;
        for (int t = 0, q = 0; t < 4; ++t) if (t != c) cc[q++] = t;
        cplx d = det3(Mm[rr[0] * 4 + cc[0]], Mm[rr[0] * 4 + cc[1]], Mm[rr[0] * 4 + cc[2]],
                      Mm[rr[1] * 4 + cc[0]], Mm[rr[1] * 4 + cc[1]], Mm[rr[1] * 4 + cc[2]],
                      Mm[rr[2] * 4 + cc[0]], Mm[rr[2] * 4 + cc[1]], Mm[rr[2] * 4 + cc[2]]);
        Cof[r * 4 + c] = ((r + c) & 1) ? cneg(d) : d;
      }
      __syncwarp();
      cplx det = cmake(0.0, 0.0);
#pragma unroll
      for (int k = 0; k < 4; ++k) cfma(det, Mm[k], Cof[k]);     // expansion along row 0
      if (lane < 16) Minv[r * 4 + c] = cdiv(Cof[c * 4 + r], det);
      const double p_acc = e_dS * det.x;
      if (fabs(det.y / det.x) > 1e-4) nonreal++;
      int acc_flag;
      if (p_acc > 1.0) { acc_flag = 1; pos += 3; }
      else { acc_flag = (u3 < p_acc) ? 1 : 0; pos += 4; }
      if (acc_flag) { nacc++; dS_sum += -log(e_dS); }
      if (lane == 0) { s_accept = acc_flag; s_newop[0] = n1; s_newop[1] = n2; s_newop[2] = n3; }
    }
    __syncthreads();
    if (s_accept) {
      if (tid < 3) {
        fs[3 * i + tid] = s_newop[tid];
        if (blockIdx.x == 0) a.hs[(size_t)3 * N * sl + 3 * i + tid] = s_newop[tid];
      }
      // A_new[r,:] = (G_eff[r, i+kN] - delta_{r,i+kN}) Minv ;  B_new[:,c] = delta * G_eff[i+kN, c]
      // 4 threads per (row, k) dot product over the pending columns.
      {
        const int task = tid >> 2, q = tid & 3;
        const int half = nown * 4;                    // tasks [0,half): A side, [half,2*half): B side
        for (int t0 = 0; t0 < 2 * half; t0 += 64) {
          const int t = t0 + task;
          cplx acc = cmake(0.0, 0.0);
          if (t < 2 * half) {
            const bool isB = t >= half;
            const int tt = isB ? t - half : t;
            const int rl = tt >> 2, k = tt & 3;
            const cplx* own = (isB ? Bown : Aown) + (size_t)rl * ldk;
            const cplx* site = (isB ? As4 : Bs4) + k * ldk;
            for (int p = q; p < np; p += 4) cfma(acc, own[p], site[p]);
          }
          acc.x += __shfl_xor_sync(0xffffffffu, acc.x, 1); acc.y += __shfl_xor_sync(0xffffffffu, acc.y, 1);
          acc.x += __shfl_xor_sync(0xffffffffu, acc.x, 2); acc.y += __shfl_xor_sync(0xffffffffu, acc.y, 2);
          if (t < 2 * half && q == 0) {
            const bool isB = t >= half;
            const int tt = isB ? t - half : t;
            const int rl = tt >> 2, k = tt & 3;
            if (!isB) {
              cplx v = cadd(gcol[tt], acc);
              if (row0 + rl == i + k * N) v.x -= 1.0;
              gcol[tt] = v;
            } else {
              grow[tt] = cadd(grow[tt], acc);
            }
          }
        }
      }
      __syncthreads();
      for (int e = tid; e < nown * 8; e += blockDim.x) {
        const bool isB = e >= nown * 4;
        const int tt = isB ? e - nown * 4 : e;
        const int rl = tt >> 2, k = tt & 3;
        cplx acc = cmake(0.0, 0.0);
        if (!isB) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) cfma(acc, gcol[rl * 4 + kk], Minv[kk * 4 + k]);
          Aown[(size_t)rl * ldk + np + k] = acc;
          a.At[(size_t)(row0 + rl) * ldk + np + k] = acc;
        } else {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) cfma(acc, Dl[k * 4 + kk], grow[rl * 4 + kk]);
          Bown[(size_t)rl * ldk + np + k] = acc;
          a.Bm[(size_t)(row0 + rl) * ldk + np + k] = acc;
        }
      }
      kc++;
      grid_barrier(a.bar, gridDim.x);
    }
    // ---- flush: G += A B with the pending 4*kc columns (DMMA, operands straight from L2 in fragment order)
    if (kc == a.kmax || (i == N - 1 && kc > 0)) {
      const int K = 4 * kc;
      const int lo = lane >> 2, lk = lane & 3;
      const int wm = warp & 1, wn = warp >> 1;           // 2 x 4 warps, warp tile 32 x 16
      const int tiles_m = (n + 63) / 64, ntiles = tiles_m * tiles_m;
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int m0 = (t % tiles_m) * 64 + wm * 32, n0 = (t / tiles_m) * 64 + wn * 16;
        double cr[4][2][2], ci[4][2][2];
#pragma unroll
        for (int x = 0; x < 4; ++x)
#pragma unroll
          for (int y = 0; y < 2; ++y) cr[x][y][0] = cr[x][y][1] = ci[x][y][0] = ci[x][y][1] = 0.0;
        for (int k0 = 0; k0 < K; k0 += 4) {
          cplx av[4], bv[2];
#pragma unroll
          for (int x = 0; x < 4; ++x) {
            const int row = m0 + x * 8 + lo;
            av[x] = row < n ? a.At[(size_t)row * ldk + k0 + lk] : cmake(0.0, 0.0);
          }
#pragma unroll
          for (int y = 0; y < 2; ++y) {
            const int col = n0 + y * 8 + lo;
            bv[y] = col < n ? a.Bm[(size_t)col * ldk + k0 + lk] : cmake(0.0, 0.0);
          }
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
#pragma unroll
        for (int x = 0; x < 4; ++x)
#pragma unroll
          for (int y = 0; y < 2; ++y)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int row = m0 + x * 8 + lo, col = n0 + y * 8 + 2 * lk + e;
              if (row < n && col < n) {
                cplx* p = a.G + (size_t)col * n + row;
                cplx v = *p;
                v.x += cr[x][y][e]; v.y += ci[x][y][e];
                *p = v;
              }
            }
      }
      kc = 0;
      grid_barrier(a.bar, gridDim.x);
    }
  }

  if (blockIdx.x == 0 && tid == 0) {
    *a.pos = pos;
    *a.accepted += nacc;
    *a.dS += dS_sum;
    if (exhausted) a.flags[0] = 1;
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

int launch_local_updates(cudaStream_t st, const LUArgs& a, int grid) {
  const int ldk = 4 * a.kmax;
  const size_t smem = sizeof(cplx) * ((size_t)2 * a.rpc * ldk + 8 * ldk + 8 * a.rpc) + sizeof(double) * 3 * a.nsites;
  static size_t smem_lim = 0;
  if (smem_lim == 0 && set_max_dynamic_smem(local_updates_kernel, &smem_lim)) return -1;
  if (smem > smem_lim) { snprintf(g_errbuf, sizeof(g_errbuf), "local_updates: shared memory %zu > %zu", smem, smem_lim); return -1; }
  LUArgs args = a;
  void* params[] = {&args};
  CUDA_TRY(cudaLaunchCooperativeKernel((const void*)local_updates_kernel, dim3(grid), dim3(256), params, smem, st));
  g_launches++;
  return 0;
}
