#include "zgemm.cuh"
#include <cstdlib>

#ifndef ZGEMM_KT
#define ZGEMM_KT 16   // k-tile of the large-tile variants (two shared-memory stages)
#endif

// Operand tiles live in shared memory in *fragment order*: for every (k-step of 4, tile of 8 rows/cols)
// the 32 lanes' elements are contiguous, element `lane` being X[o = lane/4][k = lane%4] as (re,im).
// A fragment load is then one conflict-free LDS.128 per lane, and the global->shared copy reads
// 128 B (op N on A / op T,C on B) or 64 B (the other cases) contiguous segments.
//
// One complex 8x8x4 tile product = 4 real DMMAs:  Cr += Ar*Br - Ai*Bi,  Ci += Ar*Bi + Ai*Br.
// M3 = true: "3M" complex product, S1 = Ar Br, S2 = Ai Bi, S3 = (Ar + Ai)(Br + Bi), Re = S1 - S2, Im = S3 - S1 - S2: three real
// DMMAs per tile product (the kernel is DMMA-bound), at the price of a third accumulator set (hence the smaller warp tile).
template <int WTM, int WTN, int NWM, int NWN, int KT, bool M3>
__global__ void __launch_bounds__(NWM * NWN * 32)
zgemm_kernel(int M, int N, int K, cplx alpha, const cplx* __restrict__ A, int lda, int opA,
             const cplx* __restrict__ B, int ldb, int opB, cplx beta, cplx* __restrict__ C, int ldc) {
  constexpr int BM = NWM * WTM * 8, BN = NWN * WTN * 8, KS = KT / 4;
  constexpr int NW = NWM * NWN;
  constexpr int FA = (BM / 8) * KS, FB = (BN / 8) * KS;       // fragments per stage
  constexpr int LA = (FA + NW - 1) / NW, LB = (FB + NW - 1) / NW;
  // two stages: the next k-tile is written while the current one is read, one barrier per k-tile
  extern __shared__ __align__(16) unsigned char zsm[];
  cplx* As = reinterpret_cast<cplx*>(zsm);                    // [2][FA*32]
  cplx* Bs = As + 2 * FA * 32;                                // [2][FB*32]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wm = warp % NWM, wn = warp / NWM;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int lo = lane >> 2, lk = lane & 3;

  // M3: cr = S1, ci = S2, c3 = S3
  double cr[WTM][WTN][2], ci[WTM][WTN][2], c3[M3 ? WTM : 1][M3 ? WTN : 1][2];
#pragma unroll
  for (int i = 0; i < WTM; ++i)
#pragma unroll
    for (int j = 0; j < WTN; ++j) {
      cr[i][j][0] = cr[i][j][1] = ci[i][j][0] = ci[i][j][1] = 0.0;
      if (M3) c3[M3 ? i : 0][M3 ? j : 0][0] = c3[M3 ? i : 0][M3 ? j : 0][1] = 0.0;
    }

  cplx ra[LA], rb[LB];
  auto gload = [&](int k0) {
#pragma unroll
    for (int t = 0; t < LA; ++t) {
      int f = warp + t * NW;
      cplx v = cmake(0.0, 0.0);
      if (f < FA) {
        int kk = f / (BM / 8), rt = f % (BM / 8);
        int row = m0 + rt * 8 + lo, k = k0 + kk * 4 + lk;
        if (row < M && k < K) {
          if (opA == OP_N) v = A[(size_t)k * lda + row];
          else { v = A[(size_t)row * lda + k]; if (opA == OP_C) v.y = -v.y; }
        }
      }
      ra[t] = v;
    }
#pragma unroll
    for (int t = 0; t < LB; ++t) {
      int f = warp + t * NW;
      cplx v = cmake(0.0, 0.0);
      if (f < FB) {
        int kk = f / (BN / 8), ct = f % (BN / 8);
        int col = n0 + ct * 8 + lo, k = k0 + kk * 4 + lk;
        if (col < N && k < K) {
          if (opB == OP_N) v = B[(size_t)col * ldb + k];
          else { v = B[(size_t)k * ldb + col]; if (opB == OP_C) v.y = -v.y; }
        }
      }
      rb[t] = v;
    }
  };
  auto sstore = [&](int stage) {
#pragma unroll
    for (int t = 0; t < LA; ++t) { int f = warp + t * NW; if (f < FA) As[(stage * FA + f) * 32 + lane] = ra[t]; }
#pragma unroll
    for (int t = 0; t < LB; ++t) { int f = warp + t * NW; if (f < FB) Bs[(stage * FB + f) * 32 + lane] = rb[t]; }
  };

  gload(0);
  sstore(0);
  __syncthreads();
  int stage = 0;
  for (int k0 = 0; k0 < K; k0 += KT) {
    const bool more = k0 + KT < K;
    if (more) gload(k0 + KT);
    const cplx* as = As + stage * FA * 32;
    const cplx* bs = Bs + stage * FB * 32;
#pragma unroll
    for (int kk = 0; kk < KS; ++kk) {
      cplx a[WTM], b[WTN];
#pragma unroll
      for (int i = 0; i < WTM; ++i) a[i] = as[(kk * (BM / 8) + wm * WTM + i) * 32 + lane];
#pragma unroll
      for (int j = 0; j < WTN; ++j) b[j] = bs[(kk * (BN / 8) + wn * WTN + j) * 32 + lane];
      if (M3) {
        double as3[WTM], bs3[WTN];
#pragma unroll
        for (int i = 0; i < WTM; ++i) as3[i] = a[i].x + a[i].y;
#pragma unroll
        for (int j = 0; j < WTN; ++j) bs3[j] = b[j].x + b[j].y;
#pragma unroll
        for (int i = 0; i < WTM; ++i)
#pragma unroll
          for (int j = 0; j < WTN; ++j) {
            dmma884(cr[i][j][0], cr[i][j][1], a[i].x, b[j].x);
            dmma884(ci[i][j][0], ci[i][j][1], a[i].y, b[j].y);
            dmma884(c3[M3 ? i : 0][M3 ? j : 0][0], c3[M3 ? i : 0][M3 ? j : 0][1], as3[i], bs3[j]);
          }
      } else {
#pragma unroll
        for (int i = 0; i < WTM; ++i)
#pragma unroll
          for (int j = 0; j < WTN; ++j) {
            dmma884(cr[i][j][0], cr[i][j][1], a[i].x, b[j].x);
            dmma884(cr[i][j][0], cr[i][j][1], -a[i].y, b[j].y);
            dmma884(ci[i][j][0], ci[i][j][1], a[i].x, b[j].y);
            dmma884(ci[i][j][0], ci[i][j][1], a[i].y, b[j].x);
          }
      }
    }
    if (more) sstore(stage ^ 1);   // the other stage was last read before the previous barrier
    __syncthreads();
    stage ^= 1;
  }

  const bool use_c = (beta.x != 0.0 || beta.y != 0.0);
#pragma unroll
  for (int i = 0; i < WTM; ++i)
#pragma unroll
    for (int j = 0; j < WTN; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        int row = m0 + (wm * WTM + i) * 8 + lo;
        int col = n0 + (wn * WTN + j) * 8 + 2 * lk + e;
        if (row < M && col < N) {
          const cplx prod = M3 ? cmake(cr[i][j][e] - ci[i][j][e], c3[M3 ? i : 0][M3 ? j : 0][e] - cr[i][j][e] - ci[i][j][e])
                               : cmake(cr[i][j][e], ci[i][j][e]);
          cplx acc = cmul(alpha, prod);
          cplx* p = C + (size_t)col * ldc + row;
          if (use_c) cfma(acc, beta, *p);
          *p = acc;
        }
      }
}

template <int WTM, int WTN, int NWM, int NWN, int KT, bool M3 = false>
static int zgemm_launch(cudaStream_t stream, int opA, int opB, int M, int N, int K, cplx alpha, const cplx* A, int lda,
                        const cplx* B, int ldb, cplx beta, cplx* C, int ldc) {
  constexpr int BM = NWM * WTM * 8, BN = NWN * WTN * 8;
  constexpr size_t smem = sizeof(cplx) * 2 * 32 * ((BM / 8) * (KT / 4) + (BN / 8) * (KT / 4));
  static SmemMemo memo;
  if (ensure_dynamic_smem(zgemm_kernel<WTM, WTN, NWM, NWN, KT, M3>, memo, smem)) return -1;
  dim3 g((M + BM - 1) / BM, (N + BN - 1) / BN);
  zgemm_kernel<WTM, WTN, NWM, NWN, KT, M3><<<g, NWM * NWN * 32, smem, stream>>>(M, N, K, alpha, A, lda, opA, B, ldb, opB, beta, C, ldc);
  return 0;
}

int zgemm(cudaStream_t stream, int opA, int opB, int M, int N, int K, cplx alpha, const cplx* A, int lda,
          const cplx* B, int ldb, cplx beta, cplx* C, int ldc, int num_sms) {
  if (M <= 0 || N <= 0) return 0;
  if (opA == OP_J || opB == OP_J) { snprintf(g_errbuf, sizeof(g_errbuf), "zgemm: OP_J unsupported"); return -1; }
  auto tiles = [&](int bm, int bn) { return (long)((M + bm - 1) / bm) * ((N + bn - 1) / bn); };
  const long want = (long)num_sms * 3 / 4;
  int rc;
  static const int use3m = []() { const char* e = getenv("DQMC_ZGEMM_3M"); return e ? atoi(e) : 3; }();
  if (use3m == 1 && tiles(64, 64) >= want) rc = zgemm_launch<4, 2, 2, 4, ZGEMM_KT, true>(stream, opA, opB, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc);
  else if (use3m == 2 && tiles(128, 32) >= want) rc = zgemm_launch<4, 2, 4, 2, ZGEMM_KT, true>(stream, opA, opB, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc);
  else if (use3m == 3 && tiles(64, 64) >= want) rc = zgemm_launch<4, 2, 2, 4, 32, true>(stream, opA, opB, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc);
  else if (tiles(128, 64) >= want) rc = zgemm_launch<4, 4, 4, 2, ZGEMM_KT>(stream, opA, opB, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc);
  else if (tiles(64, 64) >= want) rc = zgemm_launch<4, 2, 2, 4, ZGEMM_KT>(stream, opA, opB, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc);
  else rc = zgemm_launch<2, 2, 2, 2, 8>(stream, opA, opB, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc);
  if (rc) return rc;
  CUDA_TRY(cudaGetLastError());
  g_launches++;
  return 0;
}
