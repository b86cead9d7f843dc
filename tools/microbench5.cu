// Latency of the local-update DECISION role in isolation (one warp; same code as local_updates.cu stage 1, warp 0).
#include <cstdio>
#include <cuda_runtime.h>
#include "../dqmc_b200/csrc/common.cuh"
char g_errbuf[512]; long long g_launches = 0;
__device__ __forceinline__ cplx det3(cplx a, cplx b, cplx c, cplx d, cplx e, cplx f, cplx g, cplx h, cplx i) {
  cplx t1 = csub(cmul(e, i), cmul(f, h));
  cplx t2 = csub(cmul(d, i), cmul(f, g));
  cplx t3 = csub(cmul(d, h), cmul(e, g));
  return cadd(csub(cmul(a, t1), cmul(b, t2)), cmul(c, t3));
}
__global__ void decision_probe(double* out, int iters, int nwarps_spin) {
  __shared__ cplx g4e[16], D[16], Mm[16], Cof[16], Minv[16];
  __shared__ int s_accept;
  __shared__ volatile int stop;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x < 16) { g4e[threadIdx.x] = cmake(0.1 * threadIdx.x, 0.01); D[threadIdx.x] = cmake(0.01 * threadIdx.x, -0.02); }
  if (threadIdx.x == 0) stop = 0;
  __syncthreads();
  if (warp > 0) {   // optional interference: warps spinning on shared memory / doing FP64
    double x = lane;
    while (!stop) { x = fma(x, 1.0000001, 1e-9); }
    out[64 + threadIdx.x] = x;
    return;
  }
  long long t0 = clock64();
  double e_dS = 0.9;
  for (int it = 0; it < iters; ++it) {
    const int r = (lane >> 2) & 3, c = lane & 3;
    if (lane < 16) {
      cplx g0 = g4e[0 + 4 * c], g1 = g4e[1 + 4 * c], g2 = g4e[2 + 4 * c], g3 = g4e[3 + 4 * c];
      g0 = cmake((c == 0 ? 1.0 : 0.0) - g0.x, -g0.y); g1 = cmake((c == 1 ? 1.0 : 0.0) - g1.x, -g1.y);
      g2 = cmake((c == 2 ? 1.0 : 0.0) - g2.x, -g2.y); g3 = cmake((c == 3 ? 1.0 : 0.0) - g3.x, -g3.y);
      const cplx t0_ = cmul(D[r * 4 + 0], g0), t1 = cmul(D[r * 4 + 1], g1), t2 = cmul(D[r * 4 + 2], g2), t3 = cmul(D[r * 4 + 3], g3);
      cplx m = cadd(cadd(t0_, t1), cadd(t2, t3));
      if (r == c) m.x += 1.0;
      Mm[r * 4 + c] = m;
    }
    __syncwarp();
    if (lane < 16) {
      const int r0 = (r == 0) ? 1 : 0, r1 = (r <= 1) ? 2 : 1, r2 = (r <= 2) ? 3 : 2;
      const int c0 = (c == 0) ? 1 : 0, c1 = (c <= 1) ? 2 : 1, c2 = (c <= 2) ? 3 : 2;
      cplx d = det3(Mm[r0 * 4 + c0], Mm[r0 * 4 + c1], Mm[r0 * 4 + c2], Mm[r1 * 4 + c0], Mm[r1 * 4 + c1], Mm[r1 * 4 + c2],
                    Mm[r2 * 4 + c0], Mm[r2 * 4 + c1], Mm[r2 * 4 + c2]);
      Cof[r * 4 + c] = ((r + c) & 1) ? cneg(d) : d;
    }
    __syncwarp();
    const cplx p0 = cmul(Mm[0], Cof[0]), p1 = cmul(Mm[1], Cof[1]), p2 = cmul(Mm[2], Cof[2]), p3 = cmul(Mm[3], Cof[3]);
    const cplx det = cadd(cadd(p0, p1), cadd(p2, p3));
    const double p_acc = e_dS * det.x;
    int acc_flag = p_acc > 0.5;
    if (lane == 0) s_accept = acc_flag;
    if (lane < 16) {
      const double id = 1.0 / (det.x * det.x + det.y * det.y);
      const cplx dinv = cmake(det.x * id, -det.y * id);
      Minv[r * 4 + c] = cmul(Cof[c * 4 + r], dinv);
    }
    __syncwarp();
    if (lane < 16) g4e[lane] = cmake(0.1 * lane + 1e-3 * Minv[lane].x, 0.01 + 1e-3 * Minv[lane].y);   // feed back: next iteration depends on this one
    __syncwarp();
  }
  long long t1 = clock64();
  if (lane == 0) { out[0] = (double)(t1 - t0) / iters; out[1] = g4e[3].x; stop = 1; }
}
int main() {
  double* out; cudaMalloc(&out, 8 * 1024); double h[2];
  for (int nthreads : {32, 256}) {
    decision_probe<<<1, nthreads>>>(out, 2000, 0);
    cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
    printf("decision role, %d threads in the CTA (%s): %.0f cycles per decision  [%s]\n", nthreads,
           nthreads == 32 ? "alone" : "7 other warps spinning on FP64", h[0], cudaGetErrorString(cudaDeviceSynchronize()));
  }
}
